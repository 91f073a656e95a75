"""DuckDB BitPacking column segments: oracle pinned to the reference, GPU decode pinned to the oracle.

Fixture tests/golden/bitpacking_segments.npz = compressed segments the REFERENCE binary wrote (lifted from a
checkpointed database file by tests/golden/make_bitpacking_golden.py) + the values its SELECT returns.
CPU tests: the oracle decoder reproduces those values; its 32-value unpack equals the reference's own
fastunpack compiled into oracle/_ref; the oracle's encoder (test data at scale) is byte-identical to the
reference's writer.  GPU tests (C-ABI cubit_gpu_upload_column_segments): bit-exact against golden and oracle.
"""
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FORCED = {"f_for": "for", "f_delta": "delta_for", "f_delta32": "delta_for", "f_for32": "for", "f_const": "constant",
          "f_wide": "for"}


@pytest.fixture(scope="module")
def bp():
    d = np.load(os.path.join(ROOT, "tests", "golden", "bitpacking_segments.npz"))
    cols, meta, blob = d["seg_column"], d["seg_meta"], d["blob"]
    names = sorted(set(cols.tolist()), key=lambda n: list(cols).index(n))
    return {n: {"values": d["values_" + n], "meta": meta[cols == n], "blob": blob} for n in names}


def segments_of(col):
    return [(int(k), int(s), int(c), col["blob"][o:o + b]) for k, s, c, o, b, _e in col["meta"]]


# ------------------------------------------------------------------ CPU: oracle vs reference
def test_oracle_decodes_reference_segments(bp):
    seen = {}
    kinds = set()
    for name, col in bp.items():
        elem = int(col["meta"][0][5])
        got, modes = oracle.decode_column_segments(col["blob"], col["meta"], elem, len(col["values"]))
        assert got.dtype == col["values"].dtype and np.array_equal(got, col["values"]), name
        for k, v in modes.items():
            seen[k] = seen.get(k, 0) + v
        kinds |= set(col["meta"][:, 0].tolist())
    # every BitpackingMode and every segment kind is exercised by the fixture
    assert set(seen) == {"constant", "constant_delta", "delta_for", "for"} and kinds == {0, 1, 2}


def test_oracle_unpack_equals_reference_fastunpack():
    """oracle/_ref/libfastpfor_ref.so = /root/reference/third_party/fastpforlib/bitpacking.cpp + our shim"""
    ref = oracle.ref_fastpfor()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference; build container only)")
    import ctypes as C
    L = oracle.lib()
    rng = np.random.default_rng(5)
    u32p, u64p, u8p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
    for bit in range(0, 65):
        vals = rng.integers(0, 2**63, 32, dtype=np.uint64) * 2 + rng.integers(0, 2, 32, dtype=np.uint64)
        if bit < 64:
            vals &= np.uint64((1 << bit) - 1)
        packed = np.zeros(64 + 2, dtype=np.uint32)
        ref.ref_fastpack64(vals.ctypes.data_as(u64p), packed.ctypes.data_as(u32p), bit)   # reference packs
        want = np.zeros(32, dtype=np.uint64)
        ref.ref_fastunpack64(packed.ctypes.data_as(u32p), want.ctypes.data_as(u64p), bit)  # reference unpacks
        got = np.zeros(32, dtype=np.uint64)
        L.oracle_bp_unpack_group64(packed.view(np.uint8).ctypes.data_as(u8p), got.ctypes.data_as(u64p), bit)
        assert np.array_equal(want, vals) and np.array_equal(got, want), bit
    for bit in range(0, 33):
        vals = rng.integers(0, 2**32, 32, dtype=np.uint64).astype(np.uint32)
        if bit < 32:
            vals &= np.uint32((1 << bit) - 1)
        packed = np.zeros(32 + 2, dtype=np.uint32)
        ref.ref_fastpack32(vals.ctypes.data_as(u32p), packed.ctypes.data_as(u32p), bit)
        want = np.zeros(32, dtype=np.uint32)
        ref.ref_fastunpack32(packed.ctypes.data_as(u32p), want.ctypes.data_as(u32p), bit)
        got = np.zeros(32, dtype=np.uint32)
        L.oracle_bp_unpack_group32(packed.view(np.uint8).ctypes.data_as(u8p), got.ctypes.data_as(u32p), bit)
        assert np.array_equal(want, vals) and np.array_equal(got, want), bit


def test_oracle_encoder_is_byte_identical_to_the_reference_writer(bp):
    checked = 0
    for name, col in bp.items():
        for kind, start, count, seg in segments_of(col):
            if kind != 1:
                continue
            vals = col["values"][start:start + count]
            mine = oracle.bitpacking_encode(vals, FORCED.get(name, "auto"))
            assert len(mine) == len(seg), (name, start)
            diff = np.nonzero(mine != seg)[0]
            if count % 32 == 0:
                assert len(diff) == 0, (name, start)
            elif len(diff):
                # the reference packs the ragged last 32-value group from an uninitialised stack buffer
                # (bitpacking.hpp:43-58): only that group's padding values may differ
                n_grp = (count + 2047) // 2048
                meta_off = len(seg) - 4 * n_grp
                assert diff.min() >= meta_off - 8 - 32 * 8 and diff.max() < meta_off, (name, start)
            back, _ = oracle.bitpacking_decode(mine, vals.dtype.itemsize, count)
            assert np.array_equal(back, vals)
            checked += 1
    assert checked >= 15


@pytest.mark.parametrize("dtype", [np.int64, np.int32])
def test_oracle_encode_decode_round_trip(dtype):
    rng = np.random.default_rng(11)
    bits = np.dtype(dtype).itemsize * 8
    for width in list(range(0, bits - 1, 3)) + [bits - 2]:
        n = int(rng.integers(1, 3 * 2048 + 7))
        lo = int(rng.integers(-2**(bits - 2), 2**(bits - 2)))
        span = 1 << width
        vals = (lo + rng.integers(0, span, n, dtype=np.int64)).clip(-2**(bits - 1), 2**(bits - 1) - 1).astype(dtype)
        for mode in ("auto", "for", "delta_for"):
            seg = oracle.bitpacking_encode(vals, mode)
            back, _ = oracle.bitpacking_decode(seg, np.dtype(dtype).itemsize, n)
            assert np.array_equal(back, vals), (width, mode)
    # a group spanning the whole signed range cannot be frame-of-reference coded (Flush returns false,
    # bitpacking.cpp:288: the reference then stores the column Uncompressed)
    info = np.iinfo(dtype)
    with pytest.raises(ValueError):
        oracle.bitpacking_encode(np.array([info.min, info.max, 0], dtype=dtype))


# ------------------------------------------------------------------ GPU: C-ABI vs oracle / golden
@pytest.mark.gpu
def test_gpu_decodes_reference_segments(cubit, bp):
    for name, col in bp.items():
        want = col["values"]
        elem = want.dtype.itemsize
        _, modes = oracle.decode_column_segments(col["blob"], col["meta"], elem, len(want))
        with cubit.CubitTable(len(want)) as t:
            info = t.upload_column_segments(7, elem, segments_of(col))
            got = t.download_column(7)
            assert got.dtype == want.dtype and np.array_equal(got, want), name
            hist = {oracle.BP_MODES[i]: int(info.mode_groups[i]) for i in range(6) if info.mode_groups[i]}
            n_const_segs = int((col["meta"][:, 0] == 2).sum())
            if n_const_segs:
                modes["constant"] = modes.get("constant", 0) + n_const_segs
            assert hist == modes, name
            assert info.h2d_bytes == sum(int(b) if k != 2 else elem for k, _s, _c, _o, b, _e in col["meta"])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.int64, np.int32])
def test_gpu_decode_at_scale_then_scan_and_probe(cubit, dtype):
    """3.2 M rows in ~100 segments of every width and mode → decode on the GPU → build an index on one decoded
    column, probe another: row IDs / values / SUM equal the oracle's on the ORIGINAL arrays"""
    rng = np.random.default_rng(3)
    n = 3_200_017
    bits = np.dtype(dtype).itemsize * 8
    grp = np.arange(n) // 2048
    width = (grp * 7) % (bits - 1)
    lo = rng.integers(-2**(bits - 3), 2**(bits - 3), n // 2048 + 1)[grp]
    vals = (lo + (rng.random(n) * (2.0 ** width)).astype(np.int64)).astype(dtype)
    vals[5 * 2048:6 * 2048] = 99                                     # CONSTANT group
    vals[8 * 2048:9 * 2048] = (np.arange(2048) * 5 - 7).astype(dtype)  # CONSTANT_DELTA group
    srt = np.cumsum(rng.integers(0, 50, n)).astype(dtype)              # DELTA_FOR wins on sorted data
    key = rng.integers(0, 9, n).astype(np.int32)
    with cubit.CubitTable(n, row_base=1 << 20) as t:
        i1 = t.upload_column_segments(0, vals.dtype.itemsize, oracle.encode_column_segments(vals, 30720))
        i2 = t.upload_column_segments(1, srt.dtype.itemsize, oracle.encode_column_segments(srt, 122880))
        i3 = t.upload_column_segments(2, 4, oracle.encode_column_segments(key, 122880, "for"))
        assert np.array_equal(t.download_column(0), vals)
        assert np.array_equal(t.download_column(1), srt)
        assert np.array_equal(t.download_column(2), key)
        assert i1.mode_groups[2] >= 1 and i1.mode_groups[3] >= 1 and i2.mode_groups[4] > 1000 and i3.mode_groups[5] > 1000
        assert i1.h2d_bytes < vals.nbytes and i2.h2d_bytes < srt.nbytes // 4
        if dtype == np.int64:
            ix = t.create_index(9)
            t.build_index(ix, 2, 0)
            bv = oracle.build_index(key, 0, 9)
            want = oracle.decode(oracle.merge([[bv[2], bv[5]]]), 1 << 20)
            with t.query([[(ix, 2), (ix, 5)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0, 1], agg=cubit.AGG_SUM,
                         agg_a=1) as r:
                ids, (a, b) = r.fetch()
                assert np.array_equal(ids, want)
                assert np.array_equal(a, oracle.probe(want, vals, 1 << 20))
                assert np.array_equal(b, oracle.probe(want, srt, 1 << 20))
                assert r.sum == oracle.sum_i64(b)


@pytest.mark.gpu
def test_gpu_rejects_malformed_segments(cubit):
    vals = (np.arange(5000, dtype=np.int64) * 37) % 1001
    good = oracle.bitpacking_encode(vals, "for")
    with cubit.CubitTable(5000) as t:
        t.upload_column_segments(0, 8, [(1, 0, 5000, good)])
        assert np.array_equal(t.download_column(0), vals)

        def bad(segs, elem=8):
            with pytest.raises(cubit.CubitError):
                t.upload_column_segments(0, elem, segs)
        bad([(1, 0, 5000, good[:len(good) - 8])])                         # truncated
        bad([(1, 0, 4000, good)])                                         # does not cover the table
        bad([(1, 8, 5000, good)])                                         # gap at the start
        bad([(1, 0, 2500, good), (1, 2600, 2400, good)])                  # gap in the middle
        x = good.copy()
        x[:8] = np.frombuffer(np.uint64(len(good) + 64).tobytes(), dtype=np.uint8)
        bad([(1, 0, 5000, x)])                                            # metadata end past the segment
        x = good.copy()
        x[len(good) - 1] = 9                                              # first group's mode byte → invalid mode
        bad([(1, 0, 5000, x)])
        x = good.copy()
        x[8 + 8] = 77                                                     # width 77 > 64
        bad([(1, 0, 5000, x)])
        x = good.copy()
        x[len(good) - 4:len(good) - 1] = 0xff                             # group offset far outside
        bad([(1, 0, 5000, x)])
        bad([(3, 0, 5000, good)])                                         # unknown kind
        bad([(0, 0, 5000, vals[:100])])                                   # short uncompressed segment
        bad([(1, 0, 5000, good)], elem=2)
        # the table still answers with the last good upload
        assert np.array_equal(t.download_column(0), vals)
