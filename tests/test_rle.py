"""DuckDB RLE column segments (src/storage/compression/rle.cpp): oracle pinned to the reference, GPU decode pinned
to both.

Fixture tests/golden/rle_segments.npz = RLE segments the REFERENCE binary wrote (lifted from a checkpointed
database file by tests/golden/make_rle_golden.py) + the values its SELECT returns.  CPU: the oracle decoder
reproduces those values and the oracle encoder is byte-identical to the reference's writer.  GPU (C-ABI
cubit_gpu_upload_column_segments, kind CUBIT_SEG_RLE): bit-exact against golden and oracle, malformed segments
rejected on the host.
"""
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEG_RLE = 3


@pytest.fixture(scope="module")
def rle():
    d = np.load(os.path.join(ROOT, "tests", "golden", "rle_segments.npz"))
    cols, meta, blob = d["seg_column"], d["seg_meta"], d["blob"]
    names = sorted(set(cols.tolist()), key=lambda n: list(cols).index(n))
    return {n: {"values": d["values_" + n], "meta": meta[cols == n], "blob": blob} for n in names}


def test_oracle_decodes_and_reencodes_reference_segments(rle):
    longest = 0
    for name, col in rle.items():
        for start, count, off, nbytes, elem in col["meta"]:
            seg = col["blob"][off:off + nbytes]
            got, runs = oracle.rle_decode(seg, int(elem), int(count))
            want = col["values"][start:start + count]
            assert got.dtype == want.dtype and np.array_equal(got, want), name
            # the oracle's writer reproduces the reference's bytes: header, values, padding, run lengths
            enc = oracle.rle_encode(want)
            assert enc.tobytes() == seg.tobytes(), (name, int(start))
            cnt = np.frombuffer(seg.tobytes()[int(np.frombuffer(seg.tobytes()[:8], "<u8")[0]):], dtype="<u2")
            longest = max(longest, int(cnt.max()))
    assert longest == 65535  # the fixture holds runs split at the uint16 limit


def test_oracle_rejects_malformed_rle():
    seg = oracle.rle_encode(np.array([1, 1, 2, 3, 3, 3], dtype=np.int64))
    for bad in (seg[:4], seg[:-2]):                       # truncated header / missing last run length
        with pytest.raises(ValueError):
            oracle.rle_decode(bad, 8, 6)
    x = seg.copy()
    x[0] = 9                                              # unaligned run-length offset
    with pytest.raises(ValueError):
        oracle.rle_decode(x, 8, 6)
    with pytest.raises(ValueError):
        oracle.rle_decode(seg, 8, 7)                      # more rows than the runs hold


@pytest.mark.gpu
def test_gpu_decodes_reference_rle_segments(cubit, rle):
    for name, col in rle.items():
        n = len(col["values"])
        elem = int(col["meta"][0][4])
        t = cubit.CubitTable(n)
        segs = [(SEG_RLE, int(s), int(c), col["blob"][o:o + b + 64]) for s, c, o, b, _e in col["meta"]]  # bytes = upper bound
        info = t.upload_column_segments(0, elem, segs)
        got = t.download_column(0)
        assert np.array_equal(got, col["values"]), name
        assert info.rle_runs > 0 and info.n_launches == 1
        t.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.int64, np.int32])
def test_gpu_rle_random_runs_and_mixed_segment_kinds(cubit, dtype):
    rng = np.random.default_rng(11)
    # geometric run lengths incl. very long runs (split at 65535), values across the full range of the type
    lens = np.concatenate([rng.geometric(0.2, 20000), np.array([70000, 131073, 1, 65535, 65536]), rng.geometric(0.01, 500)])
    info_t = np.iinfo(dtype)
    vals = rng.integers(info_t.min, info_t.max, len(lens), dtype=dtype)
    col = np.repeat(vals, lens)
    n = len(col)
    # three segments: RLE, uncompressed, RLE (rows split at arbitrary points, as row groups do)
    a, b = n // 3 + 17, 2 * n // 3 + 5
    elem = col.dtype.itemsize
    segs = [(SEG_RLE, 0, a, oracle.rle_encode(col[:a])), (0, a, b - a, col[a:b].view(np.uint8)),
            (SEG_RLE, b, n - b, oracle.rle_encode(col[b:]))]
    t = cubit.CubitTable(n, seg_bits=32768)
    info = t.upload_column_segments(3, elem, segs)
    assert np.array_equal(t.download_column(3), col)
    assert info.rle_runs >= 2
    # probe through an index: the decoded column is an ordinary resident column
    key = (np.arange(n) % 5).astype(np.int64)
    t.upload_column(0, key)
    ix = t.create_index(5)
    t.build_index(ix, 0, 0)
    with t.query([[(ix, 2)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[3]) as r:
        ids, (v,) = r.fetch()
        assert np.array_equal(ids, np.nonzero(key == 2)[0]) and np.array_equal(v, col[ids])
    # malformed: run lengths cut short, zero-length run, offset outside the segment
    good = oracle.rle_encode(col[:a])
    with pytest.raises(cubit.CubitError):
        t.upload_column_segments(4, elem, [(SEG_RLE, 0, n, good)])            # runs cover a rows, not n
    z = good.copy()
    off = int(np.frombuffer(z[:8].tobytes(), "<u8")[0])
    z[off] = 0
    z[off + 1] = 0
    with pytest.raises(cubit.CubitError):
        t.upload_column_segments(4, elem, [(SEG_RLE, 0, a, z), (0, a, n - a, col[a:].view(np.uint8))])
    y = good.copy()
    y[:8] = np.frombuffer(np.uint64(len(good) + 800).tobytes(), dtype=np.uint8)
    with pytest.raises(cubit.CubitError):
        t.upload_column_segments(4, elem, [(SEG_RLE, 0, a, y), (0, a, n - a, col[a:].view(np.uint8))])
    t.close()
