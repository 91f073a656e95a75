"""The narrow-wire DataChunk hand-off (include/cubit_gpu_wire.h, csrc/cubit_wire.cu).

CPU: the host-side unpacker against wires written by an independent numpy encoder of the documented format
(every width, negative bases, 4-byte outputs, ragged last chunk, malformed input).
GPU: wires written by the device equal the wide hand-off row for row (which the parity tests pin to the oracle),
the widths are the narrowest that hold each chunk, and cubit_gpu_drain delivers every DataChunk exactly once with
usable batch indexes — single shard, several shards, several workers."""
import numpy as np
import pytest

import oracle

CHUNK = 2048
SLOT = 16384
MAGIC = 0x45524957


BITMAP = 255


def _frame(part, ascending):
    """(width field, payload bytes) of one chunk by the documented rule"""
    lo, hi = int(part.min()), int(part.max())
    rng = hi - lo
    w = 0 if rng == 0 else 1 if rng < 256 else 2 if rng < 65536 else 4 if rng < 2**32 else 8
    words = rng // 64 + 1
    if ascending and words * 16 <= len(part) * w:
        return BITMAP | (words << 8), words * 8
    return w, len(part) * w


def _encode_wire(streams, widths=None, ascending=()):
    """independent encoder of the format: streams = list of int64/int32 arrays of equal length"""
    n = len(streams[0])
    c = (n + CHUNK - 1) // CHUNK
    s = len(streams)
    data_off = (64 + s * c * 24 + 255) & ~255
    wire = np.zeros(data_off + s * c * SLOT, dtype=np.uint8)
    hdr = np.zeros(8, dtype=np.uint64)
    hdr[0] = MAGIC | (s << 32)
    hdr[1], hdr[2], hdr[3] = n, c, data_off
    wire[:64] = hdr.view(np.uint8)
    at = 0
    for ci in range(c):
        for si, a in enumerate(streams):
            wire[32 + si] = a.dtype.itemsize
            part = a.astype(np.int64)[ci * CHUNK:(ci + 1) * CHUNK]
            lo = int(part.min())
            w, _ = _frame(part, si in ascending and widths is None)
            if widths is not None:
                w = max(w, widths[si])
            d = 64 + (ci * s + si) * 24
            wire[d:d + 16] = np.array([lo, at], dtype=np.int64).view(np.uint8)
            wire[d + 16:d + 24] = np.array([w, len(part)], dtype=np.uint32).view(np.uint8)
            o = data_off + at
            if w & 255 == BITMAP:
                bits = np.zeros((w >> 8) * 64, dtype=np.uint8)
                bits[part - lo] = 1
                nbytes = (w >> 8) * 8
                wire[o:o + nbytes] = np.packbits(bits, bitorder="little")
            else:
                nbytes = len(part) * w
                if w:
                    delta = (part.view(np.uint64) - np.uint64(lo & (2**64 - 1))).astype(
                        {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w])
                    wire[o:o + nbytes] = delta.view(np.uint8)
            at += (nbytes + 15) & ~15
    return wire


def _dir(wire, n_streams, n):
    """→ base[stream, chunk], width field[stream, chunk], n[stream, chunk]; checks the frames are back to back"""
    c = (n + CHUNK - 1) // CHUNK
    d = wire[64:64 + n_streams * c * 24]
    q = d.view(np.int64).reshape(c, n_streams, 3)
    wn = d.view(np.uint32).reshape(c, n_streams, 6)
    width, cnt = wn[:, :, 4], wn[:, :, 5]
    nbytes = np.where((width & 255) == BITMAP, (width >> 8) * 8, cnt * width)
    padded = (nbytes.astype(np.int64) + 15) & ~15
    assert np.array_equal(q[:, :, 1].ravel(), np.concatenate([[0], np.cumsum(padded.ravel())[:-1]]))
    return q[:, :, 0].T, width.T, cnt.T


@pytest.mark.parametrize("n", [1, 2047, 2048, 2049, 10_000])
def test_unpack_matches_independent_encoder(cubit, n):
    rng = np.random.default_rng(n)
    ids = np.sort(rng.choice(50 * n + 10, size=n, replace=False)).astype(np.int64) + 7_000_000_000
    const = np.full(n, -42, dtype=np.int64)
    small = rng.integers(-100, 100, n).astype(np.int64)
    mid = rng.integers(-2**31 + 5, 2**31 - 5, n).astype(np.int64)
    huge = rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64)
    i32 = rng.integers(-2**31, 2**31 - 1, n).astype(np.int32)
    streams = [ids, const, small, mid, huge, i32]
    wire = _encode_wire(streams)
    assert len(wire) == cubit.wire_bytes(n, len(streams))
    for si, a in enumerate(streams):
        got = np.concatenate([cubit.wire_unpack(wire, si, c, a.dtype) for c in range((n + CHUNK - 1) // CHUNK)])
        assert np.array_equal(got, a), si
    assert len(cubit.wire_unpack(wire, 0, (n + CHUNK - 1) // CHUNK)) == 0  # past the window: empty, not an error
    # dense ascending streams as bitmap frames (one row in 2 or 4: bitmap; one in 50: deltas)
    for step in (2, 4, 50):
        asc = (np.cumsum(rng.integers(1, 2 * step, n)) + 2**40).astype(np.int64)
        w2 = _encode_wire([asc, small], ascending=(0,))
        _, width, _ = _dir(w2, 2, n)
        assert ((width[0] & 255) == BITMAP).any() == (step < 50 and n >= 2047), (step, n)
        got = np.concatenate([cubit.wire_unpack(w2, 0, c) for c in range((n + CHUNK - 1) // CHUNK)])
        assert np.array_equal(got, asc)
        assert np.array_equal(cubit.wire_unpack(w2, 0, 0, np.int32), asc[:CHUNK].astype(np.int32))
        assert np.array_equal(cubit.wire_unpack(w2, 1, 0), small[:CHUNK])
    # a wider-than-needed width is still a valid wire (what the CPU mock of the ABI writes)
    wide = _encode_wire(streams, widths=[8] * len(streams))
    for si, a in enumerate(streams):
        assert np.array_equal(cubit.wire_unpack(wide, si, 0, a.dtype), a[:CHUNK])
    # int64 stream read as 4-byte values: the low halves (what a narrowing caller would get)
    assert np.array_equal(cubit.wire_unpack(wire, 2, 0, np.int32), small[:CHUNK].astype(np.int32))


def test_unpack_property_random_streams(cubit):
    """hypothesis: any mix of value ranges, signs, chunk counts and ragged tails survives encode → unpack bit for bit"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 3 * CHUNK + 17), st.integers(0, 63), st.integers(-2**62, 2**62), st.integers(0, 2**32 - 1),
           st.booleans())
    def run(n, bits, center, seed, ascending):
        rng = np.random.default_rng(seed)
        span = 1 << bits
        a = (center + rng.integers(0, span, n, dtype=np.int64)).astype(np.int64)
        if ascending:
            a = np.unique(a)                      # strictly ascending: may travel as bitmap frames
            n = len(a)
        b = rng.integers(-2**31, 2**31 - 1, n).astype(np.int32)
        wire = _encode_wire([a, b], ascending=(0,) if ascending else ())
        nch = (n + CHUNK - 1) // CHUNK
        assert np.array_equal(np.concatenate([cubit.wire_unpack(wire, 0, c) for c in range(nch)]), a)
        assert np.array_equal(np.concatenate([cubit.wire_unpack(wire, 1, c, np.int32) for c in range(nch)]), b)
    run()


def test_scalar_widening_path_gives_the_same_answers():
    """the AVX2 loops are an optimisation of the scalar ones: run the encoder test again with them switched off"""
    import os
    import subprocess
    import sys
    env = dict(os.environ, CUBIT_WIRE_SCALAR="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-k",
                        "unpack_matches_independent_encoder"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    assert r.returncode == 0 and "5 passed" in r.stdout, r.stdout[-2000:]


def test_unpack_rejects_malformed_wires(cubit):
    a = np.arange(5000, dtype=np.int64)
    wire = _encode_wire([a])
    bad = wire.copy()
    bad[0] ^= 0xff
    with pytest.raises(cubit.CubitError):
        cubit.wire_unpack(bad, 0, 0)
    with pytest.raises(cubit.CubitError):
        cubit.wire_unpack(wire, 1, 0)  # stream out of range
    bad = wire.copy()
    bad[64 + 16] = 3  # width 3 does not exist
    with pytest.raises(cubit.CubitError):
        cubit.wire_unpack(bad, 0, 0)
    bad = wire.copy()
    bad[64 + 20:64 + 24] = np.array([4096], dtype=np.uint32).view(np.uint8)  # more values than a chunk holds
    with pytest.raises(cubit.CubitError):
        cubit.wire_unpack(bad, 0, 0)
    # a bitmap frame that holds fewer set bits than dir.n, or claims more words than a slot has
    dense = _encode_wire([np.arange(0, 9000, 3, dtype=np.int64)], ascending=(0,))
    assert dense[64 + 16] == BITMAP
    bad = dense.copy()
    off = int(dense[24:32].view(np.uint64)[0])
    bad[off] &= 0xfe
    with pytest.raises(cubit.CubitError):
        cubit.wire_unpack(bad, 0, 0)
    bad = dense.copy()
    bad[64 + 16:64 + 20] = np.array([BITMAP | (4096 << 8)], dtype=np.uint32).view(np.uint8)
    with pytest.raises(cubit.CubitError):
        cubit.wire_unpack(bad, 0, 0)


# ---------------------------------------------------------------------------------------------------- GPU
def _wire_table(cubit, n, card, seed, devices=None, row_base=0):
    rng = np.random.default_rng(seed)
    key = rng.integers(0, card, n).astype(np.int32)
    cols = {
        0: rng.integers(-2**40, 2**40, n).astype(np.int64),          # wide range: 8-byte deltas
        2: (np.arange(n, dtype=np.int64) + row_base),                  # = row id: as narrow as the row IDs
        3: rng.integers(-2**31, 2**31 - 1, n).astype(np.int32),      # 4-byte column
        4: rng.integers(1000, 1200, n).astype(np.int64),              # one byte of range
        5: np.full(n, 77, dtype=np.int64),                            # constant: width 0
        6: rng.standard_normal(n).astype(np.float64),                 # doubles travel as bit patterns
    }
    t = cubit.CubitTable(n, row_base=row_base, seg_bits=65536, devices=devices)
    for c, a in cols.items():
        t.upload_column(c, a)
    t.upload_column(1, key)
    ix = t.create_index(card)
    t.build_index(ix, 1, 0)
    return t, ix, key, cols


@pytest.mark.gpu
@pytest.mark.parametrize("n,vals", [(300_007, [3]), (1_000_003, [0, 1, 2, 3, 4, 5, 6]), (70_000, [])])
def test_device_written_wire_equals_wide_fetch(cubit, n, vals, monkeypatch):
    card = 8
    bitmaps = len(vals) == 7  # bitmap frames for the row IDs are opt-in: exercised on the dense selection
    if bitmaps:
        monkeypatch.setenv("CUBIT_WIRE_BITMAP", "1")
    t, ix, key, cols = _wire_table(cubit, n, card, 11 + n, row_base=65536 * 3)
    order = [0, 2, 3, 4, 5, 6]
    dts = [cols[c].dtype for c in order]
    groups = [[(ix, v) for v in vals]] if vals else [[(ix, 1)], [(ix, 2)]]  # the second one is empty (AND of disjoint)
    with t.query(groups, flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=order) as r:
        want_ids = np.flatnonzero(np.isin(key, vals)) + 65536 * 3 if vals else np.empty(0, np.int64)
        assert r.count == len(want_ids)
        ids, cv = r.fetch()
        assert np.array_equal(ids, want_ids)
        for c, a in zip(order, cv):
            assert np.array_equal(a.view(np.uint8), cols[c][want_ids - 65536 * 3].view(np.uint8))
        cnt = r.count
        for off, m in [(0, cnt), (1, min(cnt - 1, 5000)), (cnt // 2 + 3, min(cnt - cnt // 2 - 3, 2048)), (cnt, 0)]:
            if m < 0 or off > cnt:
                continue
            with cubit.HostBuffer(cubit.wire_bytes(m, 1 + len(order))) as hb:
                wire = hb.array
                tk = r.fetch_wire_async(off, m, wire)
                r.fetch_wait(tk)
                nch = (m + CHUNK - 1) // CHUNK
                got = np.concatenate([cubit.wire_unpack(wire, 0, c) for c in range(nch)]) if nch else np.empty(0, np.int64)
                assert np.array_equal(got, ids[off:off + m])
                for si, (c, dt) in enumerate(zip(order, dts)):
                    g = np.concatenate([cubit.wire_unpack(wire, 1 + si, ch, dt) for ch in range(nch)]) if nch else np.empty(0, dt)
                    assert np.array_equal(g.view(np.uint8), cv[si][off:off + m].view(np.uint8)), (c, off, m)
                if m:
                    base, width, cn = _dir(wire, 1 + len(order), m)
                    assert cn.sum(axis=1).tolist() == [m] * (1 + len(order))
                    # every width is the narrowest that holds the chunk's range
                    streams = [ids[off:off + m]] + [a[off:off + m] for a in cv]
                    for si, a in enumerate(streams):
                        v = a.view(np.int64) if a.dtype.itemsize == 8 else a.astype(np.int64)
                        for ch in range(nch):
                            p = v[ch * CHUNK:(ch + 1) * CHUNK]
                            w, _ = _frame(p, si == 0 and bitmaps)  # only the row-ID stream may travel as a bitmap
                            assert width[si, ch] == w and base[si, ch] == int(p.min()), (si, ch)
                    # stream 5 is the constant column; a selection of 7 values in 8 ships its row IDs as bitmaps
                    assert (width[5] == 0).all()
                    assert ((width[0] & 255) == BITMAP).all() == bitmaps
                    assert cubit.load_library().cubit_gpu_wire_payload_bytes(wire.ctypes.data) < m * 8 * (1 + len(order))
    t.close()


@pytest.mark.gpu
def test_wire_argument_errors(cubit):
    t, ix, key, cols = _wire_table(cubit, 100_000, 4, 5)
    with t.query([[(ix, 1)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0]) as r:
        n = r.count
        pageable = np.zeros(cubit.wire_bytes(n, 2), dtype=np.uint8)
        with pytest.raises(cubit.CubitError, match="page-locked"):
            r.fetch_wire_async(0, n, pageable)
        with cubit.HostBuffer(cubit.wire_bytes(n, 2)) as hb:
            with pytest.raises(cubit.CubitError, match="needed"):
                r.fetch_wire_async(0, n, hb.array[:1024])
            with pytest.raises(cubit.CubitError, match="outside result"):
                r.fetch_wire_async(1, n, hb.array)
            with pytest.raises(cubit.CubitError, match="projected columns"):
                r.fetch_wire_async(0, n, hb.array, n_cols=2)
    with t.query([[(ix, 1)]], flags=cubit.Q_VALUES, cols=[0]) as r:
        with cubit.HostBuffer(cubit.wire_bytes(r.count, 2)) as hb:
            with pytest.raises(cubit.CubitError, match="row IDs"):
                r.fetch_wire_async(0, r.count, hb.array)
            r.fetch_wait(r.fetch_wire_async(0, r.count, hb.array, rowids=False))  # values alone are fine
            assert np.array_equal(cubit.wire_unpack(hb.array, 0, 0), cols[0][np.flatnonzero(key == 1)][:CHUNK])
        with pytest.raises(cubit.CubitError):
            r.drain(rowids=True)
    with t.query([[(ix, 1)]], flags=0, agg=cubit.AGG_SUM, agg_a=0) as r:
        with pytest.raises(cubit.CubitError):
            r.drain(rowids=False, n_cols=0)
    with t.query([[(ix, 1)], [(ix, 2)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0]) as r:  # empty result
        st = r.drain(threads=4)
        assert r.count == 0 and st.rows == 0 and st.chunks == 0 and st.sum_rowids == 0
    t.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shards,threads,window", [(1, 1, 0), (1, 4, 4096), (3, 3, 8192), (2, 8, 2048)])
def test_drain_delivers_every_chunk_once_in_batch_order(cubit, shards, threads, window):
    ndev = cubit.device_count()
    devices = [i % ndev for i in range(shards)] if shards > 1 else None
    n, card = 700_001, 6
    base = 65536 * 2
    t, ix, key, cols = _wire_table(cubit, n, card, 77 + shards, devices=devices, row_base=base)
    order = [0, 3, 2]
    with t.query([[(ix, 0), (ix, 2), (ix, 5)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=order) as r:
        want = np.flatnonzero(np.isin(key, [0, 2, 5])) + base
        assert r.count == len(want)
        # checksum consumer (the bench's consumer): reads every delivered value
        st = r.drain(threads=threads, window_rows=window)
        assert st.rows == len(want) and st.chunks >= (len(want) + CHUNK - 1) // CHUNK
        assert st.sum_rowids == int(want.sum()) % 2**64
        loc = want - base
        assert st.sum_cols[0] == int(cols[0][loc].view(np.uint64).sum(dtype=np.uint64))
        assert st.sum_cols[1] == int(cols[3][loc].view(np.uint32).sum(dtype=np.uint64))
        assert st.sum_cols[2] == int(cols[2][loc].sum()) % 2**64
        assert 0 < st.wire_bytes < st.wide_bytes == len(want) * (8 + 8 + 4 + 8)
        assert 1 <= st.workers <= threads
        # callback consumer: chunks sorted by (batch index, row offset) are the result in order
        got = []

        def fn(worker, batch, row_off, ids, cv):
            got.append((batch, row_off, ids, cv))
            return 0
        st2 = r.drain(threads=threads, window_rows=window, fn=fn)
        assert st2.rows == len(want)
        got.sort(key=lambda g: (g[0], g[1]))
        assert [g[1] for g in got] == sorted(g[1] for g in got)  # batch order = row order
        assert np.array_equal(np.concatenate([g[2] for g in got]), want)
        for k, c in enumerate(order):
            assert np.array_equal(np.concatenate([g[3][k] for g in got]), cols[c][loc])
        assert all(len(g[2]) == CHUNK for g in got[:-1]) or shards > 1 or window
        # a consumer that gives up stops the drain with an error
        with pytest.raises(cubit.CubitError, match="stopped"):
            r.drain(threads=threads, window_rows=window, fn=lambda *a: 1)
        if shards > 1:
            # one wire = one device: a window inside one shard is written by that shard, one that straddles two is refused
            spans = [t.shard_info(i) for i in range(shards)]
            per_shard = [int(np.count_nonzero((loc >= r0) & (loc < r0 + nr))) for _, r0, nr in spans]
            ids_all, cv_all = r.fetch()
            with cubit.HostBuffer(cubit.wire_bytes(3 * CHUNK, 4)) as hb:
                off = per_shard[0] + 5                      # inside the second shard
                m = min(3 * CHUNK, per_shard[1] - 5)
                r.fetch_wait(r.fetch_wire_async(off, m, hb.array))
                got = np.concatenate([cubit.wire_unpack(hb.array, 0, c) for c in range((m + CHUNK - 1) // CHUNK)])
                assert np.array_equal(got, ids_all[off:off + m])
                g3 = np.concatenate([cubit.wire_unpack(hb.array, 2, c, np.int32) for c in range((m + CHUNK - 1) // CHUNK)])
                assert np.array_equal(g3, cv_all[1][off:off + m])
                with pytest.raises(cubit.CubitError, match="straddle"):
                    r.fetch_wire_async(per_shard[0] - 10, 100, hb.array)
                with pytest.raises(cubit.CubitError, match="outside result"):
                    r.fetch_wire_async(len(want) - 10, 100, hb.array)
    t.close()


@pytest.mark.gpu
def test_drain_with_nulls_passes_validity(cubit):
    """NULL-bearing projected column: the drain hands the chunk's ValidityMask words along"""
    n = 200_003
    rng = np.random.default_rng(9)
    key = rng.integers(0, 4, n).astype(np.int32)
    pay = rng.integers(0, 1000, n).astype(np.int64)
    valid = rng.random(n) > 0.1
    t = cubit.CubitTable(n, seg_bits=65536)
    t.upload_column(0, pay)
    bits = np.packbits(valid, bitorder="little")
    bits = np.concatenate([bits, np.zeros((-len(bits)) % 8, dtype=np.uint8)])
    t.upload_validity(0, bits.view(np.uint64))
    t.upload_column(1, key)
    ix = t.create_index(4)
    t.build_index(ix, 1, 0)
    with t.query([[(ix, 2)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0]) as r:
        want = np.flatnonzero(key == 2)
        st = r.drain(threads=2, window_rows=4096)
        assert st.rows == len(want) and st.sum_rowids == int(want.sum())
        seen = []

        def fn(worker, batch, row_off, ids, cv, vw):
            m = np.ones(len(ids), dtype=bool) if vw[0] is None else \
                np.unpackbits(vw[0].view(np.uint8), bitorder="little")[:len(ids)].astype(bool)
            seen.append((batch, row_off, ids, cv[0], m))
            return 0
        r.drain(threads=2, window_rows=4096, fn=fn, validity=True)
        seen.sort(key=lambda g: (g[0], g[1]))
        ids = np.concatenate([g[2] for g in seen])
        vals = np.concatenate([g[3] for g in seen])
        assert np.array_equal(ids, want)
        assert np.array_equal(np.concatenate([g[4] for g in seen]), valid[want])
        assert np.array_equal(vals[valid[want]], pay[want][valid[want]])  # NULL slots hold unspecified values
    t.close()


@pytest.mark.gpu
def test_concurrent_queries_and_drains_on_one_table(cubit):
    """four host threads, each running its own queries and draining them (two workers each) on ONE table at the same
    time — the hand-off's kernels share the table's copy streams and the per-result rings: every checksum must be right"""
    import threading
    n, card = 1_500_007, 6
    t, ix, key, cols = _wire_table(cubit, n, card, 404)
    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for it in range(6):
                vals = sorted(rng.choice(card, size=int(rng.integers(1, 4)), replace=False).tolist())
                want = np.flatnonzero(np.isin(key, vals))
                with t.query([[(ix, v) for v in vals]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0, 3]) as r:
                    st = r.drain(threads=2, window_rows=int(rng.choice([2048, 16384, 0])))
                    assert st.rows == len(want) and st.sum_rowids == int(want.sum())
                    assert st.sum_cols[0] == int(cols[0][want].view(np.uint64).sum(dtype=np.uint64))
                    assert st.sum_cols[1] == int(cols[3][want].view(np.uint32).sum(dtype=np.uint64))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    th = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for x in th:
        x.start()
    for x in th:
        x.join(timeout=300)
    assert not any(x.is_alive() for x in th), "a drain is stuck"
    assert not errors, errors
    t.close()
