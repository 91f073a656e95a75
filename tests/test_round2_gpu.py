"""GPU: the round-2 library features through the C-ABI, each against the CPU oracle (bit-exact):
incremental device-side delta ingestion + threshold-driven merge-back, compressed (roaring-style) indexes,
one table sharded over several devices behind one handle, asynchronous DataChunk hand-off, host-thread
concurrency on one table, device-side limbs for the multi-process reduce."""
import threading
import time

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _table(cubit, n, card, seed, seg_bits=65536, row_base=0, devices=None, compressed=False, col=None):
    rng = np.random.default_rng(seed)
    if col is None:
        col = rng.integers(0, card, n).astype(np.int32)
    pay = rng.integers(-2**40, 2**40, n).astype(np.int64)
    t = cubit.CubitTable(n, row_base=row_base, seg_bits=seg_bits, devices=devices)
    t.upload_column(0, pay)
    t.upload_column(1, col)
    ix = t.create_index(card, compressed=compressed)
    t.build_index(ix, 1, 0)
    return t, ix, col, pay, oracle.build_index(col, 0, card), rng


def _check_query(cubit, t, ix, groups, bv, dl, pay, row_base=0, extras=(0,)):
    og = [[bv[v] for v in grp] for grp in groups]
    od = [[dl.get(v) for v in grp] for grp in groups]
    q = oracle.merge(og, od)
    want = oracle.decode(q, row_base)
    wv = oracle.probe(want, pay, row_base)
    for extra in extras:
        with t.query([[(ix, v) for v in grp] for grp in groups], flags=cubit.Q_ROWIDS | cubit.Q_VALUES | cubit.Q_BITVECTOR | extra,
                     cols=[0], agg=cubit.AGG_SUM, agg_a=0) as r:
            assert r.count == len(want), (groups, extra)
            ids, (vals,) = r.fetch()
            assert np.array_equal(ids, want)
            assert np.array_equal(vals, wv)
            assert r.sum == oracle.sum_i64(wv)
            assert np.array_equal(r.bitvector(), q)
        with t.query([[(ix, v) for v in grp] for grp in groups], flags=extra, agg=cubit.AGG_SUM, agg_a=0) as r:
            assert r.count == len(want) and r.sum == oracle.sum_i64(wv)
    return want


# ------------------------------------------------------------------ deltas
@pytest.mark.parametrize("n,seg_bits", [(1, 65536), (65, 32768), (300_007, 32768), (1_000_003, 65536), (2_500_000, 131072)])
def test_incremental_delta_ingestion_matches_oracle(cubit, n, seg_bits):
    """add_delta in pieces (with duplicates that must cancel, across calls too), add_delta_pairs, set_delta on top:
    queries always see B_v XOR (net flips); merge-back gives the same answers with no pending entries left"""
    card = 12
    t, ix, col, pay, bv, rng = _table(cubit, n, card, 100 + n, seg_bits)
    t.set_merge_threshold(ix, 0)  # keep everything pending: the automatic merge-back has its own test
    flips = {v: [] for v in range(card)}

    def net(v):
        return oracle.delta_from_rows(np.asarray(flips[v], dtype=np.int64), n) if flips[v] else None

    # three incremental batches on values 3 and 7; the third repeats rows of the first (they cancel)
    a = rng.integers(0, n, max(1, n // 40))
    b = rng.integers(0, n, max(1, n // 60))
    t.add_delta(ix, 3, a)
    flips[3] += a.tolist()
    t.add_delta(ix, 7, b)
    flips[7] += b.tolist()
    c = np.concatenate([a[: max(1, len(a) // 3)], rng.integers(0, n, max(1, n // 90))])
    t.add_delta(ix, 3, c)
    flips[3] += c.tolist()
    # one UPDATE-statement-shaped batch: (old value, row) and (new value, row) pairs over many values
    rows = rng.integers(0, n, max(1, n // 30))
    old = col[rows].astype(np.uint32)
    new = ((old + 1) % card).astype(np.uint32)
    t.add_delta_pairs(ix, np.concatenate([old, new]), np.concatenate([rows, rows]))
    for r, o, w in zip(rows.tolist(), old.tolist(), new.tolist()):
        flips[o].append(r)
        flips[w].append(r)
    dl = {v: net(v) for v in range(card)}
    cases = [[[3]], [[7]], [[1, 3, 5, 7]], [[0, 1, 2, 3], [3, 4, 7]], [list(range(card))]]
    extras = (0, cubit.Q_UNFUSED, cubit.Q_FUSE_PROBE)
    for g in cases:
        _check_query(cubit, t, ix, g, bv, dl, pay, extras=extras)
    assert t.index_info(ix).delta_entries == sum(len(x) for x in flips.values())
    # set_delta REPLACES one value's pending rows, the others stay
    d9 = rng.integers(0, n, max(1, n // 70))
    t.set_delta(ix, 3, d9)
    flips[3] = d9.tolist()
    dl[3] = net(3)
    for g in cases:
        _check_query(cubit, t, ix, g, bv, dl, pay)
    t.set_delta(ix, 7, np.zeros(0, dtype=np.int64))
    flips[7] = []
    dl[7] = None
    _check_query(cubit, t, ix, [[7], [3, 7]], bv, dl, pay)
    # merge-back: same answers, nothing pending
    t.merge_deltas(ix)
    assert t.index_info(ix).delta_entries == 0
    merged = [oracle.merge([[bv[v]]], [[dl.get(v)]]) for v in range(card)]
    for v in range(card):
        assert np.array_equal(t.download_bitvector(ix, v), merged[v])
    for g in cases:
        _check_query(cubit, t, ix, g, merged, {}, pay)
    t.close()


def test_threshold_driven_merge_back(cubit):
    """pending deltas are folded into the bitvectors once they outgrow the threshold"""
    n, card = 600_000, 4
    t, ix, col, pay, bv, rng = _table(cubit, n, card, 7)
    t.set_merge_threshold(ix, 0.05)  # 5 % of a bitvector's bytes = 3750 bytes = 234 entries
    rows = rng.choice(n, size=200, replace=False)
    t.add_delta(ix, 1, rows)
    info = t.index_info(ix)
    assert info.delta_entries == 200 and info.auto_merges == 0
    more = rng.choice(n, size=100, replace=False)
    t.add_delta(ix, 1, more)  # 300 entries > 234: merged back automatically
    info = t.index_info(ix)
    assert info.delta_entries == 0 and info.auto_merges == 1
    want = oracle.merge([[bv[1]]], [[oracle.delta_from_rows(np.concatenate([rows, more]), n)]])
    assert np.array_equal(t.download_bitvector(ix, 1), want)
    with t.query([[(ix, 1)]], flags=cubit.Q_ROWIDS) as r:
        assert r.info.delta_entries == 0 and np.array_equal(r.fetch()[0], oracle.decode(want))
    t.close()


def test_delta_ingestion_throughput_10m_rows(cubit):
    """10 M flipped rows were 4.65 s through the host-side CSR build; the device-side ingestion has to
    take them in well under 50 ms of GPU+host time (pageable source arrays)"""
    n, card = 1_000_000_000 // 4, 100
    t = cubit.CubitTable(n)
    t.synth_column(1, 1, seed=0xC0B17, threshold=1 << 62, card=card, hot_lo=10, hot_n=10)
    ix = t.create_index(card)
    t.build_index(ix, 1, 0)
    t.drop_column(1)
    t.set_merge_threshold(ix, 0)
    rng = np.random.default_rng(3)
    rows = rng.integers(0, n, 10_000_000).astype(np.int64)
    vals = rng.integers(0, card, 10_000_000).astype(np.uint32)
    t.add_delta_pairs(ix, vals[:1000], rows[:1000])  # warm-up (pools, kernels)
    before = [t.bitvector_count(ix, v) for v in range(card)]
    t0 = time.perf_counter()
    t.add_delta_pairs(ix, vals, rows)
    with t.query([[(ix, 11)]], flags=0) as r:  # a scan right behind it waits for the ingestion on the stream
        cnt = r.count
    dt = time.perf_counter() - t0
    print("10M delta pairs ingested + first scan: %.1f ms" % (dt * 1e3))
    # net effect on B_11's popcount: rows listed an odd number of times flip
    m = np.concatenate([rows[:1000][vals[:1000] == 11], rows[vals == 11]])
    u, c = np.unique(m, return_counts=True)
    odd = u[(c & 1) == 1]
    b11 = t.download_bitvector(ix, 11)
    was_set = ((b11[odd >> 6] >> (odd & 63).astype(np.uint64)) & 1).astype(np.int64)
    assert cnt == before[11] + int((1 - 2 * was_set).sum())
    assert dt < 0.25, dt  # generous bound for a shared box; the measured figure is recorded in profiles/
    t.close()


# ------------------------------------------------------------------ compressed indexes
def _mixed_density_column(rng, n, card):
    """values whose bitvectors have empty, sparse (ARRAY), dense (BITMAP) and full segments"""
    col = rng.integers(0, card, n).astype(np.int32)
    seg = 65536
    if n > 3 * seg:
        col[seg:2 * seg] = 2                       # value 2: one FULL segment, others lose it
        col[2 * seg:3 * seg][col[2 * seg:3 * seg] == 5] = 6  # value 5: one EMPTY segment
    rare = rng.choice(n, size=max(1, n // 5000), replace=False)
    col[col == card - 1] = 0
    col[rare] = card - 1                           # value card-1: ARRAY containers everywhere
    return col


@pytest.mark.parametrize("n,seg_bits", [(70_000, 65536), (400_003, 32768), (1_500_000, 65536)])
def test_compressed_index_parity(cubit, n, seg_bits):
    rng = np.random.default_rng(n)
    card = 9
    col = _mixed_density_column(rng, n, card)
    t, ix, col, pay, bv, rng = _table(cubit, n, card, n + 1, seg_bits, compressed=True, col=col)
    vx = t.create_index(card)  # the same index verbatim, for the mixed-stream query below
    t.build_index(vx, 1, 0)
    info = t.index_info(ix)
    assert info.compressed == 1 and info.resident_bytes < info.verbatim_bytes
    for v in range(card):
        assert np.array_equal(t.download_bitvector(ix, v), bv[v]), v
        assert t.bitvector_count(ix, v) == oracle.popcount(bv[v])
    cases = [[[card - 1]], [[2]], [[5]], [[0, 2, 5, card - 1]], [[0, 1, 2], [2, 3, card - 1]], [list(range(card))]]
    for g in cases:
        _check_query(cubit, t, ix, g, bv, {}, pay, extras=(0, cubit.Q_UNFUSED, cubit.Q_FUSE_PROBE))
    # a query mixing container streams and verbatim streams
    want = oracle.decode(oracle.merge([[bv[card - 1], bv[2]], [bv[2], bv[3]]]))
    with t.query([[(ix, card - 1), (vx, 2)], [(ix, 2), (vx, 3)]], flags=cubit.Q_ROWIDS) as r:
        assert np.array_equal(r.fetch()[0], want)
    # pending deltas on containers (EMPTY, ARRAY and BITMAP segments all get flips), then merge-back
    t.set_merge_threshold(ix, 0)
    dl = {}
    for v in (5, card - 1, 0):
        rows = rng.integers(0, n, max(2, n // 300))
        t.add_delta(ix, v, rows)
        dl[v] = oracle.delta_from_rows(rows, n)
    for g in cases:
        _check_query(cubit, t, ix, g, bv, dl, pay)
    t.merge_deltas(ix)
    merged = [oracle.merge([[bv[v]]], [[dl.get(v)]]) for v in range(card)]
    for v in range(card):
        assert np.array_equal(t.download_bitvector(ix, v), merged[v]), v
    for g in cases:
        _check_query(cubit, t, ix, g, merged, {}, pay)
    # image round trip keeps the compressed form
    img = t.serialize_index(ix)
    ix2 = t.deserialize_index(img)
    assert t.index_info(ix2).compressed == 1
    for v in range(card):
        assert np.array_equal(t.download_bitvector(ix2, v), merged[v])
    t.close()


def test_compressed_index_uploads_and_append(cubit):
    """verbatim and WAH uploads into a compressed index; append rebuilds the containers from the source column"""
    n, card = 300_000, 6
    rng = np.random.default_rng(11)
    col = _mixed_density_column(rng, n, card)
    bv = oracle.build_index(col, 0, card)
    t = cubit.CubitTable(n)
    t.upload_column(1, col)
    ix = t.create_index(card, compressed=True)
    for v in range(card):
        if v % 2:
            wah, aval, anb = oracle.wah_encode(bv[v], n)
            t.upload_bitvector_wah(ix, v, wah, aval, anb)
        else:
            t.upload_bitvector(ix, v, bv[v])
    for v in range(card):
        assert np.array_equal(t.download_bitvector(ix, v), bv[v]), v
    # built from the column, then appended past several segment boundaries
    bx = t.create_index(card, compressed=True)
    t.build_index(bx, 1, 0)
    more = rng.integers(0, card, 200_000).astype(np.int32)
    t.append_rows({1: more})
    full = np.concatenate([col, more])
    bv2 = oracle.build_index(full, 0, card)
    for v in range(card):
        assert np.array_equal(t.download_bitvector(bx, v), bv2[v]), v
    want = oracle.decode(oracle.merge([[bv2[1], bv2[card - 1]]]))
    with t.query([[(bx, 1), (bx, card - 1)]], flags=cubit.Q_ROWIDS) as r:
        assert np.array_equal(r.fetch()[0], want)
    t.close()


def test_high_cardinality_compressed_index_fits(cubit):
    """the day-level l_shipdate shape: 2,526 bitvectors; verbatim this index is 2,526 * N/8 bytes, the
    containers hold it in ~2 bytes per row"""
    n, card = 6_000_000, 2526
    t = cubit.CubitTable(n)
    t.synth_column(1, 2, seed=99, card=card, hot_lo=0)
    ix = t.create_index(card, compressed=True)
    t.build_index(ix, 1, 0)
    info = t.index_info(ix)
    assert info.verbatim_bytes > 1_800_000_000
    assert info.resident_bytes < info.verbatim_bytes // 8, (info.resident_bytes, info.verbatim_bytes)
    col = t.download_column(1)
    vals = list(range(1000, 1060))  # a two-month range: 60 container streams
    want = np.flatnonzero((col >= 1000) & (col < 1060)).astype(np.int64)
    with t.query([[(ix, v) for v in vals]], flags=cubit.Q_ROWIDS) as r:
        assert np.array_equal(r.fetch()[0], want)
    t.close()


# ------------------------------------------------------------------ sharded table behind one handle
@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_sharded_table_equals_single_shard(cubit, n_shards):
    """cubit_gpu_create_sharded on one device named several times: every entry point fans out by row range and the
    answers equal the oracle's (and therefore the single-shard table's); with more GPUs the same code spreads out"""
    ndev = cubit.device_count()
    devices = [i % ndev for i in range(n_shards)]
    n, card, seg_bits = 1_000_003, 10, 32768
    t, ix, col, pay, bv, rng = _table(cubit, n, card, 31, seg_bits, row_base=seg_bits * 4, devices=devices)
    assert t.shard_count == n_shards
    spans = [t.shard_info(s) for s in range(n_shards)]
    assert spans[0][1] == 0 and sum(x[2] for x in spans) == n and all(x[1] % seg_bits == 0 for x in spans)
    for v in range(card):
        assert np.array_equal(t.download_bitvector(ix, v), bv[v])
        assert t.bitvector_count(ix, v) == oracle.popcount(bv[v])
    assert np.array_equal(t.download_column(0), pay)
    # deltas are routed to the shard that owns the row
    d4 = rng.integers(0, n, 5000)
    t.add_delta(ix, 4, d4)
    rows = rng.integers(0, n, 3000)
    vals = rng.integers(0, card, 3000).astype(np.uint32)
    t.add_delta_pairs(ix, vals, rows)
    flips = {v: rows[vals == v].tolist() for v in range(card)}
    flips[4] += d4.tolist()
    dl = {v: oracle.delta_from_rows(np.asarray(f, dtype=np.int64), n) for v, f in flips.items() if f}
    cases = [[[4]], [[1, 2, 3]], [[0, 1, 2, 3, 4], [4, 5, 6]], [list(range(card))]]
    for g in cases:
        want = _check_query(cubit, t, ix, g, bv, dl, pay, row_base=seg_bits * 4)
        # windowed fetch across shard boundaries + asynchronous hand-off
        with t.query([[(ix, v) for v in grp] for grp in g], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0]) as r:
            if len(want) > 5000:
                off = len(want) // 3
                ids, (v,) = r.fetch(off, 4000)
                assert np.array_equal(ids, want[off:off + 4000])
                oi, ov = np.empty(4000, dtype=np.int64), np.empty(4000, dtype=np.int64)
                tk = r.fetch_async(off + 1, 4000, oi, [ov])
                r.fetch_wait(tk)
                assert np.array_equal(oi, want[off + 1:off + 4001]) and np.array_equal(ov, pay[oi - seg_bits * 4])
    # stand-alone probe over rows of every shard
    some = np.sort(rng.choice(n, size=2001, replace=False)) + seg_bits * 4
    v, s = t.probe(0, some, want_sum=True)
    assert np.array_equal(v, pay[some - seg_bits * 4]) and s == oracle.sum_i64(v)
    # INSERT goes to the last shard
    more_c = rng.integers(0, card, 70_000).astype(np.int32)
    more_p = rng.integers(-99, 99, 70_000).astype(np.int64)
    t.merge_deltas(ix)
    t.append_rows({0: more_p, 1: more_c})
    merged = [oracle.merge([[bv[v]]], [[dl.get(v)]]) for v in range(card)]
    full_col = np.concatenate([col, more_c])
    # rows flipped by the merged deltas no longer follow `col`: compare the appended tail only through a fresh value
    tail = np.flatnonzero(more_c == 3).astype(np.int64) + n + seg_bits * 4
    with t.query([[(ix, 3)]], flags=cubit.Q_ROWIDS) as r:
        ids = r.fetch()[0]
        head = oracle.decode(merged[3], seg_bits * 4)
        assert np.array_equal(ids, np.concatenate([head, tail]))
    assert len(full_col) == t.n_rows
    t.close()


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2", reason="needs 2 GPUs")
def test_sharded_table_over_two_gpus(cubit):
    n, card = 3_000_017, 10
    t, ix, col, pay, bv, rng = _table(cubit, n, card, 77, devices=[0, 1])
    assert {t.shard_info(s)[0] for s in range(t.shard_count)} == {0, 1}
    for g in ([[1, 2, 3]], [[0, 1, 2, 3, 4], [4, 5, 6]]):
        _check_query(cubit, t, ix, g, bv, {}, pay)
    t.close()


# ------------------------------------------------------------------ hand-off and concurrency
def test_async_fetch_windows_overlap(cubit):
    """double-buffered hand-off: window i+1 is in flight while window i is checked"""
    import torch
    n = 4_000_000
    t, ix, col, pay, bv, rng = _table(cubit, n, 4, 5)
    want = oracle.decode(oracle.merge([[bv[0], bv[2]]]))
    win = 1 << 17
    bufs = [(torch.empty(win, dtype=torch.int64, pin_memory=True).numpy(), torch.empty(win, dtype=torch.int64, pin_memory=True).numpy())
            for _ in range(2)]
    with t.query([[(ix, 0), (ix, 2)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0]) as r:
        assert r.count == len(want)
        tickets = [None, None]
        nwin = (r.count + win - 1) // win
        tickets[0] = r.fetch_async(0, min(win, r.count), bufs[0][0], [bufs[0][1]])
        for w in range(nwin):
            if w + 1 < nwin:
                o = (w + 1) * win
                tickets[(w + 1) & 1] = r.fetch_async(o, min(win, r.count - o), bufs[(w + 1) & 1][0], [bufs[(w + 1) & 1][1]])
            r.fetch_wait(tickets[w & 1])
            m = min(win, r.count - w * win)
            assert np.array_equal(bufs[w & 1][0][:m], want[w * win:w * win + m])
            assert np.array_equal(bufs[w & 1][1][:m], pay[want[w * win:w * win + m]])
    t.close()


def test_threads_overlap_on_one_table(cubit):
    """latency-bound queries from several host threads overlap (the table lock covers planning and
    enqueueing only): 8 threads must push > 3x the queries per second of one thread, answers exact"""
    n, card = 6_001_215, 50  # config-1 shape
    t, ix, col, pay, bv, rng = _table(cubit, n, card, 9)
    expect = {v: (oracle.popcount(bv[v]), int(pay[col == v].sum())) for v in range(8)}

    def run(n_threads, per_thread):
        errs = []

        def work(k):
            try:
                for i in range(per_thread):
                    v = (k + i) % 8
                    with t.query([[(ix, v)]], flags=0, agg=cubit.AGG_SUM, agg_a=0) as r:
                        if (r.count, r.sum) != expect[v]:
                            errs.append((v, r.count, r.sum))
            except Exception as e:  # noqa: BLE001
                errs.append(repr(e))
        th = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
        t0 = time.perf_counter()
        [x.start() for x in th]
        [x.join() for x in th]
        dt = time.perf_counter() - t0
        assert not errs, errs[:3]
        return n_threads * per_thread / dt
    run(2, 50)
    one = run(1, 400)
    eight = run(8, 400)
    print("queries/s: 1 thread %.0f, 8 threads %.0f (%.2fx)" % (one, eight, eight / one))
    # ctypes releases the GIL inside the C call; the Python glue around it does not, so the ratio seen from Python
    # understates what C++ callers get (tests/cpp/host_scan_test.cpp ConcurrentQueriesOverlap measures and asserts
    # that one); here: no slowdown and exact answers under contention
    assert eight > 1.1 * one
    t.close()


def test_result_limbs_accumulate_on_device(cubit):
    """cubit_gpu_result_add_limbs: (count, 128-bit sum) → five int64 limbs added on the device (what bench.py hands to
    one NCCL all-reduce per step instead of a host round trip)"""
    import torch
    n = 500_000
    t, ix, col, pay, bv, rng = _table(cubit, n, 6, 13)
    t.upload_column(2, np.full(n, -(2**62) - 12345, dtype=np.int64))  # sums far outside 64 bits, negative
    acc = torch.zeros(10, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()  # the library adds on the table's own stream
    tot = [[0, 0], [0, 0]]
    for rep in range(3):
        for slot, col_id in enumerate((0, 2)):
            with t.query([[(ix, 1), (ix, 4)]], flags=cubit.Q_ASYNC, agg=cubit.AGG_SUM, agg_a=col_id) as r:
                r.add_limbs(acc.data_ptr() + 40 * slot)
                r.wait()
                tot[slot][0] += r.count
                tot[slot][1] += r.sum
    torch.cuda.synchronize()
    limbs = acc.cpu().tolist()
    for slot in range(2):
        c, l0, l1, l2, l3 = limbs[slot * 5:slot * 5 + 5]
        assert c == tot[slot][0]
        assert l0 + (l1 << 32) + (l2 << 64) + (l3 << 96) == tot[slot][1]
    t.close()


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2", reason="needs 2 GPUs")
def test_nccl_sharded_scan_and_rowid_gather():
    """tools/multi_gpu_check.py as a test: one rank per GPU over NCCL, row-range shards of one synthetic table,
    exact all-reduce of (COUNT, SUM), row-ID lists gathered to rank 0 with send/recv and checked there (ascending,
    length == COUNT, Σ ids == SUM(payload))"""
    import os
    import subprocess
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(root, "tools", "multi_gpu_check.py"),
           "50000017"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "multi_gpu_check ok: world=%d" % world in r.stdout, r.stdout[-2000:]
