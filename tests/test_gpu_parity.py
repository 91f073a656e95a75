"""GPU: the CUDA path, called through the C-ABI, against the oracle and the reference goldens.

Bit-exact everywhere (row IDs, values, integer aggregates); there is no float on this path.
"""
import hashlib

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

INDEX_BASE = {"quantity": 1, "discount": 0, "month": 0}
INDEX_CARD = {"quantity": 50, "discount": 11, "month": 84}
COL_PRICE, COL_DISC, COL_SHIP, COL_QTY = 0, 1, 2, 3


def digest(ids):
    return hashlib.sha256(np.ascontiguousarray(ids, dtype="<i8").tobytes()).hexdigest()


def make_lineitem_table(cubit, lineitem, seg_bits=65536, build_on_gpu=False):
    n = len(lineitem["price"])
    t = cubit.CubitTable(n, seg_bits=seg_bits)
    t.upload_column(COL_PRICE, lineitem["price"])
    t.upload_column(COL_DISC, lineitem["discount"])
    t.upload_column(COL_SHIP, lineitem["shipdate"])  # int32
    ix = {}
    for slot, c in enumerate(("quantity", "discount", "month")):
        if build_on_gpu:
            t.upload_column(10 + slot, lineitem[c])
            ix[c] = t.create_index(INDEX_CARD[c])
            t.build_index(ix[c], 10 + slot, INDEX_BASE[c])
        else:
            ix[c] = t.upload_index(oracle.build_index(lineitem[c], INDEX_BASE[c], INDEX_CARD[c]))
    return t, ix


def refs(ent, ix):
    return [[(ix[c], v - INDEX_BASE[c]) for (c, v) in grp] for grp in ent["groups"]]


@pytest.mark.parametrize("seg_bits", [32768, 65536, 131072])
@pytest.mark.parametrize("mode", ["default", "unfused", "fuse_probe"])
def test_tpch_sf001_against_reference_goldens(cubit, golden, lineitem, seg_bits, mode):
    """default = fused merge+decode kernel then bit-driven probe kernel; unfused = three kernels
    (merge, decode, gather probe); fuse_probe = everything in the scan kernel"""
    g, gids = golden
    unfused = mode == "unfused"
    t, ix = make_lineitem_table(cubit, lineitem, seg_bits, build_on_gpu=(seg_bits == 65536))
    extra = {"default": 0, "unfused": cubit.Q_UNFUSED, "fuse_probe": cubit.Q_FUSE_PROBE}[mode]
    for name, ent in g["tpch_sf001"]["answers"].items():
        want = gids["tpch_sf001/%s/ids" % name]
        with t.query(refs(ent, ix), flags=cubit.Q_ROWIDS | cubit.Q_VALUES | cubit.Q_BITVECTOR | extra,
                     cols=[COL_PRICE, COL_DISC], agg=cubit.AGG_SUM_PROD, agg_a=COL_PRICE, agg_b=COL_DISC) as r:
            assert r.count == ent["count"], name
            assert r.sum == ent["sum_price_x_discount_e4"], name
            ids, (price, disc) = r.fetch()
            assert np.array_equal(ids, want), name
            assert np.array_equal(price, lineitem["price"][want])
            assert np.array_equal(disc, lineitem["discount"][want])
            assert np.array_equal(oracle.decode(r.bitvector()), want)
            assert r.info.fused == (0 if unfused else 1)
        # SUM(price) with an int32 projected column → separate probe kernel
        with t.query(refs(ent, ix), flags=cubit.Q_ROWIDS | cubit.Q_VALUES | extra, cols=[COL_SHIP],
                     agg=cubit.AGG_SUM, agg_a=COL_PRICE) as r:
            assert r.count == ent["count"] and r.sum == ent["sum_price_cents"], name
            ids, (ship,) = r.fetch()
            assert np.array_equal(ids, want) and np.array_equal(ship, lineitem["shipdate"][want])
        # aggregate only (no row IDs materialised)
        with t.query(refs(ent, ix), flags=extra, agg=cubit.AGG_SUM, agg_a=COL_PRICE) as r:
            assert r.count == ent["count"] and r.sum == ent["sum_price_cents"], name
    t.close()


def test_q6_answer_file(cubit, golden, lineitem):
    """extension/tpch/dbgen/answers/sf0.01/q06.csv = 1193053.2253"""
    g, _ = golden
    t, ix = make_lineitem_table(cubit, lineitem)
    ent = g["tpch_sf001"]["answers"]["q6"]
    with t.query(refs(ent, ix), flags=0, agg=cubit.AGG_SUM_PROD, agg_a=COL_PRICE, agg_b=COL_DISC) as r:
        assert r.sum == 11930532253 and r.info.n_streams == 38
    t.close()


def delta_lists(gids, lineitem):
    upd = gids["tpch_sf001_delta/updated_rows"]
    newq = gids["tpch_sf001_delta/new_quantity"]
    dele = gids["tpch_sf001_delta/deleted_rows"]
    flips = {}
    for r, a, b in zip(upd.tolist(), lineitem["quantity"][upd].tolist(), newq.tolist()):
        flips.setdefault(("quantity", a), []).append(r)
        flips.setdefault(("quantity", b), []).append(r)
    for c in ("quantity", "discount", "month"):
        for r, v in zip(dele.tolist(), lineitem[c][dele].tolist()):
            flips.setdefault((c, v), []).append(r)
    return flips


@pytest.mark.parametrize("seg_bits", [32768, 131072])
def test_pending_deltas_match_sql_update_delete(cubit, golden, lineitem, seg_bits):
    g, gids = golden
    t, ix = make_lineitem_table(cubit, lineitem, seg_bits)
    for (c, v), rows in delta_lists(gids, lineitem).items():
        t.set_delta(ix[c], v - INDEX_BASE[c], np.asarray(rows))
    answers = g["tpch_sf001_delta"]["answers"]

    def check():
        for name, ent in answers.items():
            want = gids["tpch_sf001_delta/%s/ids" % name]
            for extra in (0, cubit.Q_UNFUSED, cubit.Q_FUSE_PROBE):
                with t.query(refs(ent, ix), flags=cubit.Q_ROWIDS | extra, agg=cubit.AGG_SUM, agg_a=COL_PRICE) as r:
                    ids, _ = r.fetch()
                    assert np.array_equal(ids, want), name
                    assert r.count == ent["count"] and r.sum == ent["sum_price_cents"]
    check()                      # deltas XOR-ed at query time
    some = t.query(refs(answers["q_10_19"], ix), flags=0)
    assert some.info.delta_entries > 0
    some.free()
    for c in ix:
        t.merge_deltas(ix[c])    # merge-back: B ^= D
    check()
    some = t.query(refs(answers["q_10_19"], ix), flags=0)
    assert some.info.delta_entries == 0
    some.free()
    t.close()


def random_case(rng, n, card):
    col = rng.integers(0, card, n).astype(np.int32)
    pay = rng.integers(-2**62, 2**62, n).astype(np.int64)
    return col, pay


@pytest.mark.parametrize("n,seg_bits,row_base", [
    (1, 32768, 0), (63, 65536, 64), (64, 65536, 0), (65, 131072, 128), (32768, 32768, 0), (32769, 32768, 32768 * 3),
    (65536 * 3 + 5, 65536, 65536), (1_000_003, 65536, 0), (5_000_011, 131072, 131072 * 7), (3_333_333, 32768, 0)])
def test_random_tables_all_paths(cubit, n, seg_bits, row_base):
    rng = np.random.default_rng(n)
    card = 16
    col, pay = random_case(rng, n, card)
    bv = oracle.build_index(col, 0, card)
    t = cubit.CubitTable(n, row_base=row_base, seg_bits=seg_bits)
    small = pay >> 40
    t.upload_column(0, pay)
    t.upload_column(1, col)
    t.upload_column(2, small)
    ix = t.create_index(card)
    t.build_index(ix, 1, 0)
    for v in range(card):
        assert np.array_equal(t.download_bitvector(ix, v), bv[v])
        assert t.bitvector_count(ix, v) == oracle.popcount(bv[v])
    # pending deltas on two bitvectors (with a duplicate row that must cancel)
    d3 = rng.integers(0, n, max(1, n // 50))
    d7 = np.concatenate([rng.integers(0, n, max(1, n // 80)), d3[:1], d3[:1]])
    t.set_delta(ix, 3, d3)
    t.set_delta(ix, 7, d7)
    dl = [None] * card
    dl[3] = oracle.delta_from_rows(d3, n)
    dl[7] = oracle.delta_from_rows(d7, n)
    cases = [
        [[0]], [[3]], [list(range(card))], [[1, 3, 5, 7]], [[1, 3, 5, 7], [3, 7, 9]], [[2], [2]], [[2], [4]],
        [[0, 1, 2, 3], [3, 4, 5], [3, 7]],
    ]
    for groups in cases:
        og = [[bv[v] for v in grp] for grp in groups]
        od = [[dl[v] for v in grp] for grp in groups]
        q = oracle.merge(og, od)
        want = oracle.decode(q, row_base)
        wv = oracle.probe(want, pay, row_base)
        for extra in (0, cubit.Q_UNFUSED, cubit.Q_FUSE_PROBE):
            with t.query([[(ix, v) for v in grp] for grp in groups],
                         flags=cubit.Q_ROWIDS | cubit.Q_VALUES | cubit.Q_BITVECTOR | extra, cols=[0, 1],
                         agg=cubit.AGG_SUM, agg_a=0) as r:
                assert r.count == len(want), (groups, extra)
                ids, (vals, cv) = r.fetch()
                assert np.array_equal(ids, want)
                assert np.array_equal(vals, wv)
                assert np.array_equal(cv, col[want - row_base])
                assert r.sum == oracle.sum_i64(wv)
                assert np.array_equal(r.bitvector(), q)
                # int64-only projection → the bit-driven probe (default) / in-kernel probe paths
                with t.query([[(ix, v) for v in grp] for grp in groups], flags=cubit.Q_VALUES | extra, cols=[0],
                             agg=cubit.AGG_SUM_PROD, agg_a=2, agg_b=2) as r2:
                    ws = oracle.probe(want, small, row_base)
                    sp, ovf = oracle.sum_prod_i64(ws, ws)
                    assert not ovf and r2.count == len(want) and r2.sum == sp
                    assert np.array_equal(r2.fetch(rowids=False)[1][0], wv)
                if len(want) > 5:  # windowed fetch, the DataChunk hand-off (≤2048 rows per call)
                    i2, (v2, _) = r.fetch(offset=3, n=min(2048, len(want) - 3))
                    assert np.array_equal(i2, want[3:3 + len(i2)]) and np.array_equal(v2, wv[3:3 + len(i2)])
    # stand-alone probe (DataTable::Fetch analog)
    some = np.sort(rng.choice(n, size=min(n, 1001), replace=False)) + row_base
    vals, s = t.probe(0, some, want_sum=True)
    assert np.array_equal(vals, pay[some - row_base]) and s == oracle.sum_i64(vals)
    cvals, _ = t.probe(1, some)
    assert np.array_equal(cvals, col[some - row_base])
    t.close()


def test_edge_bitvectors(cubit):
    n = 200_003
    t = cubit.CubitTable(n)
    zeros = np.zeros(t.n_words, dtype=np.uint64)
    ones = np.full(t.n_words, ~np.uint64(0), dtype=np.uint64)
    ones[-1] = np.uint64((1 << (n % 64)) - 1)
    ix = t.upload_index(np.stack([zeros, ones]))
    t.synth_column(0, 0)
    with t.query([[(ix, 0)]], flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=0) as r:
        assert r.count == 0 and r.sum == 0 and len(r.fetch()[0]) == 0
    with t.query([[(ix, 1)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0], agg=cubit.AGG_SUM, agg_a=0) as r:
        ids, (v,) = r.fetch()
        assert np.array_equal(ids, np.arange(n)) and np.array_equal(v, ids) and r.sum == n * (n - 1) // 2
    with t.query([[(ix, 1)], [(ix, 0)]], flags=cubit.Q_ROWIDS) as r:
        assert r.count == 0
    with pytest.raises(cubit.CubitError):   # bits beyond n_rows are rejected
        bad = ones.copy()
        bad[-1] = ~np.uint64(0)
        t.upload_bitvector(ix, 0, bad)
    with pytest.raises(cubit.CubitError):
        t.query([[(ix, 5)]])
    with pytest.raises(cubit.CubitError):
        t.query([])
    with pytest.raises(cubit.CubitError):
        t.set_delta(ix, 0, np.array([n]))
    t.close()


def test_sum_prod_overflow_is_an_error(cubit):
    n = 1000
    t = cubit.CubitTable(n)
    t.upload_column(0, np.full(n, 2**40, dtype=np.int64))
    ones = np.zeros(t.n_words, dtype=np.uint64)
    ones[0] = 1
    ix = t.upload_index(ones[None, :])
    with pytest.raises(cubit.CubitError) as e:
        t.query([[(ix, 0)]], flags=0, agg=cubit.AGG_SUM_PROD, agg_a=0, agg_b=0)
    assert "Overflow" in str(e.value)
    t.close()


def test_synth_generator_and_gpu_index_build_match_reference_sql(cubit, golden):
    g, _ = golden
    for tag, ent in g["synthetic"].items():
        n = ent["n_rows"]
        t = cubit.CubitTable(n)
        t.synth_column(1, 1, seed=ent["seed"], threshold=int(ent["threshold"]), card=ent["card"], hot_lo=10, hot_n=10)
        t.synth_column(0, 0)
        assert np.array_equal(t.download_column(1),
                              oracle.synth_column(1, n, 0, ent["seed"], int(ent["threshold"]), ent["card"], 10, 10))
        ix = t.create_index(ent["card"])
        t.build_index(ix, 1, 0)
        with t.query([[(ix, v) for v in range(10, 20)]], flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=0) as r:
            ids, _ = r.fetch()
            assert r.count == ent["count"] and r.sum == ent["sum_rowid"] and digest(ids) == ent["ids_sha256"]
        t.close()


def test_sf1_shape_config1(cubit):
    """config 1 shape: 6,001,215 rows, 50-value index, equality predicate + SUM (data synthetic:
    the SF1 table itself cannot travel; its reference answers are pinned on the CPU side)"""
    n = 6_001_215
    rng = np.random.default_rng(1)
    qty = rng.integers(1, 51, n).astype(np.int32)
    price = rng.integers(90_000, 10_500_000, n).astype(np.int64)
    t = cubit.CubitTable(n)
    t.upload_column(0, price)
    t.upload_column(1, qty)
    ix = t.create_index(50)
    t.build_index(ix, 1, 1)
    want = np.flatnonzero(qty == 24)
    with t.query([[(ix, 23)]], flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=0) as r:
        ids, _ = r.fetch()
        assert np.array_equal(ids, want) and r.sum == int(price[want].sum())
    t.close()


@pytest.mark.parametrize("sel", ["1e-4", "1e-3", "1e-2", "0.1", "0.25", "0.5"])
def test_full_size_properties_1b_rows(cubit, sel):
    """BASELINE config 2 at full size (10^9 rows): size-independent properties.
    payload = row id, so SUM(payload) == SUM(row ids); the OR of disjoint value bitvectors
    must select exactly Σ popcount(B_v) rows; row IDs strictly ascending; every returned
    row's value lies in the predicate range."""
    from fractions import Fraction
    n = 1_000_000_000
    thr = int(Fraction(sel) * (1 << 64))
    t = cubit.CubitTable(n)
    t.synth_column(1, 1, seed=0xC0B17, threshold=thr, card=100, hot_lo=10, hot_n=10)
    ix = t.create_index(100)
    t.build_index(ix, 1, 0)
    t.synth_column(0, 0)
    total = sum(t.bitvector_count(ix, v) for v in range(100))
    assert total == n
    expect = sum(t.bitvector_count(ix, v) for v in range(10, 20))
    assert abs(expect / n - float(Fraction(sel))) < 0.01 * float(Fraction(sel)) + 1e-6
    with t.query([[(ix, v) for v in range(10, 20)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[1],
                 agg=cubit.AGG_SUM, agg_a=0) as r:
        assert r.count == expect
        fused_sum = r.sum
        chunk = 1 << 25
        prev, acc = -1, 0
        for off in range(0, r.count, chunk):
            ids, (vals,) = r.fetch(off, min(chunk, r.count - off))
            assert ids[0] > prev and (np.diff(ids) > 0).all()
            assert vals.min() >= 10 and vals.max() <= 19
            prev = int(ids[-1])
            acc += int(ids.sum(dtype=np.uint64))
        assert prev < n and acc == fused_sum
    # the three-kernel path agrees with the fused one
    with t.query([[(ix, v) for v in range(10, 20)]], flags=cubit.Q_ROWIDS | cubit.Q_UNFUSED, agg=cubit.AGG_SUM,
                 agg_a=0) as r2:
        assert r2.count == expect and r2.sum == fused_sum
    t.close()


def test_q6_shape_three_indexes_k38_against_oracle(cubit):
    """config 3 shape at a size the oracle finishes in seconds: three indexes (84 month bins, 11 discounts,
    50 quantities), Q = (12 months) AND (3 discounts) AND (23 quantities), k = 38, probe price & discount,
    SUM(price*discount); device generators vs their oracle restatement included."""
    n = 3_000_017
    t = cubit.CubitTable(n, row_base=131072, seg_bits=131072)
    t.synth_column(0, 3, seed=21, threshold=10_410_000, hot_lo=90_000)
    t.synth_column(1, 3, seed=22, threshold=11, hot_lo=0)
    t.synth_column(2, 2, seed=23, card=50, hot_lo=1)
    t.synth_column(3, 2, seed=24, card=84, hot_lo=0)
    price = oracle.synth_column(3, n, 131072, 21, 10_410_000, 0, 90_000, 0)
    disc = oracle.synth_column(3, n, 131072, 22, 11, 0, 0, 0)
    qty = oracle.synth_column(2, n, 131072, 23, 0, 50, 1, 0)
    month = oracle.synth_column(2, n, 131072, 24, 0, 84, 0, 0)
    for cid, ref in ((0, price), (1, disc), (2, qty), (3, month)):
        assert np.array_equal(t.download_column(cid), ref)
    ixd, ixq, ixm = t.create_index(11), t.create_index(50), t.create_index(84)
    t.build_index(ixd, 1, 0)
    t.build_index(ixq, 2, 1)
    t.build_index(ixm, 3, 0)
    bd, bq, bm = oracle.build_index(disc, 0, 11), oracle.build_index(qty, 1, 50), oracle.build_index(month, 0, 84)
    groups = [[(ixm, m) for m in range(24, 36)], [(ixd, v) for v in (5, 6, 7)], [(ixq, v) for v in range(23)]]
    og = [[bm[m] for m in range(24, 36)], [bd[v] for v in (5, 6, 7)], [bq[v] for v in range(23)]]
    want = oracle.decode(oracle.merge(og), 131072)
    wp, wd = oracle.probe(want, price, 131072), oracle.probe(want, disc, 131072)
    ws, ovf = oracle.sum_prod_i64(wp, wd)
    assert not ovf and len(want) > 1000
    for extra in (0, cubit.Q_UNFUSED, cubit.Q_FUSE_PROBE):
        with t.query(groups, flags=cubit.Q_ROWIDS | cubit.Q_VALUES | extra, cols=[0, 1], agg=cubit.AGG_SUM_PROD,
                     agg_a=0, agg_b=1) as r:
            assert r.info.n_streams == 38 and r.count == len(want) and r.sum == ws
            ids, (p, d) = r.fetch()
            assert np.array_equal(ids, want) and np.array_equal(p, wp) and np.array_equal(d, wd)
        with t.query(groups, flags=extra, agg=cubit.AGG_SUM_PROD, agg_a=0, agg_b=1) as r:
            assert r.count == len(want) and r.sum == ws
    t.close()


def test_concurrent_queries_on_one_table(cubit):
    """the table handle is thread-safe (calls serialise on its stream): 8 host threads, mixed predicates"""
    import threading
    n = 2_000_003
    rng = np.random.default_rng(5)
    col = rng.integers(0, 20, n).astype(np.int32)
    pay = rng.integers(-10**9, 10**9, n).astype(np.int64)
    bv = oracle.build_index(col, 0, 20)
    t = cubit.CubitTable(n)
    t.upload_column(0, pay)
    t.upload_column(1, col)
    ix = t.create_index(20)
    t.build_index(ix, 1, 0)
    errors = []

    def work(k):
        try:
            for rep in range(6):
                vals = [(k + rep + j) % 20 for j in range(1 + (k % 4))]
                want = oracle.decode(oracle.merge([[bv[v] for v in vals]]))
                with t.query([[(ix, v) for v in vals]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0],
                             agg=cubit.AGG_SUM, agg_a=0) as r:
                    ids, (v,) = r.fetch()
                    assert np.array_equal(ids, want) and np.array_equal(v, pay[want]) and r.sum == int(pay[want].sum())
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    th = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errors, errors
    t.close()


def test_maximum_stream_count(cubit):
    """CUBIT_MAX_STREAMS = 64 bitvectors in one query (4 groups of 16); 65 is rejected"""
    n = 700_001
    rng = np.random.default_rng(9)
    col = rng.integers(0, 80, n).astype(np.int32)
    bv = oracle.build_index(col, 0, 80)
    t = cubit.CubitTable(n, seg_bits=32768)
    t.upload_column(1, col)
    ix = t.create_index(80)
    t.build_index(ix, 1, 0)
    d = rng.integers(0, n, 3000)
    t.set_delta(ix, 70, d)          # stream 62: exercises the second half of the delta offset table
    groups = [list(range(g * 16, g * 16 + 16)) for g in range(4)]
    groups[1] = list(range(0, 16))  # overlap so the AND is not empty: (0..15) AND (0..15) AND ...
    groups[2] = list(range(8, 24))
    groups[3] = [v for v in range(8, 16)] + [70, 71, 72, 73, 74, 75, 76, 77]
    groups[3], groups[0] = groups[0], groups[3]
    dl = [[oracle.delta_from_rows(d, n) if v == 70 else None for v in g] for g in groups]
    want = oracle.decode(oracle.merge([[bv[v] for v in g] for g in groups], dl))
    assert len(want) > 0
    with t.query([[(ix, v) for v in g] for g in groups], flags=cubit.Q_ROWIDS) as r:
        assert r.info.n_streams == 64 and np.array_equal(r.fetch()[0], want)
    with pytest.raises(cubit.CubitError):
        t.query([[(ix, v) for v in range(65)]])
    t.close()


@pytest.mark.parametrize("extra_name", ["default", "unfused", "fuse_probe"])
def test_double_sum_within_1e12(cubit, extra_name):
    """SUM over a DOUBLE column (north_star: float aggregates within 1e-12 relative of the reference order sum)"""
    import math
    extra = {"default": 0, "unfused": cubit.Q_UNFUSED, "fuse_probe": cubit.Q_FUSE_PROBE}[extra_name]
    n = 4_000_037
    rng = np.random.default_rng(17)
    col = rng.integers(0, 10, n).astype(np.int32)
    val = rng.normal(0.0, 1e6, n) + 1e9          # float64, large common offset: cancellation-free but long sum
    bv = oracle.build_index(col, 0, 10)
    t = cubit.CubitTable(n)
    t.upload_column(0, val)
    t.upload_column(1, col)
    ix = t.create_index(10)
    t.build_index(ix, 1, 0)
    TOL = 1e-12  # relative tolerance stated by north_star
    for vals_sel in ([3], [0, 1, 2, 3, 4, 5, 6], list(range(10))):
        want = oracle.decode(oracle.merge([[bv[v] for v in vals_sel]]))
        ref_plain, ref_comp = oracle.sum_f64(val[want])      # DuckDB order (plain) and compensated
        exact = math.fsum(val[want].tolist())
        for flags in (extra, cubit.Q_ROWIDS | cubit.Q_VALUES | extra):
            with t.query([[(ix, v) for v in vals_sel]], flags=flags, cols=[0] if flags & cubit.Q_VALUES else (),
                         agg=cubit.AGG_SUM_F64, agg_a=0) as r:
                assert r.count == len(want)
                assert abs(r.sum_f64 - ref_plain) <= TOL * abs(ref_plain)
                assert abs(r.sum_f64 - exact) <= TOL * abs(exact)
                if flags & cubit.Q_VALUES:
                    ids, (got,) = r.fetch()
                    assert np.array_equal(ids, want) and np.array_equal(got, val[want])   # values are bit-exact
    t.close()


def test_scan_while_deltas_change(cubit):
    """concurrentloop-style stress (test/sql/parallelism/interquery/*): one thread keeps replacing /
    clearing / merging pending deltas while others scan; every scan must see exactly one of the consistent states"""
    import threading
    n = 1_500_007
    rng = np.random.default_rng(23)
    col = rng.integers(0, 8, n).astype(np.int32)
    bv = oracle.build_index(col, 0, 8)
    t = cubit.CubitTable(n, seg_bits=32768)
    t.upload_column(1, col)
    ix = t.create_index(8)
    t.build_index(ix, 1, 0)
    d = np.unique(rng.integers(0, n, 20_000))
    base = oracle.decode(oracle.merge([[bv[2], bv[5]]]))
    flipped = oracle.decode(oracle.merge([[bv[2], bv[5]]], [[oracle.delta_from_rows(d, n), None]]))
    assert len(base) != len(flipped)
    states = {digest(base), digest(flipped)}
    stop = threading.Event()
    errors = []

    def writer():
        try:
            i = 0
            while not stop.is_set():
                t.set_delta(ix, 2, d if i % 2 == 0 else np.zeros(0, dtype=np.int64))
                i += 1
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def reader():
        try:
            for _ in range(40):
                with t.query([[(ix, 2), (ix, 5)]], flags=cubit.Q_ROWIDS) as r:
                    ids, _ = r.fetch()
                    if digest(ids) not in states:
                        errors.append("scan saw a state that never existed (%d rows)" % len(ids))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    w = threading.Thread(target=writer)
    rs = [threading.Thread(target=reader) for _ in range(3)]
    w.start()
    [x.start() for x in rs]
    [x.join() for x in rs]
    stop.set()
    w.join()
    assert not errors, errors[:3]
    t.close()


def test_high_cardinality_index_build(cubit):
    """cardinality above one shared-memory tile (400 values per pass) → multi-pass GPU index build;
    values outside the indexed domain are ignored (NULL-like)"""
    n, card = 300_011, 1000
    rng = np.random.default_rng(31)
    col = rng.integers(-50, card + 50, n).astype(np.int64)
    bv = oracle.build_index(col, 0, card)
    t = cubit.CubitTable(n, seg_bits=32768)
    t.upload_column(0, col)
    ix = t.create_index(card)
    t.build_index(ix, 0, 0)
    for v in (0, 1, 399, 400, 401, 799, 800, 999):
        assert np.array_equal(t.download_bitvector(ix, v), bv[v]), v
    assert sum(t.bitvector_count(ix, v) for v in range(card)) == int(((col >= 0) & (col < card)).sum())
    want = oracle.decode(oracle.merge([[bv[v] for v in range(380, 420)]]))
    with t.query([[(ix, v) for v in range(380, 420)]], flags=cubit.Q_ROWIDS) as r:
        assert np.array_equal(r.fetch()[0], want)
    t.close()


def test_zero_copy_rowid_tensor(cubit):
    """sharding.result_rowids_tensor: the device-resident row-ID list as a torch tensor without a copy
    (what the multi-GPU result gather sends over NCCL)"""
    import importlib
    import torch
    sharding = importlib.import_module("duckdb-cubit_b200.sharding")
    n = 500_003
    rng = np.random.default_rng(41)
    col = rng.integers(0, 6, n).astype(np.int32)
    t = cubit.CubitTable(n, row_base=65536)
    t.upload_column(1, col)
    ix = t.create_index(6)
    t.build_index(ix, 1, 0)
    with t.query([[(ix, 2), (ix, 4)]], flags=cubit.Q_ROWIDS) as r:
        ids, _ = r.fetch()
        ten = sharding.result_rowids_tensor(r, torch.device("cuda", 0))
        assert ten.data_ptr() == r.info.d_rowids and ten.numel() == r.count
        assert np.array_equal(ten.cpu().numpy(), ids)
        assert sharding.gather_sorted(ten, None) is ten
    t.close()


@pytest.mark.parametrize("keep_raw,seg_bits", [(False, 65536), (True, 65536), (False, 131072), (False, 32768)])
def test_bit_packed_columns_are_lossless_on_every_probe_path(cubit, keep_raw, seg_bits):
    """cubit_gpu_pack_column: FOR-bit-packed int64 columns decode to exactly the raw values on every probe
    path (bit-driven, gather over row IDs, in-scan fused, stand-alone probe) for every width 0..64"""
    n = 1_200_011
    rng = np.random.default_rng(77)
    col = rng.integers(0, 12, n).astype(np.int32)
    # blocks of 1024 rows get different ranges: constant, 1 bit, ..., 63 bits, full 64-bit range, negatives
    blk = np.arange(n) // 1024
    width = (blk % 66).astype(np.int64)                       # 0..65 (≥64 → full range)
    lo = rng.integers(-2**40, 2**40, n // 1024 + 1)[blk]
    span = np.where(width >= 63, 2**62, (1 << np.minimum(width, 62)))
    a = lo + (rng.random(n) * span).astype(np.int64)
    full = width >= 64
    a[full] = rng.integers(-2**63, 2**63 - 1, int(full.sum()), dtype=np.int64)
    b = rng.integers(0, 1000, n).astype(np.int64)             # small second column for SUM_PROD
    b[: 5 * 1024] = 7                                         # constant blocks (width 0)
    bv = oracle.build_index(col, 0, 12)
    t = cubit.CubitTable(n, row_base=65536 * 2, seg_bits=seg_bits)
    t.upload_column(0, a)
    t.upload_column(1, col)
    t.upload_column(2, b)
    ix = t.create_index(12)
    t.build_index(ix, 1, 0)
    pa, pb = t.pack_column(0, keep_raw=keep_raw), t.pack_column(2, keep_raw=keep_raw)
    assert pa < n * 8 and pb < n * 2                          # b needs ≤ 10 bits per value
    assert np.array_equal(t.download_column(0), a) and np.array_equal(t.download_column(2), b)
    for vals in ([3], [0, 1, 2, 3, 4, 5, 6, 7], list(range(12))):       # sparse-ish … all rows
        want = oracle.decode(oracle.merge([[bv[v] for v in vals]]), 65536 * 2)
        wa, wb = oracle.probe(want, a, 65536 * 2), oracle.probe(want, b, 65536 * 2)
        for extra in (0, cubit.Q_UNFUSED, cubit.Q_FUSE_PROBE):
            with t.query([[(ix, v) for v in vals]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES | extra, cols=[0, 2],
                         agg=cubit.AGG_SUM, agg_a=0) as r:
                ids, (ga, gb) = r.fetch()
                assert np.array_equal(ids, want) and np.array_equal(ga, wa) and np.array_equal(gb, wb)
                assert r.sum == oracle.sum_i64(wa)
            with t.query([[(ix, v) for v in vals]], flags=extra, agg=cubit.AGG_SUM_PROD, agg_a=2, agg_b=2) as r:
                sp, ovf = oracle.sum_prod_i64(wb, wb)
                assert not ovf and r.sum == sp
    some = np.sort(rng.choice(n, 5001, replace=False)) + 65536 * 2
    got, s = t.probe(0, some, want_sum=True)
    assert np.array_equal(got, a[some - 65536 * 2]) and s == oracle.sum_i64(got)
    with pytest.raises(cubit.CubitError):
        t.pack_column(1)                                      # 4-byte columns are not packed
    if not keep_raw:
        with pytest.raises(cubit.CubitError):
            ix2 = t.create_index(4)
            t.build_index(ix2, 2, 0)                          # raw form gone: build first, then pack
    t.close()


@pytest.mark.parametrize("seg_bits", [32768, 65536])
def test_append_rows_extends_columns_and_indexes(cubit, seg_bits):
    """cubit_gpu_append_rows (INSERT): appended rows take the next row ids, built indexes are extended on the
    GPU from the new rows only, uploaded bitvectors get zeros, pending deltas survive, and every query answer
    equals the oracle's on the full arrays — across word, build-tile, segment and capacity boundaries"""
    rng = np.random.default_rng(99)
    base = seg_bits * 3
    sizes = [50_001, 13, 4096 - 13, 70_000, 1, 300_000, 64]   # appended one after another
    total = sum(sizes)
    key = rng.integers(0, 9, total).astype(np.int32)
    pay = rng.integers(-2**50, 2**50, total).astype(np.int64)
    n = sizes[0]
    t = cubit.CubitTable(n, row_base=base, seg_bits=seg_bits)
    t.upload_column(0, pay[:n])
    t.upload_column(1, key[:n])
    ix = t.create_index(9)
    t.build_index(ix, 1, 0)
    up = t.create_index(2)                                       # an UPLOADED index (no source column)
    manual = (np.arange(n) % 3 == 0)
    t.upload_bitvector(up, 1, np.packbits(np.pad(manual, (0, -n % 64)), bitorder="little").view(np.uint64))
    drows = np.array([5, 77, 40_000])
    t.set_delta(ix, 4, drows)                                    # pending delete/update flips stay pending
    for add in sizes[1:]:
        t.append_rows({0: pay[n:n + add], 1: key[n:n + add]})
        n += add
        assert t.n_rows == n
        assert np.array_equal(t.download_column(0), pay[:n]) and np.array_equal(t.download_column(1), key[:n])
        bv = oracle.build_index(key[:n], 0, 9)
        for v in (0, 4, 8):
            assert np.array_equal(t.download_bitvector(ix, v), bv[v]), (n, v)
        d4 = oracle.delta_from_rows(drows, n)
        want = oracle.decode(oracle.merge([[bv[3], bv[4], bv[5]]], [[None, d4, None]]), base)
        with t.query([[(ix, 3), (ix, 4), (ix, 5)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0],
                     agg=cubit.AGG_SUM, agg_a=0) as r:
            ids, (vals,) = r.fetch()
            assert np.array_equal(ids, want) and np.array_equal(vals, oracle.probe(want, pay[:n], base))
            assert r.sum == oracle.sum_i64(vals)
        m = np.zeros(n, dtype=bool)
        m[:sizes[0]] = manual
        with t.query([[(up, 1)]], flags=cubit.Q_ROWIDS) as r:   # new rows of an uploaded bitvector are 0
            assert np.array_equal(r.fetch()[0], np.nonzero(m)[0] + base)
    t.merge_deltas(ix)
    bv = oracle.build_index(key, 0, 9)
    assert np.array_equal(t.download_bitvector(ix, 4), bv[4] ^ oracle.delta_from_rows(drows, total))
    with pytest.raises(cubit.CubitError):
        t.append_rows({0: pay[:10]})                             # every resident column must be supplied
    with pytest.raises(cubit.CubitError):
        t.append_rows({0: pay[:10], 1: key[:10].astype(np.int64)})  # wrong width
    t.close()


def test_single_bitvector_aggregate_runs_on_the_bitvector_itself(cubit, golden, lineitem):
    """equality predicate + SUM (config 1's query), no deltas: one launch (the bit-driven probe on B_v), and the
    same answers as the general path; a pending delta on that bitvector brings the scan kernel back"""
    g, gids = golden
    t, ix = make_lineitem_table(cubit, lineitem, 32768)
    ent = g["tpch_sf001"]["answers"]["q_eq_24"]
    with t.query(refs(ent, ix), flags=0, agg=cubit.AGG_SUM, agg_a=COL_PRICE) as r:
        assert r.count == ent["count"] and r.sum == ent["sum_price_cents"] and r.info.n_launches == 1
    with t.query(refs(ent, ix), flags=0, agg=cubit.AGG_SUM_PROD, agg_a=COL_PRICE, agg_b=COL_DISC) as r:
        assert r.count == ent["count"] and r.sum == ent["sum_price_x_discount_e4"] and r.info.n_launches == 1
    with t.query(refs(ent, ix), flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=COL_PRICE) as r:   # row IDs wanted: general path
        assert r.count == ent["count"] and r.sum == ent["sum_price_cents"] and r.info.n_launches == 2
        assert np.array_equal(r.fetch()[0], gids["tpch_sf001/q_eq_24/ids"])
    # delete one selected row through a pending delta: XOR at query time needs the merge kernel again
    victim = int(gids["tpch_sf001/q_eq_24/ids"][5])
    t.set_delta(ix["quantity"], 24 - INDEX_BASE["quantity"], np.array([victim], dtype=np.int64))
    with t.query(refs(ent, ix), flags=0, agg=cubit.AGG_SUM, agg_a=COL_PRICE) as r:
        assert r.count == ent["count"] - 1 and r.sum == ent["sum_price_cents"] - int(lineitem["price"][victim])
        assert r.info.n_launches == 2
    t.merge_deltas(ix["quantity"])
    with t.query(refs(ent, ix), flags=0, agg=cubit.AGG_SUM, agg_a=COL_PRICE) as r:
        assert r.count == ent["count"] - 1 and r.info.n_launches == 1
    t.close()
