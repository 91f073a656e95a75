"""CPU: pin the oracle against what the reference DuckDB produced (tests/golden/)."""
import hashlib

import numpy as np
import pytest

import oracle

INDEX_BASE = {"quantity": 1, "discount": 0, "month": 0}
INDEX_CARD = {"quantity": 50, "discount": 11, "month": 84}


def digest(ids):
    return hashlib.sha256(np.ascontiguousarray(ids, dtype="<i8").tobytes()).hexdigest()


def build_indexes(cols):
    return {c: oracle.build_index(cols[c], INDEX_BASE[c], INDEX_CARD[c]) for c in INDEX_BASE}


def groups_of(ent, idx):
    return [[idx[c][v - INDEX_BASE[c]] for (c, v) in grp] for grp in ent["groups"]]


def test_tpch_sf001_answers(golden, lineitem):
    g, gids = golden
    idx = build_indexes(lineitem)
    for name, ent in g["tpch_sf001"]["answers"].items():
        ids = oracle.decode(oracle.merge(groups_of(ent, idx)))
        assert np.array_equal(ids, gids["tpch_sf001/%s/ids" % name]), name
        assert len(ids) == ent["count"] and digest(ids) == ent["ids_sha256"]
        price = oracle.probe(ids, lineitem["price"])
        assert oracle.sum_i64(price) == ent["sum_price_cents"]
        sp, ovf = oracle.sum_prod_i64(price, oracle.probe(ids, lineitem["discount"]))
        assert not ovf and sp == ent["sum_price_x_discount_e4"]


def test_q6_matches_answer_file(golden):
    g, _ = golden
    # extension/tpch/dbgen/answers/sf0.01/q06.csv → 1193053.2253 ; sf1 → 123141078.2283
    assert g["tpch_sf001"]["answers"]["q6"]["sum_price_x_discount_e4"] == 11930532253
    assert g["tpch_sf1"]["answers"]["q6"]["sum_price_x_discount_e4"] == 1231410782283
    assert g["tpch_sf1"]["answers"]["q_eq_24"]["count"] == 119971
    assert g["tpch_sf1"]["answers"]["q_eq_24"]["sum_price_cents"] == 431592967032
    assert g["tpch_sf1"]["oracle_reproduced_here"] and g["tpch_sf01"]["oracle_reproduced_here"]


def delta_lists(gids, lineitem):
    """UPDATE v→w of row r flips r in D_v and D_w; DELETE of a row with value v flips r in D_v"""
    upd = gids["tpch_sf001_delta/updated_rows"]
    newq = gids["tpch_sf001_delta/new_quantity"]
    dele = gids["tpch_sf001_delta/deleted_rows"]
    flips = {}  # (column, value) -> rows
    oldq = lineitem["quantity"][upd]
    for r, a, b in zip(upd.tolist(), oldq.tolist(), newq.tolist()):
        flips.setdefault(("quantity", a), []).append(r)
        flips.setdefault(("quantity", b), []).append(r)
    for c in ("quantity", "discount", "month"):
        for r, v in zip(dele.tolist(), lineitem[c][dele].tolist()):
            flips.setdefault((c, v), []).append(r)
    return flips


def test_delta_semantics_match_sql_update_delete(golden, lineitem):
    g, gids = golden
    idx = build_indexes(lineitem)
    n = len(lineitem["price"])
    flips = delta_lists(gids, lineitem)
    dl = {c: [None] * INDEX_CARD[c] for c in INDEX_BASE}
    for (c, v), rows in flips.items():
        dl[c][v - INDEX_BASE[c]] = oracle.delta_from_rows(np.asarray(rows), n)
    for name, ent in g["tpch_sf001_delta"]["answers"].items():
        grp = groups_of(ent, idx)
        dgr = [[dl[c][v - INDEX_BASE[c]] for (c, v) in gg] for gg in ent["groups"]]
        ids = oracle.decode(oracle.merge(grp, dgr))
        assert np.array_equal(ids, gids["tpch_sf001_delta/%s/ids" % name]), name
        assert oracle.sum_i64(oracle.probe(ids, lineitem["price"])) == ent["sum_price_cents"]
        # numpy restatement agrees
        assert np.array_equal(oracle.np_decode(oracle.np_merge(grp, dgr)), ids)


def test_synthetic_generator_against_reference_sql(golden):
    g, _ = golden
    for tag, ent in g["synthetic"].items():
        v = oracle.synth_column(1, ent["n_rows"], 0, ent["seed"], int(ent["threshold"]), ent["card"], ent["hot_lo"],
                                ent["hot_n"])
        hist = np.bincount(v, minlength=ent["card"])
        assert [[i, int(c)] for i, c in enumerate(hist) if c] == ent["histogram"]
        bv = oracle.build_index(v, 0, ent["card"])
        ids = oracle.decode(oracle.merge([[bv[i] for i in range(10, 20)]]))
        assert len(ids) == ent["count"] and int(ids.sum()) == ent["sum_rowid"] and digest(ids) == ent["ids_sha256"]
        # bitvectors built without the column agree
        direct = oracle.synth_bitvectors(ent["n_rows"], 0, ent["seed"], int(ent["threshold"]), ent["card"], 10, 10,
                                         10, 10, n_threads=3)
        assert np.array_equal(direct, bv[10:20])


def test_art_many_matches_known_answers():
    """test/sql/index/art/scan/test_art_many_matches.test: [0,1,0,1,...] x 1024 and x 2048"""
    for reps in (1024, 2048):
        col = np.tile(np.array([0, 1], dtype=np.int32), reps)
        bv = oracle.build_index(col, 0, 2)
        cnt = lambda vals: len(oracle.decode(oracle.merge([[bv[v] for v in vals]])))  # noqa: E731
        assert cnt([0]) == reps            # i<1, i=0
        assert cnt([0, 1]) == 2 * reps     # i<=1, i>=0
        assert cnt([1]) == reps            # i=1, i>0
        ids = oracle.decode(bv[1])
        assert np.array_equal(ids, np.arange(1, 2 * reps, 2))  # sorted, unique (art.cpp:974-985)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 1000, 4096, 65536 + 7, 200_003])
def test_c_vs_numpy_random(n):
    rng = np.random.default_rng(n)
    col = rng.integers(0, 12, n).astype(np.int64)
    bv = oracle.build_index(col, 0, 12)
    assert np.array_equal(bv, oracle.np_build_index(col, 0, 12))
    groups = [[bv[0], bv[3], bv[5]], [bv[3], bv[5], bv[7]]]
    deltas = [[oracle.delta_from_rows(rng.integers(0, n, 5), n), None, None],
              [None, oracle.delta_from_rows(rng.integers(0, n, 9), n), None]]
    q = oracle.merge(groups, deltas)
    assert np.array_equal(q, oracle.np_merge(groups, deltas))
    for seg in (1, 7, 512, 1024):  # the result does not depend on the segment size
        assert np.array_equal(oracle.merge(groups, deltas, seg_words=seg), q)
    ids = oracle.decode(q, row_base=128)
    assert np.array_equal(ids, oracle.np_decode(q, 128))
    assert (np.diff(ids) > 0).all()
    pay = rng.integers(-2**62, 2**62, n).astype(np.int64)
    vals = oracle.probe(ids, pay, row_base=128)
    assert np.array_equal(vals, pay[ids - 128])
    assert oracle.sum_i64(vals) == oracle.np_sum(vals)
    cnt, i2, v2, s2 = oracle.scan_mt(groups, payload=pay, row_base=128, n_threads=3, deltas=deltas)
    assert cnt == len(ids) and np.array_equal(i2[:cnt], ids) and np.array_equal(v2[:cnt], vals)
    assert s2 == oracle.np_sum(vals)


def test_sum_128bit_carry_and_overflow_flag():
    big = np.array([2**63 - 1] * 5 + [-2**63] * 3 + [-1, 1, 7], dtype=np.int64)
    assert oracle.sum_i64(big) == sum(int(x) for x in big)
    s, ovf = oracle.sum_prod_i64(np.array([2**40], dtype=np.int64), np.array([2**40], dtype=np.int64))
    assert ovf
    s, ovf = oracle.sum_prod_i64(np.array([-3, 5], dtype=np.int64), np.array([7, 11], dtype=np.int64))
    assert not ovf and s == 34


def test_empty_and_full():
    n = 1000
    z = np.zeros((n + 63) // 64, dtype=np.uint64)
    assert len(oracle.decode(z)) == 0
    f = oracle.build_index(np.zeros(n, dtype=np.int32), 0, 1)[0]
    assert np.array_equal(oracle.decode(f), np.arange(n))
    assert len(oracle.decode(oracle.merge([[f], [z]]))) == 0


def test_double_sum_oracle():
    import math
    rng = np.random.default_rng(3)
    v = rng.normal(0, 1e6, 200_000) + 1e9
    plain, comp = oracle.sum_f64(v)
    exact = math.fsum(v.tolist())
    assert abs(comp - exact) <= 1e-15 * abs(exact)
    assert abs(plain - exact) <= 1e-12 * abs(exact)
