"""NULLs in probed columns: validity masks of the projected values and NULL-skipping aggregates, against the
reference's own answers (tests/golden/nulls.json, produced by the reference DuckDB through
tests/golden/make_null_golden.py) and against the oracle on random tables.

Reference semantics: StandardColumnData::FetchRow = validity.FetchRow + data
(src/storage/table/standard_column_data.cpp:169-178, ValidityFetchRow validity_uncompressed.cpp:381);
SUM skips NULL inputs (aggregate_executor.hpp:118-146); SUM over no non-NULL input is NULL.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COL_PRICE, COL_DISC = 0, 1


def validity_words(valid):
    b = np.packbits(np.asarray(valid, dtype=bool), bitorder="little")
    b = np.concatenate([b, np.zeros((-len(b)) % 8, dtype=np.uint8)])
    return b.view("<u8").astype(np.uint64)


def mask_bits(words, n):
    return np.unpackbits(np.ascontiguousarray(words, dtype="<u8").view(np.uint8), bitorder="little")[:n].astype(bool)


def sha(a, dt):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=dt).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def nulls():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "nulls.json")))


def null_patterns(n):
    r = np.arange(n, dtype=np.int64)
    return (r % 7 == 3) | ((r >= 1000) & (r < 1100)), r % 11 == 5  # as punched by make_null_golden.py


def bitmaps(lineitem):
    n = len(lineitem["price"])
    rng_bv = np.zeros((n + 63) // 64, dtype=np.uint64)  # rows [1000, 1100) as one more bitvector
    for r in range(1000, 1100):
        rng_bv[r // 64] |= np.uint64(1) << np.uint64(r % 64)
    return {"quantity": (oracle.build_index(lineitem["quantity"], 1, 50), 1),
            "discount": (oracle.build_index(lineitem["discount"], 0, 11), 0), "range": rng_bv}


def groups_of(name, ent):
    if ent["groups"] is not None:
        return ent["groups"]
    assert name == "all_null_price"
    return [[("quantity", 24)], [("range", 0)]]


def test_oracle_reproduces_reference_null_answers(nulls, lineitem):
    n = len(lineitem["price"])
    assert n == nulls["n_rows"]
    pn, dn = null_patterns(n)
    vp, vd = validity_words(~pn), validity_words(~dn)
    bm = bitmaps(lineitem)
    for name, ent in nulls["answers"].items():
        g = [[bm["range"] if c == "range" else bm[c][0][v - bm[c][1]] for (c, v) in grp] for grp in groups_of(name, ent)]
        ids = oracle.decode(oracle.merge(g))
        assert len(ids) == ent["count"] and sha(ids, "<i8") == ent["ids_sha256"], name
        pm, nv = oracle.probe_validity(ids, vp)
        dm, _ = oracle.probe_validity(ids, vd)
        assert nv == ent["count_p"]
        assert sha(np.packbits(mask_bits(pm, len(ids)), bitorder="little"), "u1") == ent["p_valid_sha256"]
        assert sha(np.packbits(mask_bits(dm, len(ids)), bitorder="little"), "u1") == ent["d_valid_sha256"]
        vals = oracle.probe(ids, lineitem["price"])
        assert sha(vals[mask_bits(pm, len(ids))], "<i8") == ent["p_values_sha256"]
        s, rows, ovf = oracle.sum_nulls(ids, lineitem["price"], vp)
        assert not ovf and rows == ent["count_p"] and (s == ent["sum_p"] if rows else ent["sum_p"] is None)
        s, rows, ovf = oracle.sum_nulls(ids, lineitem["price"], vp, lineitem["discount"], vd)
        assert not ovf and rows == ent["count_pd"] and (s == ent["sum_pd"] if rows else ent["sum_pd"] is None)


@pytest.mark.gpu
@pytest.mark.parametrize("seg_bits", [32768, 65536])
@pytest.mark.parametrize("packed", [False, True])
def test_gpu_null_semantics_against_reference_goldens(cubit, nulls, lineitem, seg_bits, packed):
    n = len(lineitem["price"])
    pn, dn = null_patterns(n)
    bm = bitmaps(lineitem)
    t = cubit.CubitTable(n, seg_bits=seg_bits)
    t.upload_column(COL_PRICE, lineitem["price"])
    t.upload_column(COL_DISC, lineitem["discount"])
    if packed:
        t.pack_column(COL_PRICE, keep_raw=False)
    t.upload_validity(COL_PRICE, validity_words(~pn))
    t.upload_validity(COL_DISC, validity_words(~dn))
    ix = {"quantity": t.upload_index(bm["quantity"][0]), "discount": t.upload_index(bm["discount"][0]),
          "range": t.upload_index(bm["range"][None, :])}
    base = {"quantity": 1, "discount": 0, "range": 0}
    for name, ent in nulls["answers"].items():
        refs = [[(ix[c], v - base[c]) for (c, v) in grp] for grp in groups_of(name, ent)]
        with t.query(refs, flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[COL_PRICE, COL_DISC], agg=cubit.AGG_SUM,
                     agg_a=COL_PRICE) as r:
            assert r.count == ent["count"] and r.agg_rows == ent["count_p"], name
            assert (r.sum == ent["sum_p"]) if ent["count_p"] else (ent["sum_p"] is None and r.sum == 0), name
            ids, (price, disc) = r.fetch()
            assert sha(ids, "<i8") == ent["ids_sha256"], name
            pm, p_all = r.fetch_validity(0)
            dm, d_all = r.fetch_validity(1)
            pv, dv = mask_bits(pm, len(ids)), mask_bits(dm, len(ids))
            assert sha(np.packbits(pv, bitorder="little"), "u1") == ent["p_valid_sha256"], name
            assert sha(np.packbits(dv, bitorder="little"), "u1") == ent["d_valid_sha256"], name
            assert p_all == bool(pv.all()) and d_all == bool(dv.all())
            assert sha(price[pv], "<i8") == ent["p_values_sha256"], name
            assert np.array_equal(disc[dv], lineitem["discount"][ids][dv])
            # DataChunk-sized fetches at offsets that are not word aligned
            for off, cnt in ((0, min(2048, len(ids))), (min(77, len(ids)), min(1000, len(ids) - min(77, len(ids)))),
                             (len(ids), 0)):
                w, _ = r.fetch_validity(0, off, cnt)
                assert np.array_equal(mask_bits(w, cnt), pv[off:off + cnt]), (name, off, cnt)
        with t.query(refs, flags=0, agg=cubit.AGG_SUM_PROD, agg_a=COL_PRICE, agg_b=COL_DISC) as r:
            assert r.count == ent["count"] and r.agg_rows == ent["count_pd"], name
            assert (r.sum == ent["sum_pd"]) if ent["count_pd"] else r.sum == 0, name
    t.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 63, 4097, 200_003])
def test_gpu_nulls_random_tables_vs_oracle(cubit, n):
    rng = np.random.default_rng(n)
    key = rng.integers(0, 8, n).astype(np.int64)
    a = rng.integers(-10**12, 10**12, n).astype(np.int64)
    b = rng.integers(-10**5, 10**5, n).astype(np.int64)
    va, vb = rng.random(n) > 0.3, rng.random(n) > 0.05
    t = cubit.CubitTable(n, row_base=5_000_000_000, seg_bits=32768)
    t.upload_column(0, a)
    t.upload_column(1, b)
    t.upload_column(2, key)
    t.upload_validity(0, validity_words(va))
    t.upload_validity(1, validity_words(vb))
    ix = t.create_index(8)
    t.build_index(ix, 2, 0)
    bv = oracle.build_index(key, 0, 8)
    for vals in ([3], [0, 1, 2, 3, 4, 5, 6, 7], [1, 6]):
        ids = oracle.decode(oracle.merge([[bv[v] for v in vals]]), row_base=5_000_000_000)
        with t.query([[(ix, v) for v in vals]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0, 1],
                     agg=cubit.AGG_SUM_PROD, agg_a=0, agg_b=1) as r:
            want, rows, ovf = oracle.sum_nulls(ids, a, validity_words(va), b, validity_words(vb), row_base=5_000_000_000)
            assert not ovf and r.count == len(ids) and r.agg_rows == rows and r.sum == want
            got_ids, (ga, gb) = r.fetch()
            assert np.array_equal(got_ids, ids)
            ma, _ = r.fetch_validity(0)
            mb, _ = r.fetch_validity(1)
            wa, _ = oracle.probe_validity(ids, validity_words(va), row_base=5_000_000_000)
            wb, _ = oracle.probe_validity(ids, validity_words(vb), row_base=5_000_000_000)
            assert np.array_equal(ma, wa) and np.array_equal(mb, wb)
            loc = ids - 5_000_000_000
            assert np.array_equal(ga[va[loc]], a[loc][va[loc]]) and np.array_equal(gb[vb[loc]], b[loc][vb[loc]])
    # a column without a mask reports all-valid; dropping a mask restores the NULL-free paths
    t.upload_validity(1, None)
    with t.query([[(ix, 3)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[1], agg=cubit.AGG_SUM, agg_a=1) as r:
        ids = oracle.decode(bv[3])
        assert r.agg_rows == r.count == len(ids) and r.sum == int(b[ids].sum())
        w, allv = r.fetch_validity(0)
        assert allv and mask_bits(w, len(ids)).all()
    # appended rows are valid
    extra = 70
    t.append_rows({0: rng.integers(0, 100, extra).astype(np.int64), 1: np.ones(extra, dtype=np.int64),
                   2: np.full(extra, 3, dtype=np.int64)})
    with t.query([[(ix, 3)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0]) as r:
        got_ids, _ = r.fetch()
        w, _ = r.fetch_validity(0)
        m = mask_bits(w, len(got_ids))
        old = got_ids - 5_000_000_000 < n
        assert np.array_equal(m[old], va[(got_ids - 5_000_000_000)[old]]) and m[~old].all() and (~old).sum() == extra
    t.close()
