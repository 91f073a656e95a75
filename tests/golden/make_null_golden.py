#!/usr/bin/env python
"""NULL semantics of the column probe, pinned to the REFERENCE itself (the DuckDB shell built from
/root/reference; build container only): TPC-H SF0.01 lineitem with a deterministic NULL pattern punched into
l_extendedprice and l_discount by reference SQL, then, per predicate,
    SELECT rowid, p, d ... ORDER BY rowid     → which result rows are NULL (validity masks of the probe)
    SELECT count(*), count(p), sum(p), count(p*d), sum(p*d)
The NULL pattern is a pure function of rowid (restated in tests), so only the answers are committed:
tests/golden/nulls.json.   Usage: python tests/golden/make_null_golden.py [--duckdb /path/to/duckdb]
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

# p (price) is NULL when rowid % 7 == 3 or rowid in [1000, 1100); d (discount) is NULL when rowid % 11 == 5
P_NULL_SQL = "(rowid % 7 = 3 OR (rowid >= 1000 AND rowid < 1100))"
D_NULL_SQL = "(rowid % 11 = 5)"

PREDICATES = [
    ("q_eq_24", "q = 24", [[("quantity", 24)]]),
    ("q_10_19", "q BETWEEN 10 AND 19", [[("quantity", v) for v in range(10, 20)]]),
    ("q_lt_24_d_5_7", "q < 24 AND dk BETWEEN 5 AND 7", [[("quantity", v) for v in range(1, 24)],
                                                        [("discount", v) for v in (5, 6, 7)]]),
    ("all_null_price", "q = 24 AND rowid >= 1000 AND rowid < 1100", None),  # SUM over only-NULL inputs is NULL
]


def p_null(r):
    return (r % 7 == 3) | ((r >= 1000) & (r < 1100))


def d_null(r):
    return r % 11 == 5


def validity_words(valid):
    """bool per row → DuckDB ValidityMask words (bit r%64 of word r/64)"""
    b = np.packbits(np.asarray(valid, dtype=bool), bitorder="little")
    b = np.concatenate([b, np.zeros((-len(b)) % 8, dtype=np.uint8)])
    return b.view("<u8").astype(np.uint64)


def sql(duck, db, stmt):
    r = subprocess.run([duck, db, "-csv", "-noheader", "-nullvalue", "NULL", "-c", stmt], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        raise RuntimeError("duckdb failed: %s\n%s" % (stmt, r.stderr))
    return r.stdout


def digest(a, dt):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=dt).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--duckdb", default=os.environ.get("CUBIT_REF_DUCKDB", "/tmp/duckdb_build/duckdb"))
    args = ap.parse_args()
    duck = args.duckdb
    li = np.load(os.path.join(HERE, "tpch_sf001.npz"))
    n = len(li["quantity"])
    rows = np.arange(n, dtype=np.int64)
    out = {"reference": sql(duck, ":memory:", "select version()").strip(), "generator": "tests/golden/make_null_golden.py",
           "p_null": "rowid % 7 == 3 or 1000 <= rowid < 1100", "d_null": "rowid % 11 == 5", "n_rows": int(n), "answers": {}}
    with tempfile.TemporaryDirectory() as tmp:
        db = os.path.join(tmp, "nulls.db")
        sql(duck, db, "CALL dbgen(sf=0.01)")
        # q = quantity (1..50), dk = discount key (0..10, never NULL: the index is built on it), p / d with NULLs
        sql(duck, db, "CREATE TABLE t AS SELECT CAST(l_quantity AS BIGINT) AS q, CAST(l_discount*100 AS BIGINT) AS dk, "
                      "CAST(l_extendedprice*100 AS BIGINT) AS p, CAST(l_discount*100 AS BIGINT) AS d "
                      "FROM lineitem ORDER BY rowid")
        sql(duck, db, "UPDATE t SET p = NULL WHERE %s" % P_NULL_SQL)
        sql(duck, db, "UPDATE t SET d = NULL WHERE %s" % D_NULL_SQL)
        chk = sql(duck, db, "SELECT count(*), count(p), count(d) FROM t").strip().split(",")
        assert int(chk[0]) == n and int(chk[1]) == n - int(p_null(rows).sum()) and int(chk[2]) == n - int(d_null(rows).sum())
        for name, where, groups in PREDICATES:
            lines = sql(duck, db, "SELECT rowid, p, d FROM t WHERE %s ORDER BY rowid" % where).strip().splitlines()
            ids = np.array([int(x.split(",")[0]) for x in lines], dtype=np.int64)
            pv = np.array([x.split(",")[1] != "NULL" for x in lines], dtype=bool)
            dv = np.array([x.split(",")[2] != "NULL" for x in lines], dtype=bool)
            pvals = np.array([0 if x.split(",")[1] == "NULL" else int(x.split(",")[1]) for x in lines], dtype=np.int64)
            agg = sql(duck, db, "SELECT count(*), count(p), CAST(sum(p) AS HUGEINT), count(p*d), CAST(sum(p*d) AS HUGEINT) "
                                "FROM t WHERE %s" % where).strip().split(",")
            ent = {"where": where, "groups": groups, "count": int(agg[0]), "count_p": int(agg[1]),
                   "sum_p": None if agg[2] == "NULL" else int(agg[2]), "count_pd": int(agg[3]),
                   "sum_pd": None if agg[4] == "NULL" else int(agg[4]),
                   "ids_sha256": digest(ids, "<i8"), "p_valid_sha256": digest(np.packbits(pv, bitorder="little"), "u1"),
                   "d_valid_sha256": digest(np.packbits(dv, bitorder="little"), "u1"),
                   "p_values_sha256": digest(pvals[pv], "<i8")}
            # the reference's NULL positions are exactly the punched pattern
            assert (pv == ~p_null(ids)).all() and (dv == ~d_null(ids)).all()
            # ... and the CPU oracle reproduces the reference on the same data
            price = li["price"].astype(np.int64)
            disc = li["discount"].astype(np.int64)
            vp = validity_words(~p_null(rows))
            vd = validity_words(~d_null(rows))
            mask, nv = oracle.probe_validity(ids, vp)
            assert nv == ent["count_p"]
            s, r_, ovf = oracle.sum_nulls(ids, price, vp)
            assert not ovf and r_ == ent["count_p"] and (s == ent["sum_p"] if r_ else ent["sum_p"] is None)
            s2, r2, ovf = oracle.sum_nulls(ids, price, vp, disc, vd)
            assert not ovf and r2 == ent["count_pd"] and (s2 == ent["sum_pd"] if r2 else ent["sum_pd"] is None)
            out["answers"][name] = ent
    with open(os.path.join(HERE, "nulls.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote nulls.json:", {k: (v["count"], v["count_p"], v["sum_p"]) for k, v in out["answers"].items()})


if __name__ == "__main__":
    main()
