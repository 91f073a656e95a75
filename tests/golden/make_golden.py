#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the REFERENCE itself.

The reference binary is the DuckDB shell built from /root/reference (by the driver at
/tmp/duckdb_build/duckdb; rebuild recipe in SURVEY.md §8c).  It only exists in the build
container, so this script is run there once and its small outputs are committed:

  tpch_sf001.npz       TPC-H SF0.01 lineitem columns (60,175 rows) in rowid order
  golden_ids.npz       the row-ID lists / delta row lists belonging to golden.json
  golden.json          query-level known answers produced by reference SQL:
                         - row-ID sets / COUNT / SUM for bitmap-expressible predicates
                         - the same after SQL UPDATE/DELETE (pending-delta semantics)
                         - TPC-H Q6 at SF0.01 (and the answer file value)
                         - synthetic cfg-2 table (SplitMix64 generator) answers
                         - SF0.1 / SF1 aggregate answers + a record that the CPU oracle
                           reproduced them here (data too large to commit)
Usage:  python tests/golden/make_golden.py [--duckdb /path/to/duckdb] [--skip-sf1]
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

COLS_SQL = ("CAST(l_quantity*100 AS BIGINT) AS q, CAST(l_extendedprice*100 AS BIGINT) AS p, "
            "CAST(l_discount*100 AS BIGINT) AS d, CAST(l_shipdate - DATE '1970-01-01' AS BIGINT) AS s, "
            "(year(l_shipdate)-1992)*12 + month(l_shipdate)-1 AS m")


def sql(duck, db, stmt):
    r = subprocess.run([duck, db, "-csv", "-noheader", "-c", stmt], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True)
    if r.returncode != 0:
        raise RuntimeError("duckdb failed: %s\n%s" % (stmt, r.stderr))
    return r.stdout


def sql_ints(duck, db, stmt):
    out = sql(duck, db, stmt).strip()
    if not out:
        return np.zeros((0,), dtype=np.int64)
    rows = [[int(x) for x in line.split(",")] for line in out.splitlines()]
    return np.asarray(rows, dtype=np.int64)


def ids_digest(ids):
    return hashlib.sha256(np.ascontiguousarray(ids, dtype="<i8").tobytes()).hexdigest()


def export_lineitem(duck, db, tmp):
    path = os.path.join(tmp, "li.csv")
    sql(duck, db, "COPY (SELECT rowid AS r, %s FROM lineitem ORDER BY rowid) TO '%s' (HEADER false)" % (COLS_SQL, path))
    a = np.loadtxt(path, delimiter=",", dtype=np.int64, ndmin=2)
    os.unlink(path)
    assert (a[:, 0] == np.arange(len(a))).all(), "rowid is not the dense table position"
    return {"quantity": a[:, 1], "price": a[:, 2], "discount": a[:, 3], "shipdate": a[:, 4], "month": a[:, 5]}


# predicates expressed both as reference SQL and as bitmap groups over
# (quantity index: values 1..50 → value_id v-1; discount index: 0..10; month index: 0..83)
PREDICATES = [
    ("q_eq_24", "l_quantity = 24", [[("quantity", 24)]]),
    ("q_10_19", "l_quantity BETWEEN 10 AND 19", [[("quantity", v) for v in range(10, 20)]]),
    ("q_lt_24", "l_quantity < 24", [[("quantity", v) for v in range(1, 24)]]),
    ("q_ge_50", "l_quantity >= 50", [[("quantity", 50)]]),
    ("d_5_7", "l_discount BETWEEN 0.05 AND 0.07", [[("discount", v) for v in (5, 6, 7)]]),
    ("q6", "l_shipdate >= DATE '1994-01-01' AND l_shipdate < DATE '1995-01-01' AND "
           "l_discount BETWEEN 0.05 AND 0.07 AND l_quantity < 24",
     [[("month", m) for m in range(24, 36)], [("discount", v) for v in (5, 6, 7)],
      [("quantity", v) for v in range(1, 24)]]),
]


def reference_answers(duck, db, keep_ids):
    out = {}
    for name, where, groups in PREDICATES:
        ids = sql_ints(duck, db, "SELECT rowid FROM lineitem WHERE %s ORDER BY rowid" % where).reshape(-1)
        agg = sql_ints(duck, db,
                       "SELECT count(*), CAST(coalesce(sum(l_extendedprice),0)*100 AS HUGEINT), "
                       "CAST(coalesce(sum(l_extendedprice*l_discount),0)*10000 AS HUGEINT) FROM lineitem WHERE %s"
                       % where).reshape(-1)
        ent = {"where": where, "groups": groups, "count": int(agg[0]), "sum_price_cents": int(agg[1]),
               "sum_price_x_discount_e4": int(agg[2]), "ids_sha256": ids_digest(ids)}
        assert len(ids) == ent["count"]
        if keep_ids:
            ent["ids"] = ids.tolist()
        out[name] = ent
    return out


def oracle_answers(cols):
    """the CPU oracle on the exported columns (index build → merge → decode → probe → SUM)"""
    qv = cols["quantity"] // 100  # 1..50
    idx = {"quantity": (oracle.build_index(qv, 1, 50), 1), "discount": (oracle.build_index(cols["discount"], 0, 11), 0),
           "month": (oracle.build_index(cols["month"], 0, 84), 0)}
    out = {}
    for name, _where, groups in PREDICATES:
        g = [[idx[c][0][v - idx[c][1]] for (c, v) in grp] for grp in groups]
        ids = oracle.decode(oracle.merge(g))
        price = oracle.probe(ids, cols["price"])
        disc = oracle.probe(ids, cols["discount"])
        sp, ovf = oracle.sum_prod_i64(price, disc)
        assert not ovf
        out[name] = {"count": int(len(ids)), "sum_price_cents": oracle.sum_i64(price), "sum_price_x_discount_e4": sp,
                     "ids_sha256": ids_digest(ids)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--duckdb", default=os.environ.get("CUBIT_REF_DUCKDB", "/tmp/duckdb_build/duckdb"))
    ap.add_argument("--skip-sf1", action="store_true")
    args = ap.parse_args()
    duck = args.duckdb
    golden = {"reference": sql(duck, ":memory:", "select version()").strip(),
              "generator": "tests/golden/make_golden.py"}
    with tempfile.TemporaryDirectory() as tmp:
        # ---------------- TPC-H SF0.01: data + answers committed
        db = os.path.join(tmp, "sf001.db")
        sql(duck, db, "CALL dbgen(sf=0.01)")
        cols = export_lineitem(duck, db, tmp)
        np.savez_compressed(os.path.join(HERE, "tpch_sf001.npz"),
                            quantity=(cols["quantity"] // 100).astype(np.uint8), price=cols["price"].astype(np.int32),
                            discount=cols["discount"].astype(np.uint8), shipdate=cols["shipdate"].astype(np.int16),
                            month=cols["month"].astype(np.uint8))
        assert (cols["quantity"] % 100 == 0).all()
        ref = reference_answers(duck, db, keep_ids=True)
        orc = oracle_answers(cols)
        for k in ref:
            for f in ("count", "sum_price_cents", "sum_price_x_discount_e4", "ids_sha256"):
                assert ref[k][f] == orc[k][f], (k, f, ref[k][f], orc[k][f])
        q6 = sql(duck, db, "PRAGMA tpch(6)").strip()
        golden["tpch_sf001"] = {"n_rows": int(len(cols["price"])), "answers": ref, "pragma_tpch_6": q6,
                                "answer_file": "extension/tpch/dbgen/answers/sf0.01/q06.csv"}
        assert ref["q6"]["sum_price_x_discount_e4"] == int(round(float(q6) * 10000))

        # ---------------- pending deltas: SQL UPDATE / DELETE on the same table
        rng = np.random.default_rng(0xDE17A)
        n = len(cols["price"])
        chosen = rng.choice(n, size=n // 100, replace=False)
        upd, dele = np.sort(chosen[: len(chosen) // 2]), np.sort(chosen[len(chosen) // 2:])
        newq = (cols["quantity"][upd] // 100) % 50 + 1
        with open(os.path.join(tmp, "upd.csv"), "w") as f:
            for r, v in zip(upd.tolist(), newq.tolist()):
                f.write("%d,%d\n" % (r, v))
        with open(os.path.join(tmp, "del.csv"), "w") as f:
            for r in dele.tolist():
                f.write("%d\n" % r)
        sql(duck, db, "CREATE TABLE upd(rid BIGINT, newq BIGINT); COPY upd FROM '%s';"
                      "CREATE TABLE del(rid BIGINT); COPY del FROM '%s';"
                      "UPDATE lineitem SET l_quantity = upd.newq FROM upd WHERE lineitem.rowid = upd.rid;"
                      "DELETE FROM lineitem WHERE rowid IN (SELECT rid FROM del);"
            % (os.path.join(tmp, "upd.csv"), os.path.join(tmp, "del.csv")))
        ref_d = reference_answers(duck, db, keep_ids=True)
        golden["tpch_sf001_delta"] = {"updated_rows": upd.tolist(), "new_quantity": newq.tolist(),
                                      "deleted_rows": dele.tolist(), "answers": ref_d}

        # ---------------- synthetic cfg-2 style table through reference SQL
        n_syn, seed, card = 200_000, 0xC0B17, 100
        syn = {}
        for tag, s_num, s_den in (("s1e-2", 1, 100), ("s0.25", 1, 4)):
            thr = (s_num << 64) // s_den
            v = oracle.synth_column(1, n_syn, 0, seed, thr, card, 10, 10)
            path = os.path.join(tmp, "syn.csv")
            np.savetxt(path, v, fmt="%d")
            sdb = os.path.join(tmp, "syn_%s.db" % tag)
            sql(duck, sdb, "CREATE TABLE t(v INTEGER); COPY t FROM '%s';" % path)
            ids = sql_ints(duck, sdb, "SELECT rowid FROM t WHERE v BETWEEN 10 AND 19 ORDER BY rowid").reshape(-1)
            agg = sql_ints(duck, sdb, "SELECT count(*), CAST(coalesce(sum(rowid),0) AS HUGEINT) FROM t "
                                      "WHERE v BETWEEN 10 AND 19").reshape(-1)
            hist = sql_ints(duck, sdb, "SELECT v, count(*) FROM t GROUP BY v ORDER BY v")
            syn[tag] = {"n_rows": n_syn, "seed": seed, "threshold": str(thr), "card": card, "hot_lo": 10, "hot_n": 10,
                        "count": int(agg[0]), "sum_rowid": int(agg[1]), "ids_sha256": ids_digest(ids),
                        "histogram": hist.tolist()}
            bv = oracle.build_index(v, 0, card)
            oids = oracle.decode(oracle.merge([[bv[i] for i in range(10, 20)]]))
            assert ids_digest(oids) == syn[tag]["ids_sha256"] and int(oids.sum()) == syn[tag]["sum_rowid"]
        golden["synthetic"] = syn

        # ---------------- larger scale factors: answers only (oracle reproduced them here)
        for sf, tag in ((0.1, "tpch_sf01"), (1, "tpch_sf1")):
            if args.skip_sf1 and sf == 1:
                continue
            db2 = os.path.join(tmp, tag + ".db")
            sql(duck, db2, "CALL dbgen(sf=%s)" % sf)
            c2 = export_lineitem(duck, db2, tmp)
            r2 = reference_answers(duck, db2, keep_ids=False)
            o2 = oracle_answers(c2)
            for k in r2:
                for f in ("count", "sum_price_cents", "sum_price_x_discount_e4", "ids_sha256"):
                    assert r2[k][f] == o2[k][f], (tag, k, f)
                r2[k].pop("groups")
            golden[tag] = {"n_rows": int(len(c2["price"])), "answers": r2, "pragma_tpch_6": sql(duck, db2,
                           "PRAGMA tpch(6)").strip(), "oracle_reproduced_here": True}
            os.unlink(db2)
    # bulky integer lists go to a compressed npz next to the json
    arrays = {}
    for tag in ("tpch_sf001", "tpch_sf001_delta"):
        for name, ent in golden[tag]["answers"].items():
            arrays["%s/%s/ids" % (tag, name)] = np.asarray(ent.pop("ids"), dtype=np.int64)
    for f in ("updated_rows", "new_quantity", "deleted_rows"):
        arrays["tpch_sf001_delta/" + f] = np.asarray(golden["tpch_sf001_delta"].pop(f), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "golden_ids.npz"), **arrays)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    print("wrote golden.json;", {k: (v.get("n_rows") if isinstance(v, dict) else v) for k, v in golden.items()})


if __name__ == "__main__":
    main()
