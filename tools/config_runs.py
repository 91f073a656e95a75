#!/usr/bin/env python
"""BASELINE.json configs 1, 3, 4 and the per-GPU slice of config 5 on one B200: full-size runs with
size-independent correctness properties and CUDA-event kernel timings (Q_TIMING).  These are the
"parity-test cases, not bench lines" of the measurement contract; results go to profiles/.
Usage: python tools/config_runs.py [--out profiles/r1_configs] [--only cfg1,cfg3,cfg4,cfg5]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M64
    return x ^ (x >> np.uint64(31))


def synth_value_np(seed, rows, threshold, card, hot_lo, hot_n):
    """numpy restatement of the kind-1 generator for scattered rows (tools only)"""
    with np.errstate(over="ignore"):
        z = splitmix64(np.uint64(seed) + rows.astype(np.uint64))
    y = z >> np.uint64(7)
    hot = z < np.uint64(threshold)
    v = np.where(hot, np.uint64(hot_lo) + y % np.uint64(hot_n), 0).astype(np.int64)
    cold = (y % np.uint64(card - hot_n)).astype(np.int64)
    cold = np.where(cold >= hot_lo, cold + hot_n, cold)
    return np.where(hot, v, cold)


def timed(cubit, t, groups, reps=5, **kw):
    kw = dict(kw)
    kw["flags"] = kw.get("flags", 0) | cubit.Q_TIMING
    plan = cubit.QueryPlan(groups, **kw)
    ms = []
    out = None
    for i in range(reps + 2):
        with t.execute(plan) as r:
            if i >= 2:
                ms.append((r.info.ms_scan, r.info.ms_probe))
            out = dict(count=r.count, sum=r.sum, k=r.info.n_streams, launches=r.info.n_launches,
                       algo_bytes=int(r.info.algo_bytes_scan + r.info.algo_bytes_probe),
                       delta_entries=int(r.info.delta_entries))
    ms.sort(key=lambda x: x[0] + x[1])
    m = ms[len(ms) // 2]
    out.update(ms_scan=round(m[0], 4), ms_probe=round(m[1], 4), ms_total=round(m[0] + m[1], 4))
    return out


def cfg1(cubit):
    """TPC-H SF1 shape: 6,001,215 rows, 50-value index, equality predicate + SUM(l_extendedprice)"""
    n = 6_001_215
    t = cubit.CubitTable(n)
    t.synth_column(1, 2, seed=11, card=50, hot_lo=1)               # l_quantity
    t.synth_column(0, 3, seed=12, threshold=10_410_000, hot_lo=90_000)  # l_extendedprice (cents)
    ix = t.create_index(50)
    t.build_index(ix, 1, 1)
    g = [[(ix, 23)]]
    agg = timed(cubit, t, g, agg=cubit.AGG_SUM, agg_a=0)
    ids = timed(cubit, t, g, flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0], agg=cubit.AGG_SUM, agg_a=0)
    # end-to-end latency of the synchronous C-ABI call (host predicate in → COUNT/SUM on the host), wall clock
    lat = {}
    for name, kw in (("aggregate_only", dict(agg=cubit.AGG_SUM, agg_a=0)),
                     ("rowids_values_first_chunk", dict(flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0], agg=cubit.AGG_SUM, agg_a=0))):
        plan = cubit.QueryPlan(g, **kw)
        ids_h, val_h = np.empty(2048, dtype=np.int64), [np.empty(2048, dtype=np.int64)]
        ts = []
        for i in range(300):
            t0 = time.perf_counter()
            with t.execute(plan) as r:
                c = r.count
                if "flags" in kw:
                    r.fetch(0, min(2048, c), out_ids=ids_h, out_cols=val_h)
            ts.append(time.perf_counter() - t0)
        ts = sorted(ts[50:])
        lat[name] = {"median_us": round(ts[len(ts) // 2] * 1e6, 1), "p10_us": round(ts[len(ts) // 10] * 1e6, 1)}
    qty, price = t.download_column(1), t.download_column(0)
    want = np.flatnonzero(qty == 24)
    assert agg["count"] == ids["count"] == len(want) and agg["sum"] == ids["sum"] == int(price[want].sum())
    t.close()
    return {"n_rows": n, "k": 1, "selected": len(want), "aggregate_only": agg, "rowids_values_sum": ids,
            "e2e_latency_sync_call": lat,
            "note": "launch-latency dominated (750 KB bitvector): report microseconds, not % roofline",
            "check": "COUNT and SUM equal a host evaluation of the downloaded columns"}


def cfg3(cubit):
    """TPC-H SF100 shape Q6: (12 months) AND (3 discounts) AND (23 quantities), k = 38, SUM(price*discount)"""
    n = 600_037_902
    t = cubit.CubitTable(n)
    t.synth_column(0, 3, seed=21, threshold=10_410_000, hot_lo=90_000)   # l_extendedprice
    t.synth_column(1, 3, seed=22, threshold=11, hot_lo=0)                # l_discount (cents) as int64
    ix_d = t.create_index(11)
    t.build_index(ix_d, 1, 0)
    t.synth_column(2, 2, seed=23, card=50, hot_lo=1)                     # l_quantity
    ix_q = t.create_index(50)
    t.build_index(ix_q, 2, 1)
    t.synth_column(2, 2, seed=24, card=84, hot_lo=0)                     # month bin of l_shipdate
    ix_m = t.create_index(84)
    t.build_index(ix_m, 2, 0)
    t.drop_column(2)
    gm = [(ix_m, m) for m in range(24, 36)]
    gd = [(ix_d, v) for v in (5, 6, 7)]
    gq = [(ix_q, v) for v in range(0, 23)]
    q6 = [gm, gd, gq]
    agg = timed(cubit, t, q6, agg=cubit.AGG_SUM_PROD, agg_a=0, agg_b=1)
    mat = timed(cubit, t, q6, flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0, 1], agg=cubit.AGG_SUM_PROD, agg_a=0, agg_b=1)
    unf = timed(cubit, t, q6, reps=1, flags=cubit.Q_ROWIDS | cubit.Q_UNFUSED, agg=cubit.AGG_SUM_PROD, agg_a=0, agg_b=1)
    assert agg["count"] == mat["count"] == unf["count"] and agg["sum"] == mat["sum"] == unf["sum"]
    # property: the merged bitvector equals the AND of the three single-group bitvectors
    qs = []
    for g in q6:
        with t.query([g], flags=cubit.Q_BITVECTOR) as r:
            qs.append(r.bitvector())
    with t.query(q6, flags=cubit.Q_BITVECTOR | cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0, 1]) as r:
        qall = r.bitvector()
        assert np.array_equal(qall, qs[0] & qs[1] & qs[2])
        ids, (price, disc) = r.fetch()
        assert (np.diff(ids) > 0).all() and disc.min() >= 5 and disc.max() <= 7
        assert int((price * disc).sum()) == agg["sum"]
        bits = np.unpackbits(qall.view(np.uint8), bitorder="little")
        assert np.array_equal(np.flatnonzero(bits), ids)
    t.close()
    return {"n_rows": n, "k": 38, "groups": [12, 3, 23], "selected": agg["count"], "selectivity": agg["count"] / n,
            "aggregate_only": agg, "rowids_values_sum": mat, "three_kernel_path": unf,
            "binning": "l_shipdate binned by month (84 bins); synthetic uniform columns of the SF100 lineitem shape",
            "check": "Q == Q_month & Q_disc & Q_qty; decode(Q) == row IDs; sum(price*disc) over fetched rows == fused SUM; "
                     "fused == three-kernel path"}


def cfg4(cubit):
    """read-mostly mix: cfg2 scan (s = 0.1) with pending update/delete deltas on 1 % of the rows"""
    from fractions import Fraction
    n = 1_000_000_000
    seed, thr = 0xC0B17, int(Fraction("0.1") * (1 << 64))
    t = cubit.CubitTable(n)
    t.synth_column(1, 1, seed=seed, threshold=thr, card=100, hot_lo=10, hot_n=10)
    ix = t.create_index(100)
    t.build_index(ix, 1, 0)
    t.drop_column(1)
    t.synth_column(0, 0)
    g = [[(ix, v) for v in range(10, 20)]]
    base = timed(cubit, t, g, flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=0)
    # 1 % of the rows: half updates v -> (v+1) % 100, half deletes
    rng = np.random.default_rng(0xDE17A)
    rows = np.unique(rng.integers(0, n, n // 100, dtype=np.int64))
    v = synth_value_np(seed, rows, thr, 100, 10, 10)
    is_upd = (np.arange(len(rows)) & 1) == 0
    # what the DML produces: one (value, row) pair per flipped bit — every touched row leaves B_v, an updated row
    # also enters B_(v+1); ingested incrementally on the device (cubit_gpu_add_delta_pairs), in four statements
    t.set_merge_threshold(ix, 1.0)                                   # keep the deltas pending: this config measures the XOR
    pv = np.concatenate([v, (v[is_upd] + 1) % 100]).astype(np.uint32)
    pr = np.concatenate([rows, rows[is_upd]])
    t0 = time.time()
    for part in range(4):
        t.add_delta_pairs(ix, pv[part::4], pr[part::4])
    with t.query(g, flags=0) as r0:                                  # first scan: the ingestion has completed behind it
        pass
    set_s = time.time() - t0
    inr = (v >= 10) & (v <= 19)
    expect = base["count"] - int(inr.sum()) + int((is_upd & (v >= 9) & (v <= 18)).sum())
    with_d = timed(cubit, t, g, flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=0)
    assert with_d["count"] == expect, (with_d["count"], expect)
    assert with_d["delta_entries"] > 0
    t0 = time.time()
    t.merge_deltas(ix)
    merge_s = time.time() - t0
    merged = timed(cubit, t, g, flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=0)
    assert merged["count"] == expect and merged["sum"] == with_d["sum"] and merged["delta_entries"] == 0
    t.close()
    return {"n_rows": n, "k": 10, "delta_rows": int(len(rows)), "no_deltas": base, "deltas_xor_at_query_time": with_d,
            "after_merge_back": merged, "delta_pairs": int(len(pr)), "ingest_4_statements_and_first_scan_s": round(set_s, 4),
            "merge_back_s": round(merge_s, 3),
            "check": "COUNT == base - rows leaving the range + rows entering it (host arithmetic on the delta list); "
                     "query-time XOR == merge-back (COUNT, SUM)"}


def cfg5(cubit):
    """per-GPU slice of the 16e9-row / 8-GPU config: 2e9 rows, 8 predicates"""
    from fractions import Fraction
    n = 2_000_000_000
    t = cubit.CubitTable(n)
    t.synth_column(1, 1, seed=0xC0B17, threshold=int(Fraction("0.1") * (1 << 64)), card=100, hot_lo=10, hot_n=10)
    ix = t.create_index(100)
    t.build_index(ix, 1, 0)
    t.drop_column(1)
    t.synth_column(0, 0)
    or8 = [[(ix, v) for v in range(10, 18)]]
    and44 = [[(ix, v) for v in (10, 11, 12, 13)], [(ix, v) for v in (12, 13, 14, 15)]]
    a = timed(cubit, t, or8, flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=0)
    b = timed(cubit, t, and44, flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=0)
    c12 = sum(t.bitvector_count(ix, v) for v in (12, 13))
    c8 = sum(t.bitvector_count(ix, v) for v in range(10, 18))
    assert a["count"] == c8 and b["count"] == c12
    t.close()
    return {"n_rows": n, "or_of_8": a, "and_of_two_or_of_4": b,
            "check": "disjoint value bitvectors: OR-of-8 selects the sum of the 8 popcounts, "
                     "(10..13) AND (12..15) selects popcount(12)+popcount(13)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_configs"))
    ap.add_argument("--only", default="cfg1,cfg3,cfg4,cfg5")
    args = ap.parse_args()
    cubit = importlib.import_module("duckdb-cubit_b200")
    res = {}
    for name, fn in (("cfg1", cfg1), ("cfg3", cfg3), ("cfg4", cfg4), ("cfg5", cfg5)):
        if name in args.only.split(","):
            t0 = time.time()
            res[name] = fn(cubit)
            res[name]["wall_s"] = round(time.time() - t0, 1)
            print(name, json.dumps(res[name])[:600], flush=True)
    with open(args.out + ".json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
