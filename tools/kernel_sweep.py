#!/usr/bin/env python
"""Per-variant kernel timings of the scan path (CUDA events inside the library, Q_TIMING).
Variants isolate the streaming merge from the look-back, the row-ID emission and the probe.
Usage: python tools/kernel_sweep.py [--rows N] [--sels 1e-4,0.5] [--reps 5] [--seg-bits 65536] [--only fused]
"""
import argparse
import importlib
import json
import os
import sys
from fractions import Fraction

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--sels", default="1e-4,1e-2,0.1,0.5")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--seg-bits", type=int, default=65536)
    ap.add_argument("--only", default="")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--pack", action="store_true", help="store the payload column FOR-bit-packed")
    ap.add_argument("--keep-raw", action="store_true", help="with --pack: keep the raw column too (hybrid)")
    ap.add_argument("--payload-bits", type=int, default=0,
                    help="0: payload = row id (packs to 10 bits); b > 0: uniform random payload in [90000, 90000 + 2^b)")
    args = ap.parse_args()
    cubit = importlib.import_module("duckdb-cubit_b200")
    t = cubit.CubitTable(args.rows, seg_bits=args.seg_bits)
    if args.payload_bits:
        t.synth_column(0, 3, seed=0xFEED, threshold=1 << args.payload_bits, hot_lo=90000)
    else:
        t.synth_column(0, 0)
    if args.pack:
        print(json.dumps({"packed_payload_bytes": t.pack_column(0, keep_raw=args.keep_raw)}), flush=True)
    variants = {
        "count": dict(flags=0),
        "bitvector": dict(flags=cubit.Q_BITVECTOR),
        "rowids": dict(flags=cubit.Q_ROWIDS),
        "agg_only": dict(flags=0, agg=cubit.AGG_SUM, agg_a=0),
        "fused": dict(flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0], agg=cubit.AGG_SUM, agg_a=0),
        "unfused": dict(flags=cubit.Q_ROWIDS | cubit.Q_VALUES | cubit.Q_UNFUSED, cols=[0], agg=cubit.AGG_SUM, agg_a=0),
    }
    if args.only:
        variants = {k: v for k, v in variants.items() if k in args.only.split(",")}
    out = []
    for s in args.sels.split(","):
        thr = int(Fraction(s) * (1 << 64))
        t.synth_column(1, 1, seed=0xC0B17, threshold=thr, card=100, hot_lo=10, hot_n=10)
        ix = t.create_index(100)
        t.build_index(ix, 1, 0)
        groups = [[(ix, v) for v in range(10, 10 + args.k)]]
        for name, kw in variants.items():
            kw = dict(kw)
            kw["flags"] = kw["flags"] | cubit.Q_TIMING
            plan = cubit.QueryPlan(groups, **kw)
            ms, msp = [], []
            for i in range(args.reps + 2):
                with t.execute(plan) as r:
                    if i >= 2:
                        ms.append(r.info.ms_scan)
                        msp.append(r.info.ms_probe)
                    by = r.info.algo_bytes_scan + r.info.algo_bytes_probe
                    cnt = r.count
                    path = r.info.probe_path
            ms.sort()
            msp.sort()
            m = ms[len(ms) // 2]
            mp = msp[len(msp) // 2]
            rec = {"sel": s, "variant": name, "count": cnt, "ms_scan": round(m, 4), "ms_probe": round(mp, 4), "probe_path": path,
                   "algo_GBps": round(by / ((m + mp) * 1e-3) / 1e9, 1), "Grows_per_s": round(args.rows / ((m + mp) * 1e-3) / 1e9, 1)}
            print(json.dumps(rec), flush=True)
            out.append(rec)
    t.close()


if __name__ == "__main__":
    main()
