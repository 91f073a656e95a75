#!/usr/bin/env python
"""Row-returning hand-off: the narrow wire (cubit_gpu_drain) against the wide copies, by worker count and window size.
Usage: python tools/drain_sweep.py [--rows N] [--sels 0.5,0.1,1e-2] [--threads 2,4,8,16] [--windows 131072,262144]
       [--payload-bits 0|24]
"""
import argparse
import importlib
import json
import os
import sys
import time
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=500_000_000)
    ap.add_argument("--sels", default="0.5,0.1,1e-2")
    ap.add_argument("--threads", default="2,4,8,16")
    ap.add_argument("--windows", default="0,262144")
    ap.add_argument("--payload-bits", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    cubit = importlib.import_module("duckdb-cubit_b200")
    t = cubit.CubitTable(args.rows)
    if args.payload_bits:
        t.synth_column(0, 3, seed=0xFEED, threshold=1 << args.payload_bits, hot_lo=90000)
    else:
        t.synth_column(0, 0)
    t.pack_column(0, keep_raw=True)
    for s in args.sels.split(","):
        thr = int(Fraction(s) * (1 << 64))
        t.synth_column(1, 1, seed=0xC0B17, threshold=thr, card=100, hot_lo=10, hot_n=10)
        ix = t.create_index(100)
        t.build_index(ix, 1, 0)
        plan = cubit.QueryPlan([[(ix, v) for v in range(10, 20)]], cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0],
                               agg=cubit.AGG_SUM, agg_a=0)
        with t.execute(plan) as r:
            cnt = r.count
            # wide, double-buffered (what bench.py's e2e_full_materialize times)
            win = 1 << 22
            bufs = [(cubit.HostBuffer(win * 8), cubit.HostBuffer(win * 8)) for _ in range(2)]
            arrs = [(a.array.view(np.int64), b.array.view(np.int64)) for a, b in bufs]
            best = None
            for _ in range(args.reps):
                t0 = time.perf_counter()
                nwin = (cnt + win - 1) // win
                tk = [None, None]
                if nwin:
                    tk[0] = r.fetch_async(0, min(win, cnt), arrs[0][0], [arrs[0][1]])
                for w in range(nwin):
                    if w + 1 < nwin:
                        o = (w + 1) * win
                        tk[(w + 1) & 1] = r.fetch_async(o, min(win, cnt - o), arrs[(w + 1) & 1][0], [arrs[(w + 1) & 1][1]])
                    r.fetch_wait(tk[w & 1])
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            for a, b in bufs:
                a.free()
                b.free()
            print(json.dumps({"s": s, "rows": cnt, "path": "wide", "ms": best * 1e3, "rows_per_s": cnt / best,
                              "pcie_GBps": cnt * 16 / best / 1e9}), flush=True)
            for win in [int(x) for x in args.windows.split(",")]:
                for th in [int(x) for x in args.threads.split(",")]:
                    best, st = None, None
                    for _ in range(args.reps):
                        t0 = time.perf_counter()
                        st = r.drain(threads=th, window_rows=win)
                        dt = time.perf_counter() - t0
                        best = dt if best is None else min(best, dt)
                    assert st.rows == cnt and st.sum_rowids == int(r.sum) % 2**64 if not args.payload_bits else st.rows == cnt
                    print(json.dumps({"s": s, "rows": cnt, "path": "narrow", "threads": th, "window_rows": win,
                                      "ms": best * 1e3, "rows_per_s": cnt / best, "wire_bytes_per_row": st.wire_bytes / max(1, cnt),
                                      "pcie_GBps": st.wire_bytes / best / 1e9}), flush=True)
    t.close()


if __name__ == "__main__":
    main()
