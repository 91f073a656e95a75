#!/usr/bin/env python
"""Timing probe for container (compressed-index) scans: k ARRAY-container streams vs k verbatim streams on the same
rows (CUDA events inside the library).  Usage: python tools/cmp_probe.py [--rows 6000000]"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=6_000_000)
    ap.add_argument("--card", type=int, default=2526)
    a = ap.parse_args()
    cubit = importlib.import_module("duckdb-cubit_b200")
    t = cubit.CubitTable(a.rows)
    t.synth_column(1, 2, seed=99, card=a.card, hot_lo=0)
    cx = t.create_index(a.card, compressed=True)
    t.build_index(cx, 1, 0)
    vx = t.create_index(64)
    t.build_index(vx, 1, 1000)  # the same values 1000..1063, verbatim
    info = t.index_info(cx)
    print(json.dumps({"rows": a.rows, "card": a.card, "resident": int(info.resident_bytes), "verbatim": int(info.verbatim_bytes)}))
    for k in (1, 2, 4, 8, 16, 31, 60):
        row = {"k": k}
        for name, ix, base in (("containers", cx, 1000), ("verbatim", vx, 0)):
            for flags, label in ((0, "count"), (cubit.Q_ROWIDS, "rowids")):
                best = None
                for _ in range(7):
                    with t.query([[(ix, base + v) for v in range(k)]], flags=flags | cubit.Q_TIMING) as r:
                        ms = r.info.ms_scan
                        cnt = r.count
                    best = ms if best is None or ms < best else best
                row["%s_%s_ms" % (name, label)] = round(best, 4)
                row["%s_count" % name] = cnt
        print(json.dumps(row), flush=True)
    t.close()


if __name__ == "__main__":
    main()
