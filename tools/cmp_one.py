#!/usr/bin/env python
"""one container-scan query (k ARRAY streams) a few times — the target of an ncu capture"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cubit = importlib.import_module("duckdb-cubit_b200")
k = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 6_000_000
t = cubit.CubitTable(rows)
t.synth_column(1, 2, seed=99, card=2526, hot_lo=0)
cx = t.create_index(2526, compressed=True)
t.build_index(cx, 1, 0)
for _ in range(3):
    with t.query([[(cx, 1000 + v) for v in range(k)]], flags=cubit.Q_TIMING) as r:
        print(r.count, r.info.ms_scan)
t.close()
