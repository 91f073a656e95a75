#!/usr/bin/env python
"""a compressed index over DENSE bitvectors (BITMAP containers: cardinality 100, k = 10, s = 0.1) against the verbatim
index of the same column — what the container path costs when nothing is sparse.  Usage: python tools/cmp_dense.py [rows]"""
import importlib, json, os, sys
from fractions import Fraction
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cubit = importlib.import_module("duckdb-cubit_b200")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
t = cubit.CubitTable(rows)
t.synth_column(1, 1, seed=0xC0B17, threshold=int(Fraction("0.1") * (1 << 64)), card=100, hot_lo=10, hot_n=10)
out = {}
for name, comp in (("verbatim", False), ("compressed", True)):
    ix = t.create_index(100, compressed=comp)
    t.build_index(ix, 1, 0)
    g = [[(ix, v) for v in range(10, 20)]]
    for variant, flags in (("count", 0), ("rowids", cubit.Q_ROWIDS)):
        ms = []
        for _ in range(5):
            with t.query(g, flags=flags | cubit.Q_TIMING) as r:
                ms.append(r.info.ms_scan)
                cnt = r.count
        out["%s_%s_ms" % (name, variant)] = min(ms)
        out["count"] = cnt
    if comp:
        info = t.index_info(ix)
        out["resident_bytes"], out["verbatim_bytes"] = int(info.resident_bytes), int(info.verbatim_bytes)
print(json.dumps(out))
t.close()
