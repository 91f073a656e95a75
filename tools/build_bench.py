#!/usr/bin/env python
"""Index-build timing (CREATE INDEX analog): 10^9 rows, cardinality 100.
Wall-clock around the synchronous C-ABI calls (each ends with a stream sync)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cubit = importlib.import_module("duckdb-cubit_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
t = cubit.CubitTable(n)
t.synth_column(1, 1, seed=1, threshold=(1 << 64) // 10, card=100, hot_lo=10, hot_n=10)
ix = t.create_index(100)
for i in range(4):
    t0 = time.perf_counter(); t.build_index(ix, 1, 0); dt = time.perf_counter() - t0
    by = n * 4 + 100 * ((n + 63) // 64) * 8
    print("build %d: %.3f ms  %.1f G rows/s  %.0f GB/s (4 B/row read + 12.5 B/row written)" % (i, dt * 1e3, n / dt / 1e9, by / dt / 1e9), flush=True)
t0 = time.perf_counter(); c = [t.bitvector_count(ix, v) for v in range(100)]; dt = time.perf_counter() - t0
print("popcount of 100 bitvectors: %.3f ms, total %d" % (dt * 1e3, sum(c)))
assert sum(c) == n
