import importlib, os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from fractions import Fraction
cubit = importlib.import_module("duckdb-cubit_b200")
n = 1_000_000_000
t = cubit.CubitTable(n)
t.synth_column(1, 1, seed=0xC0B17, threshold=int(Fraction("0.1") * (1 << 64)), card=100, hot_lo=10, hot_n=10)
for rep in range(3):
    ix = t.create_index(100)
    t0 = time.time(); t.build_index(ix, 1, 0); dt = time.time() - t0
    print("build card=100: %.2f ms -> %.2f TB/s (4 GB in + 12.5 GB out)" % (dt * 1e3, 16.5e9 / dt / 1e12), flush=True)
t.synth_column(2, 2, seed=5, card=2526, hot_lo=0)
t.close()
