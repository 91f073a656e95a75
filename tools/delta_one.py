#!/usr/bin/env python
"""config 4 in isolation: the s = 0.1 cfg2 scan with pending deltas on 1 % of the rows (XOR at query time), a few
times — the target of an ncu capture / A-B timing.  Usage: python tools/delta_one.py [--rows N] [--reps 5]"""
import argparse
import importlib
import json
import os
import sys
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from config_runs import synth_value_np, timed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    cubit = importlib.import_module("duckdb-cubit_b200")
    n = a.rows
    seed, thr = 0xC0B17, int(Fraction("0.1") * (1 << 64))
    t = cubit.CubitTable(n)
    t.synth_column(1, 1, seed=seed, threshold=thr, card=100, hot_lo=10, hot_n=10)
    ix = t.create_index(100)
    t.build_index(ix, 1, 0)
    t.drop_column(1)
    t.synth_column(0, 0)
    g = [[(ix, v) for v in range(10, 20)]]
    base = timed(cubit, t, g, reps=a.reps, flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=0)
    rng = np.random.default_rng(0xDE17A)
    rows = np.unique(rng.integers(0, n, n // 100, dtype=np.int64))
    v = synth_value_np(seed, rows, thr, 100, 10, 10)
    is_upd = (np.arange(len(rows)) & 1) == 0
    t.set_merge_threshold(ix, 1.0)
    t.add_delta_pairs(ix, np.concatenate([v, (v[is_upd] + 1) % 100]).astype(np.uint32), np.concatenate([rows, rows[is_upd]]))
    with_d = timed(cubit, t, g, reps=a.reps, flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=0)
    cnt = timed(cubit, t, g, reps=a.reps, flags=0)
    print(json.dumps({"no_deltas_ms_scan": base["ms_scan"], "deltas_ms_scan": with_d["ms_scan"],
                      "overhead": round(with_d["ms_scan"] / base["ms_scan"] - 1, 4), "count_only_with_deltas_ms": cnt["ms_scan"],
                      "delta_entries": with_d["delta_entries"]}))
    t.close()


if __name__ == "__main__":
    main()
