#!/usr/bin/env python
"""GPU decode of DuckDB BitPacking column segments (cubit_gpu_upload_column_segments) at scale.

Makes a TPC-H-like column on the host (l_extendedprice-like cents, 24 bits; a sorted key for DELTA_FOR; an
l_discount-like 4-bit column), encodes it with the oracle's writer (byte-identical to the reference's, see
tests/test_bitpacking.py), then times
    compressed upload + GPU decode   vs   plain upload of the decoded array   vs   the CPU decoder (oracle)
Algorithmic bytes of the decode kernel = compressed bytes read + 8 B/row written.
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (test infrastructure: makes the input and is the CPU baseline)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 27)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    cubit = importlib.import_module("duckdb-cubit_b200")
    n = args.rows
    rng = np.random.default_rng(1)
    cols = {
        "price_24bit_for": (90000 + rng.integers(0, 10404951, n)).astype(np.int64),
        "sorted_key_delta_for": np.cumsum(rng.integers(0, 60, n)).astype(np.int64),
        "discount_4bit": rng.integers(0, 11, n).astype(np.int64),
    }
    peak = 6451.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    with cubit.CubitTable(n) as t:
        for name, vals in cols.items():
            t0 = time.perf_counter()
            segs = oracle.encode_column_segments(vals, 122880 if "price" not in name else 86016)
            enc_s = time.perf_counter() - t0
            comp = sum(len(s[3]) for s in segs)
            best = None
            for _ in range(args.reps):
                t0 = time.perf_counter()
                info = t.upload_column_segments(0, 8, segs)
                wall = time.perf_counter() - t0
                if best is None:
                    best = (wall, info.ms_decode, info.h2d_bytes, info.n_groups)
                best = (min(wall, best[0]), min(info.ms_decode, best[1]), info.h2d_bytes, info.n_groups)
            ok = bool(np.array_equal(t.download_column(0), vals))
            plain = None
            for _ in range(args.reps):
                t0 = time.perf_counter()
                t.upload_column(1, vals)
                w = time.perf_counter() - t0
                plain = w if plain is None or w < plain else plain
            t0 = time.perf_counter()
            sample = segs[:max(1, len(segs) // 16)]
            rows_s = 0
            for _k, _s, c, b in sample:
                oracle.bitpacking_decode(b, 8, c)
                rows_s += c
            cpu_s = time.perf_counter() - t0
            algo = comp + 8 * n
            print(json.dumps({
                "column": name, "rows": n, "segments": len(segs), "groups": int(best[3]), "bit_exact": ok,
                "compressed_bytes": comp, "bits_per_value": round(8.0 * comp / n, 2),
                "decode_kernel_ms": round(best[1], 4), "decode_algo_GBps": round(algo / (best[1] * 1e-3) / 1e9, 1),
                "decode_frac_of_measured_hbm_peak": round(algo / (best[1] * 1e-3) / 1e9 / peak, 3),
                "decode_Grows_per_s": round(n / (best[1] * 1e-3) / 1e9, 1),
                "upload_compressed_plus_decode_wall_s": round(best[0], 4),
                "upload_decoded_plain_wall_s": round(plain, 4),
                "cpu_oracle_decode_Mrows_per_s_1thread": round(rows_s / cpu_s / 1e6, 1),
                "host_encode_s": round(enc_s, 2)}), flush=True)


if __name__ == "__main__":
    main()
