#!/usr/bin/env python
"""BASELINE.json config 5: ONE synthetic 16·10^9-row table, row-range sharded over the GPUs of the box, "8 predicates"
(both readings: OR over 8 value bitvectors, and AND of two OR-of-4 groups), STRONG scaling: the table is the same
for every GPU count (the generator is seeded by the global row id, so any shard count yields identical data).
Run under torchrun (one rank per GPU) or plainly for 1 GPU.  Per query: CUDA-event kernel times inside the
library, max over ranks; COUNT / SUM all-reduced exactly over NCCL and checked against size-independent
properties (disjoint value bitvectors → COUNT = Σ popcounts; payload = global row id → SUM = Σ row ids of the
shard's result, compared with the sum of the materialised row IDs).
Usage: torchrun --nproc-per-node G tools/cfg5_scaling.py [--rows 16000000000] [--out gpurun_out/cfg5_G.json]"""
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
    ap.add_argument("--rows", type=int, default=16_000_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--no-payload", action="store_true", help="merge + decode only (row IDs), no payload probe")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    cubit = importlib.import_module("duckdb-cubit_b200")
    sharding = importlib.import_module("duckdb-cubit_b200.sharding")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    seg = 65536
    lo, hi = sharding.shard_ranges(args.rows, world, seg)[rank]
    n = hi - lo
    with_payload = n <= 8_000_000_000 and not args.no_payload  # 8 B/row payload + index + row IDs must fit 180 GB
    t = cubit.CubitTable(n, row_base=lo, seg_bits=seg, device=local)
    t.set_stream(torch.cuda.current_stream().cuda_stream)
    # key column: 10 % of the rows fall uniformly on values 10..19 (0.01 each), the rest on the other 90 values
    t.synth_column(1, 1, seed=0xC0B17, threshold=int(Fraction("0.1") * (1 << 64)), card=100, hot_lo=10, hot_n=10)
    ix = t.create_index(18)          # only the bitvectors of values 0..17 are kept (the queries read 10..17)
    t.build_index(ix, 1, 0)
    t.drop_column(1)
    if with_payload:
        t.synth_column(0, 0)         # payload = global row id
    pop = [t.bitvector_count(ix, v) for v in range(18)]
    queries = {"or_of_8": ([[(ix, v) for v in range(10, 18)]], sum(pop[10:18])),
               "and_of_two_or_of_4": ([[(ix, v) for v in (10, 11, 12, 13)], [(ix, v) for v in (12, 13, 14, 15)]],
                                      pop[12] + pop[13])}
    res = {"n_rows_total": args.rows, "n_gpus": world, "rows_per_gpu": n, "payload": with_payload, "scaling": "strong"}
    for name, (groups, want_local) in queries.items():
        kw = dict(flags=cubit.Q_ROWIDS | cubit.Q_TIMING)
        if with_payload:
            kw.update(agg=cubit.AGG_SUM, agg_a=0)
        plan = cubit.QueryPlan(groups, **kw)
        times = []
        for i in range(args.reps + 2):
            if world > 1:
                dist.barrier()
            with t.execute(plan) as r:
                ms = r.info.ms_scan + r.info.ms_probe
                if i >= 2:
                    times.append(ms)
                if i == 0:
                    assert r.count == want_local, (name, r.count, want_local)
                    ids = sharding.result_rowids_tensor(r, dev)
                    assert ids.numel() == r.count and bool((ids[1:] > ids[:-1]).all())
                    assert int(ids[0]) >= lo and int(ids[-1]) < hi
                    # Σ row ids of the shard in exact integer arithmetic (chunked: int64 partial sums stay below 2^63)
                    s = 0
                    for c0 in range(0, ids.numel(), 1 << 24):
                        s += int(ids[c0:c0 + (1 << 24)].sum())
                    if with_payload:
                        assert r.sum == s, (name, r.sum, s)
                    cnt_g, sum_g = sharding.allreduce_aggregate(r.count, s, dist if world > 1 else None, dev)
                    algo = r.info.algo_bytes_scan + r.info.algo_bytes_probe
        times.sort()
        med = torch.tensor([times[len(times) // 2]], dtype=torch.float64, device=dev)
        ab = torch.tensor([float(algo)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(med, op=dist.ReduceOp.MAX)
            dist.all_reduce(ab, op=dist.ReduceOp.SUM)
        ms = float(med.item())
        res[name] = {"k": 8, "selected": cnt_g, "sum_rowids": str(sum_g), "ms_max_over_ranks": round(ms, 4),
                     "rows_per_s": args.rows / (ms * 1e-3), "algo_GBps_all_gpus": float(ab.item()) / (ms * 1e-3) / 1e9}
    if rank == 0:
        line = json.dumps(res)
        print(line, flush=True)
        if args.out:
            with open(args.out, "w") as f:
                f.write(line + "\n")
    if world > 1:
        dist.barrier()
    t.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
