#!/usr/bin/env python
"""Multi-GPU correctness of the sharded path over NCCL (run under torchrun, one rank per GPU):
row-range shards of one global synthetic table, per-shard scan, exact aggregate all-reduce, and the
result gather of the row-ID lists to rank 0 (NCCL send/recv).  Properties checked on rank 0:
gathered list strictly ascending, length == all-reduced COUNT, Σ ids == all-reduced SUM(payload)
(payload = global row id), every shard's slice lies in its row range."""
import importlib
import os
import sys
from fractions import Fraction

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    cubit = importlib.import_module("duckdb-cubit_b200")
    sharding = importlib.import_module("duckdb-cubit_b200.sharding")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_global = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_003
    seg = 65536
    lo, hi = sharding.shard_ranges(n_global, world, seg)[rank]
    t = cubit.CubitTable(hi - lo, row_base=lo, seg_bits=seg, device=local)
    t.set_stream(torch.cuda.current_stream().cuda_stream)
    t.synth_column(1, 1, seed=0xC0B17, threshold=int(Fraction("0.01") * (1 << 64)), card=100, hot_lo=10, hot_n=10)
    ix = t.create_index(100)
    t.build_index(ix, 1, 0)
    t.synth_column(0, 0)
    with t.query([[(ix, v) for v in range(10, 18)]], flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=0) as r:
        cnt, tot = sharding.allreduce_aggregate(r.count, r.sum, dist, dev)
        local_ids = sharding.result_rowids_tensor(r, dev)
        assert local_ids.numel() == r.count
        if r.count:
            assert int(local_ids[0]) >= lo and int(local_ids[-1]) < hi
        full = sharding.gather_sorted(local_ids, dist, dst=0)
        if rank == 0:
            assert full.numel() == cnt, (full.numel(), cnt)
            assert bool((full[1:] > full[:-1]).all())
            # Σ ids fits int64 here (n_global * count < 2^63)
            assert int(full.sum()) == tot, (int(full.sum()), tot)
            print("multi_gpu_check ok: world=%d rows=%d selected=%d sum=%d" % (world, n_global, cnt, tot))
    dist.barrier()
    t.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
