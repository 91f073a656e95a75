"""CPU, world_size 2 over gloo: the multi-GPU host logic (row-range shards + exact aggregate all-reduce).
Each rank runs the ORACLE on its shard (the GPU kernels need a B200); what is under test is the sharding
arithmetic, the global row IDs and the limb-split 128-bit all-reduce that bench.py uses over NCCL."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_rows, seg_bits, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    sharding = importlib.import_module("duckdb-cubit_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_ranges(n_rows, world, seg_bits)[rank]
    # every rank generates ITS rows of the same global table (row-offset seeded generator)
    thr = (1 << 64) // 7
    col = oracle.synth_column(1, hi - lo, lo, 0xC0B17, thr, 100, 10, 10)
    pay = np.arange(lo, hi, dtype=np.int64) * 3 - 5_000_000_000_000
    bv = oracle.build_index(col, 0, 100)
    ids = oracle.decode(oracle.merge([[bv[v] for v in range(10, 20)]]), row_base=lo)
    vals = oracle.probe(ids, pay, row_base=lo)
    cnt, tot = sharding.allreduce_aggregate(len(ids), oracle.sum_i64(vals), dist)
    np.save(os.path.join(out_dir, "ids_%d.npy" % rank), ids)
    import torch
    full = sharding.gather_sorted(torch.from_numpy(ids), dist, dst=0)   # result gather (send/recv)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
    else:
        assert full is None
    if rank == 0:
        np.save(os.path.join(out_dir, "agg.npy"), np.array([cnt, tot >> 64, tot & ((1 << 64) - 1)], dtype=object),
                allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,seg_bits", [(300_001, 65536), (1_000_000, 131072)])
def test_two_rank_shards_equal_single_table(tmp_path, n_rows, seg_bits):
    import oracle
    sharding = importlib.import_module("duckdb-cubit_b200.sharding")
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_rows, seg_bits, str(tmp_path)), nprocs=world, join=True)
    # single-table oracle
    thr = (1 << 64) // 7
    col = oracle.synth_column(1, n_rows, 0, 0xC0B17, thr, 100, 10, 10)
    pay = np.arange(n_rows, dtype=np.int64) * 3 - 5_000_000_000_000
    bv = oracle.build_index(col, 0, 100)
    want = oracle.decode(oracle.merge([[bv[v] for v in range(10, 20)]]))
    got = np.concatenate([np.load(os.path.join(tmp_path, "ids_%d.npy" % r)) for r in range(world)])
    assert np.array_equal(got, want)  # concatenation in rank order is globally sorted
    assert np.array_equal(np.load(os.path.join(tmp_path, "gathered.npy")), want)  # gathered over the process group
    agg = np.load(os.path.join(tmp_path, "agg.npy"), allow_pickle=True)
    total = (int(agg[1]) << 64) + int(agg[2])
    assert int(agg[0]) == len(want) and total == oracle.sum_i64(pay[want])
    assert total < 0  # exercises the two's complement limb path


def test_shard_ranges_and_limbs():
    sharding = importlib.import_module("duckdb-cubit_b200.sharding")
    for n, w, seg in [(1, 8, 65536), (65536 * 8, 8, 65536), (10**9, 8, 65536), (16 * 10**9, 8, 131072), (1000, 3, 32768)]:
        r = sharding.shard_ranges(n, w, seg)
        assert r[0][0] == 0 and r[-1][1] == n
        for (a, b), (c, d) in zip(r, r[1:]):
            assert b == c and a % seg == 0 and c % seg == 0
    for cnt, tot in [(0, 0), (5, -1), (2**40, 2**100 + 12345), (7, -(2**120))]:
        limbs = sharding.to_limbs(cnt, tot)
        assert all(0 <= x < 2**32 for x in limbs[:4])
        assert sharding.from_limbs(limbs) == (cnt, tot)
        # element-wise sum over "ranks" stays exact
        s = [a + b for a, b in zip(limbs, sharding.to_limbs(3, -17))]
        assert sharding.from_limbs(s) == (cnt + 3, tot - 17)
