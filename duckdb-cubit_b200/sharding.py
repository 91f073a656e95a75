"""Row-range sharding of a table over the GPUs of one box.

Bit r of every bitvector and row r of every column depend only on r, so the table is cut
into contiguous row ranges aligned to the CUBIT segment size; rank g owns rows
[lo_g, hi_g) of every bitvector, delta list and column.  The scan itself needs no
collective.  Afterwards:
  * COUNT / SUM  → one all-reduce of a few int64 limbs (exact: 128-bit sums are split into
    32-bit limbs so the int64 all-reduce cannot overflow),
  * row-ID lists → already global (row_base = lo_g) and globally sorted in rank order.
The reference's own parallel unit is the 122,880-row row group handed out by
RowGroupCollection::NextParallelScan (src/storage/table/row_group_collection.cpp:174-224).
"""

N_LIMBS = 5  # 4 x 32-bit limbs of the 128-bit sum + 1 limb for the count


def shard_ranges(n_rows, world_size, seg_bits):
    """→ [(lo, hi)] per rank; every lo is a multiple of seg_bits; ranges tile [0, n_rows)."""
    n_seg = (n_rows + seg_bits - 1) // seg_bits
    out = []
    for g in range(world_size):
        s0 = n_seg * g // world_size
        s1 = n_seg * (g + 1) // world_size
        out.append((min(s0 * seg_bits, n_rows), min(s1 * seg_bits, n_rows)))
    return out


def to_limbs(count, total):
    """(count, signed 128-bit sum) → N_LIMBS python ints, each < 2^32 (sum as two's complement)"""
    u = total & ((1 << 128) - 1)
    return [(u >> (32 * i)) & 0xFFFFFFFF for i in range(4)] + [count]


def from_limbs(limbs):
    """inverse of to_limbs after an element-wise SUM over ranks (limbs may exceed 2^32)"""
    u = sum(int(limbs[i]) << (32 * i) for i in range(4)) & ((1 << 128) - 1)
    if u >= 1 << 127:
        u -= 1 << 128
    return int(limbs[4]), u


def allreduce_aggregate(count, total, dist=None, device=None):
    """exact global (COUNT, SUM) over all ranks; torch.distributed does the plumbing"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return count, total
    import torch
    t = torch.tensor(to_limbs(count, total), dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return from_limbs(t.tolist())


class _DeviceArray:
    """zero-copy view of a device buffer owned by the library (CUDA array interface v2)"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def result_rowids_tensor(result, device):
    """the result's row-ID list as a torch tensor on `device` WITHOUT a copy (valid until result.free())"""
    import torch
    n = result.count
    ptr = result.info.d_rowids
    if n == 0 or not ptr:
        return torch.empty(0, dtype=torch.int64, device=device)
    return torch.as_tensor(_DeviceArray(ptr, n, "<i8"), device=device)


def gather_sorted(local, dist, dst=0):
    """Gather every rank's sorted row-ID tensor on rank `dst` (north_star: "per-shard row-ID lists are
    concatenated with shard offsets"; NCCL point-to-point over NVLink when the tensors are on GPUs).
    Shards are contiguous row ranges in rank order and their row IDs are already global, so plain
    concatenation in rank order IS the globally sorted list.  Returns the full tensor on dst, None elsewhere."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    counts[rank] = local.numel()
    dist.all_reduce(counts)  # an all-gather of the per-shard counts
    counts = counts.tolist()
    if rank == dst:
        out = torch.empty(sum(counts), dtype=local.dtype, device=local.device)
        off = [0]
        for c in counts:
            off.append(off[-1] + c)
        out[off[rank]:off[rank + 1]] = local
        ops = [dist.P2POp(dist.irecv, out[off[r]:off[r + 1]], r) for r in range(world) if r != dst and counts[r]]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out
    if local.numel():
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst)]):
            req.wait()
    return None
