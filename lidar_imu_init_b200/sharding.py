"""Scan sharding for frames of >= 1M points: the map is replicated on every GPU, scan points
are block-partitioned (contiguous shards keep each rank's queries spatially compact), and the only exchange per ICP
iteration is one all-reduce (sum) of the 160-double accumulator block [HtH 144 | Htr 12 | res_sq | m | pad 2]."""
from __future__ import annotations

ACC_DOUBLES = 160


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous shard; sizes differ by at most one; every point belongs to exactly one rank."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_accumulators(acc):
    """Sum the accumulator block across ranks in place (torch tensor on the rank's device; NCCL on GPUs, gloo on CPU).
    A no-op when torch.distributed is not initialised (single GPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def unpack_accumulators(acc):
    """-> (HtH 12x12, Htr 12, res_sq, m) from the 160-double block (numpy or torch, on the host)."""
    a = acc.detach().cpu().numpy() if hasattr(acc, "detach") else acc
    return a[:144].reshape(12, 12).copy(), a[144:156].copy(), float(a[156]), int(round(float(a[157])))
