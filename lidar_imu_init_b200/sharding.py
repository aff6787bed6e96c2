"""Scan sharding for frames of >= 1M points, host-side helpers.

The multi-GPU mechanics live behind the C-ABI (include/liinit_gpu.h: liinit_comm_unique_id / liinit_comm_init): the map is
replicated on every GPU, every rank uploads the same frame, the library cuts it into equal slots and all-reduces the 160-double
accumulator block [HtH 144 | Htr 12 | res_sq | m | pad 2] inside liinit_icp_iterate. What is left for the host application is
to hand the 128-byte NCCL id from one rank to the others -- over whatever channel it has; `attach_comm` does it over
torch.distributed (gloo or nccl), which is how bench.py and the tests launch their ranks."""
from __future__ import annotations

ACC_DOUBLES = 160


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of rank's slot, the rule of the library (liinit_gpu.cu set_scan): equal slots of ceil(n / world) points, the last
    ones clipped to n -- equal slots are what lets the per-point results be all-gathered in place."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard arguments")
    cnt = -(-n // world)
    lo = min(rank * cnt, n)
    return lo, min(lo + cnt, n)


def attach_comm(g, rank: int, world: int):
    """Collective over the ranks of an initialised torch.distributed group: rank 0 draws the NCCL id through the library,
    everyone receives it and attaches its context (liinit_comm_init). No-op for world == 1."""
    if world <= 1:
        return
    import torch.distributed as dist
    box = [g.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    g.comm_init(box[0], world, rank)


def allreduce_accumulators(acc):
    """Sum an accumulator block across the ranks of torch.distributed in place (gloo on CPU; used by the CPU tests of the
    sharding rule, where the oracle stands in for the device pass). A no-op when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def unpack_accumulators(acc):
    """-> (HtH 12x12, Htr 12, res_sq, m) from the 160-double block (numpy or torch, on the host)."""
    a = acc.detach().cpu().numpy() if hasattr(acc, "detach") else acc
    return a[:144].reshape(12, 12).copy(), a[144:156].copy(), float(a[156]), int(round(float(a[157])))
