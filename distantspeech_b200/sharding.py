"""Stream sharding across GPUs.

Streams are independent units -- the reference has no cross-stream state -- so a
job is split into contiguous blocks of streams, one process per GPU, with NO
collective on the data path.  ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
tests) is used only to gather a few output streams for validation and to take the
max of the per-rank timings.
"""
from __future__ import annotations


def shard_bounds(n_streams: int, rank: int, world: int):
    """Contiguous block ``streams[lo:hi]`` owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(n_streams, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gather_validation_streams(local, dst: int = 0):
    """Gather one tensor per rank onto ``dst`` (validation only, outside any timed region).
    Returns the list on ``dst`` and ``None`` elsewhere; with no process group it returns ``[local]``."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    world, rank = dist.get_world_size(), dist.get_rank()
    out = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, out, dst=dst)
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (the job time is the slowest rank's device time)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
