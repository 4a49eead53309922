"""Host-side sharding of independent images over ranks/devices.

The path has no exchange step: image i goes to device/rank i mod G (the same rule
ikc_resize_batch applies inside one process), results never cross devices, and the only
cross-rank traffic is the benchmark's barrier and max-over-ranks timing.  No data-path collective.
"""
from __future__ import annotations


def shard_indices(n_jobs: int, world: int, rank: int) -> list[int]:
    """Indices of the jobs rank `rank` of `world` owns: round robin, job i -> rank i mod world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return list(range(rank, n_jobs, world))


def aggregate_throughput(local_units: float, local_ms: float, dist=None, device=None):
    """(total_units, max_ms, units_per_second): units summed over ranks, time = max over ranks."""
    if dist is None or not dist.is_initialized():
        return local_units, local_ms, local_units / (local_ms * 1e-3)
    import torch
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    u = torch.tensor([local_units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()), float(t.item()), float(u.item()) / (float(t.item()) * 1e-3)
