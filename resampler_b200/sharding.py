"""Multi-GPU plumbing: streams are independent (no shared mutable state), so the
path shards by stream with NO data-path collective.  One process per GPU; torch.distributed is
used only for the barrier and for combining per-rank timings / counts."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous range of stream ids owned by `rank` (s -> s * G / N): [lo, hi)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    lo = n_total * rank // world_size
    hi = n_total * (rank + 1) // world_size
    return lo, hi


def all_reduce_scalar(dist, value: float, op: str = "max", device=None) -> float:
    """max / sum of a python float over all ranks (identity when dist is None)."""
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def job_throughput(dist, produced_local: int, seconds_local: float, device=None) -> float:
    """Whole-job samples/s: units of all ranks / slowest rank's time."""
    total = all_reduce_scalar(dist, float(produced_local), "sum", device)
    t_max = all_reduce_scalar(dist, float(seconds_local), "max", device)
    return total / t_max if t_max > 0 else 0.0
