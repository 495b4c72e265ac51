"""Replica sharding helpers (host logic shared by bench.py and the tests).

The path shards trivially over images (frozen BatchNorm statistics: no cross-image state), so
multi-GPU = independent replicas, contiguous index ranges, no data-path collective.  The only
communication is the barrier + MAX reduction of the measured time.
"""
from __future__ import annotations


def split_contiguous(n: int, g: int):
    """[(begin, end)] of the g contiguous, as-even-as-possible shards of range(n) — same rule as
    libroomnet's replica scheduler (roomnet_b200/csrc/api.cpp, Infer)."""
    out, b = [], 0
    for r in range(g):
        m = n // g + (1 if r < n % g else 0)
        out.append((b, b + m))
        b += m
    return out


def reduce_max(value: float, device=None) -> float:
    """MAX over ranks of a scalar (per-rank elapsed time); identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank: int, world: int, steps: int, max_seconds: float) -> float:
    """Whole-job throughput: units all ranks processed / the slowest rank's time."""
    return world * units_per_rank * steps / max_seconds
