"""Replica sharding helpers for the multi-GPU run (pairs are independent, no collective on the data path).

One process per GPU; every rank holds the same weights and its own slice of the pairs.  The only communication is
bookkeeping AFTER the timed region: a barrier and a MAX-reduction of the per-rank device time."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of `n_items` pairs owned by `rank` (first n_items % world ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """MAX-reduce a host scalar over the default process group (identity when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(items_this_rank: int, ms_this_rank: float, device=None) -> float:
    """Whole-job items/s = all ranks' items / max over ranks of the device time."""
    total = sum_over_ranks(items_this_rank, device)
    worst = max_over_ranks(ms_this_rank, device)
    return total / (worst / 1e3)
