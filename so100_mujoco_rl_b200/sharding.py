"""Env-index sharding across GPUs: one process per GPU, no collective on the step path.

Environments are independent, so rank r of W owns the contiguous global env ids [lo, hi) and passes `lo` as
`env_offset`; the device RNG is keyed by the GLOBAL env id, which makes every trajectory independent of W
(SURVEY.md §8 e).  `torch.distributed` is used only to agree on timing (barrier + max over ranks) in bench.py.
"""
from __future__ import annotations

import os


def shard_range(total_envs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition: sizes differ by at most one and cover [0, total_envs) exactly once."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    if total_envs < 0:
        raise ValueError("total_envs must be non-negative")
    lo = (total_envs * rank) // world
    hi = (total_envs * (rank + 1)) // world
    return lo, hi


def dist_env() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment, defaulting to a single process."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
