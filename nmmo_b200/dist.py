"""Multi-GPU plumbing: environments shard by global index, one process per GPU.

The reference has no distributed code; environments are independent, so
the step needs no collective.  The only exchange is the episode-statistics vector the trainer
logs (means of the finished-agent infos, /root/reference/reinforcement_learning/clean_pufferl.py:381-390):
plain sums and counts, all-reduced (NCCL on GPUs, gloo in CPU tests) and divided afterwards so the
result equals ``np.mean`` over the concatenated per-rank lists.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous env-index range of `rank`: (env_base, n_local).  Remainders go to the low ranks."""
    q, r = divmod(int(total_envs), int(world_size))
    n = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, n


def global_seeds(seed: int, env_base: int, n_local: int) -> np.ndarray:
    """Per-env seeds derived from the GLOBAL env index, so results do not depend on the GPU count."""
    return (np.arange(n_local, dtype=np.uint64) + np.uint64(env_base) + np.uint64(seed)).astype(np.uint64)


def reduce_stats(sums: np.ndarray, counts: np.ndarray, counters: np.ndarray, device=None):
    """All-reduce (sum) of the episode-stat vector across ranks; identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sums, counts, counters
    n1, n2 = len(sums), len(counts)
    t = torch.tensor(np.concatenate([sums, counts, counters.astype(np.float64)]), dtype=torch.float64,
                     device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    v = t.cpu().numpy()
    return v[:n1], v[n1:n1 + n2], v[n1 + n2:].astype(np.uint64)


def stat_means(sums: np.ndarray, counts: np.ndarray) -> np.ndarray:
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(counts > 0, sums / np.maximum(counts, 1), np.nan)
