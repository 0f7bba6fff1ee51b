"""Independent-shard partitioning of a profile batch over ranks / GPUs.

Profiles never communicate and the 4 chains of a profile stay together, so the multi-GPU path is a
contiguous split with NO data-path collective; results are gathered on the host.  The only communication a
multi-rank run performs is the barrier / max-over-ranks timing of bench.py, expressed here over
torch.distributed so it can be tested with the gloo backend on CPU.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [first, last) profile range of `rank`; the same rule foct_sample() uses for devices[]."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def aggregate(stats_max: np.ndarray, stats_sum: np.ndarray, dist=None):
    """max-over-ranks of `stats_max` (times) and sum-over-ranks of `stats_sum` (units processed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(stats_max, dtype=np.float64), np.asarray(stats_sum, dtype=np.float64)
    import torch

    mx = torch.tensor(np.asarray(stats_max, dtype=np.float64))
    sm = torch.tensor(np.asarray(stats_sum, dtype=np.float64))
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return mx.numpy(), sm.numpy()


def gather_rows(local: np.ndarray, dist=None) -> np.ndarray:
    """Host-side gather of per-profile result rows (rank order = profile order)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    return np.concatenate(out, axis=0)
