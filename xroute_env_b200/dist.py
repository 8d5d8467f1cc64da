"""Multi-GPU plumbing: one process per GPU, each owning a contiguous shard of the
environment batch; the only collective on this path is a SUM all-reduce of the small
int64 episode/congestion statistics vector (XR_BUF_STATS).  Mirrors how the reference
scales: one independent environment per worker process
(``/root/reference/baseline/A3C/discrete_A3C.py:246-247``)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ._lib import STAT_NAMES


def shard_range(n_envs_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [first, first+count) of the global environment ids owned by `rank`."""
    base, rem = divmod(n_envs_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allreduce_stats(stats: torch.Tensor) -> dict:
    """SUM-all-reduce a stats vector (CUDA tensor under NCCL, CPU tensor under gloo) and
    return it as a dict keyed by STAT_NAMES."""
    t = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(STAT_NAMES, t.cpu().tolist())}
