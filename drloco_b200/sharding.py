"""Multi-GPU plumbing: one process per GPU, environments sharded contiguously, no collective inside step.

The only exchanges are all-reduce(sum) of small packed float64 vectors: the VecNormalize batch moments
every step and the episode-statistics sums at the logging cadence.  ``torch.distributed`` (NCCL on the GPU box, gloo in
the CPU tests) carries them.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the global env indices owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total_envs, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """in-place sum over ranks (no-op for a single process)."""
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def merge_moments(mean, var, count, bsum, bsumsq, bn):
    """RunningMeanStd.update_from_moments from packed sums (same arithmetic as csrc/vecnorm.cu::chan_merge);
    works on torch tensors or floats.  Used by the CPU tests of the N>1 path."""
    bmean = bsum / bn
    bvar = bsumsq / bn - bmean * bmean
    if isinstance(bvar, torch.Tensor):
        bvar = torch.clamp(bvar, min=0.0)
    else:
        import numpy as np
        bvar = np.maximum(bvar, 0.0)
    delta = bmean - mean
    tot = count + bn
    new_mean = mean + delta * bn / tot
    m2 = var * count + bvar * bn + delta * delta * count * bn / tot
    return new_mean, m2 / tot, tot
