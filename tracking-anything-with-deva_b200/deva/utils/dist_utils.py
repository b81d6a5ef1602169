"""Clip-parallel multi-GPU helpers (one process per GPU; the reference's only inference-time distribution is
manual ``--start/--count`` video sharding, evaluation/eval_with_detections.py:54-55)."""
from typing import List

import torch
import torch.distributed as dist


def assign_clips(num_clips: int, world_size: int, rank: int, lengths: List[int] = None) -> List[int]:
    """Clip indices owned by ``rank``.  Without lengths: round-robin.  With per-clip frame counts: greedy
    longest-first balancing (deterministic, identical on every rank)."""
    if lengths is None:
        return list(range(rank, num_clips, world_size))
    assert len(lengths) == num_clips
    load = [0] * world_size
    owner = [0] * num_clips
    for i in sorted(range(num_clips), key=lambda j: (-lengths[j], j)):
        r = min(range(world_size), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += lengths[i]
    return [i for i in range(num_clips) if owner[i] == rank]


def max_over_ranks(value: float, device=None) -> float:
    """max of a scalar over all ranks (timing is reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])
