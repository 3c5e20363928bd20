"""Multi-GPU plumbing: scenes are independent units of the forward path, so N GPUs
= N processes that each own a disjoint shard of the scene stream; there is NO data-path collective.
The only communication is the throughput aggregation (max of the per-rank device times)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of scene indices to ranks (scene i -> rank i % world)."""
    return list(range(rank, n_items, world))


def aggregate_times(times: Sequence[float], device=None) -> List[float]:
    """Element-wise MAX over ranks of per-rank elapsed times (works on gloo/CPU and nccl/GPU)."""
    t = torch.tensor(list(times), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def whole_job_throughput(units_per_rank: int, world: int, elapsed_max: float) -> float:
    """Units processed by ALL ranks divided by the slowest rank's time."""
    return world * units_per_rank / elapsed_max
