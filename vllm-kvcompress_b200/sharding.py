"""Request sharding for the multi-GPU layout: the path has no exchange
step, so GPU g owns requests {i : i mod G == g}, its own cache, tables and metrics.  The only
collectives are a start-up broadcast (stands in for the one-time weight broadcast) and the
max-over-ranks reduction of the timed region.  Backend-agnostic (nccl on GPUs, gloo in tests)."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_requests(num_requests: int, world_size: int, rank: int) -> List[int]:
    """Round-robin request ids owned by `rank`."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, num_requests, world_size))


def broadcast_config(values: List[int], device) -> List[int]:
    """Rank 0's workload header to every rank (one small broadcast at start-up)."""
    t = torch.tensor(values, dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, 0)
    return [int(v) for v in t.tolist()]


def max_over_ranks(values: List[float], device) -> List[float]:
    """Whole-job time of a step = the slowest rank's device time."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def aggregate_throughput(tokens_per_rank: int, ms_per_step: float) -> float:
    """tokens/s of the whole job: all ranks' tokens over the max-over-ranks step time."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    return tokens_per_rank * world / (ms_per_step * 1e-3)
