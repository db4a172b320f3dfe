"""Multi-GPU execution of the rollout path: independent trajectories sharded across ranks.

Nothing couples trajectories on the step path (the batch is just the leading array dimension in the
reference, src/data_generation.py:117,153-155), so each rank owns a contiguous block of the ensemble and runs
its own Plan; there is NO collective on the data path.  torch.distributed is used only for the timing barrier
and the max-over-ranks reduction of the elapsed time (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_batch(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, start+count) of `total` trajectories owned by `rank`; sizes differ by <= 1."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad rank/world/total")
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def barrier(device=None) -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value: float, device="cpu") -> float:
    """Slowest rank's value (multi-GPU numbers are always the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank: float, seconds_this_rank: float, device="cpu") -> float:
    """Whole-job throughput = units processed by all ranks / slowest rank's time."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)
