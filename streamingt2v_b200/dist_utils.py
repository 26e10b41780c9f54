"""torch.distributed plumbing for the one-process-per-GPU launch (bench.py --gpus N).

The denoiser step does not shard in round 1 ("replicas only"): ranks run independent replicas, and the
only cross-rank traffic is the timing reduction.  Backend nccl on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def max_over_ranks(values: Sequence[float], device="cpu") -> List[float]:
    """Element-wise MAX of per-rank timings (a multi-GPU number is the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def shard_items(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced partition of independent work items (replica videos / enhance chunks) over ranks."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def aggregate_throughput(units_per_rank: float, world: int, max_ms: float) -> float:
    """Whole-job units/s = units all ranks processed / slowest rank's time."""
    return world * units_per_rank / (max_ms * 1e-3)
