"""torch.distributed plumbing for the one-process-per-GPU launch (bench.py --gpus N).

The throughput benchmark runs independent replicas ("replicas only"): the only cross-rank
traffic is the timing reduction.  Two opt-in latency modes use the partitions SURVEY.md section 8(e) names — the
classifier-free-guidance halves of a denoise step (2 ranks, one all-gather of the network output per step) and the
groups of <= 8 frames of the VAE decode — through `all_gather_cat`.  Backend nccl on GPUs, gloo in the CPU tests.
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


def all_gather_cat(t: torch.Tensor) -> torch.Tensor:
    """Concatenate every rank's (equal-shape) tensor along dim 0 in rank order; identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, 0)


def broadcast_from_rank0(t: torch.Tensor) -> torch.Tensor:
    """In-place broadcast of rank 0's tensor to every rank (identity without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = t.contiguous()
        dist.broadcast(t, src=0)
    return t


def broadcast_object(obj):
    """Rank 0's picklable object on every rank (identity without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    return obj


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_items(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced partition of independent work items (replica videos / enhance chunks) over ranks."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def aggregate_throughput(units_per_rank: float, world: int, max_ms: float) -> float:
    """Whole-job units/s = units all ranks processed / slowest rank's time."""
    return world * units_per_rank / (max_ms * 1e-3)
