"""Env-batch sharding over the GPUs of one node (one process per GPU, torchrun).

Envs are independent (no cross-env term anywhere on the path), so the global batch is cut into
contiguous equal shards: rank r owns global envs [offset, offset + count).  Counter-based RNG
streams are keyed by GLOBAL env id, so results do not depend on the number of GPUs.  There is no
per-step collective; the only communication is one all-reduce (NCCL on GPUs, gloo in CPU tests) of
a small statistics vector per rollout.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(count, offset) of rank's contiguous shard; the first `total % world` ranks get one extra env."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return count, offset


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun) and initialises the process group.
    Returns (rank, world_size, local_rank); a single process needs no group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local_rank


def reduce_stats(stats: torch.Tensor, op=dist.ReduceOp.SUM, group=None) -> torch.Tensor:
    """In-place all-reduce of a rollout statistics vector; identity for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=op, group=group)
    return stats
