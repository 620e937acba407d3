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


def average_gradients(parameters, world_size: int, group=None) -> None:
    """Data-parallel learner step: one flat all-reduce(SUM) of all gradients, divided by the world size
    (0.47 MB for resnet_b_s -- latency-bound, NVLS / one-shot territory on NVSwitch)."""
    if world_size <= 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world_size)
    offset = 0
    for g in grads:
        g.copy_(flat[offset:offset + g.numel()].view_as(g))
        offset += g.numel()


def global_mean_std(values: torch.Tensor, group=None):
    """Mean and unbiased std of the concatenation of every rank's `values` (three-number all-reduce): the
    advantage normalisation of RolloutBuffer.get_data_loader (rollout_buffer.py:96-99) over the whole batch."""
    v = values.reshape(-1).double()
    moments = torch.stack([v.sum(), (v * v).sum(), torch.tensor(float(v.numel()), dtype=torch.float64, device=v.device)])
    reduce_stats(moments, group=group)
    mean = moments[0] / moments[2]
    var = (moments[1] - moments[2] * mean * mean) / (moments[2] - 1).clamp(min=1)
    return mean.float(), var.clamp(min=0).sqrt().float()


def common_minibatches(local_batches: int, world_size: int, group=None, device=None) -> int:
    """The number of optimiser steps per epoch every rank can join: min over ranks of the local minibatch count
    (ranks with uneven shards would otherwise issue different numbers of gradient all-reduces and hang)."""
    if world_size <= 1:
        return int(local_batches)
    t = torch.tensor([int(local_batches)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return int(t.item())


def broadcast_module(module: torch.nn.Module, optimizer=None, src: int = 0, group=None) -> None:
    """Data-parallel start-up: every replica takes rank `src`'s parameters and buffers (BatchNorm statistics
    included) and, if given, its optimiser's tensor state -- the ranks then stay in lock-step through
    average_gradients without relying on identical seeding."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
        if optimizer is not None:
            for state in optimizer.state.values():
                for v in state.values():
                    if torch.is_tensor(v):
                        dist.broadcast(v, src=src, group=group)


def pin_to_gpu_numa(local_rank: int):
    """Best effort: restrict this process to the CPUs of the NUMA node the GPU hangs off (read from sysfs), so that the
    host side of the per-slab copies and launches of mnk_step_host_loop does not cross sockets when 8 ranks share one
    host.  Returns a short description, or None if nothing was changed."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            raise LookupError("no numa_node")
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            raise LookupError("empty cpu list")
        os.sched_setaffinity(0, allowed)
        return f"numa node {node}: {len(allowed)} cpus"
    except Exception:
        pass
    try:        # containers often hide the PCI device's numa_node: ask the driver for the GPU's ideal CPUs instead
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(torch.cuda.get_device_properties(local_rank).pci_bus_id_str
                                                      if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id_str")
                                                      else f"{torch.cuda.get_device_properties(local_rank).pci_domain_id:08x}:"
                                                           f"{torch.cuda.get_device_properties(local_rank).pci_bus_id:02x}:"
                                                           f"{torch.cuda.get_device_properties(local_rank).pci_device_id:02x}.0")
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed or allowed == os.sched_getaffinity(0):
            return None
        os.sched_setaffinity(0, allowed)
        return f"nvml cpu affinity: {len(allowed)} cpus"
    except Exception:
        return None
