"""Multi-GPU plumbing: one process per GPU, environments sharded by rank with NO rollout-time collective; the
only exchanges are the DDP gradient all-reduce inside ``loss.backward()`` and three scalars for whole-buffer
advantage normalisation.  Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU (tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_envs(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(env_offset, n_local): contiguous, balanced shards; the offset keys the per-env RNG stream so that the
    union over ranks equals a single-process run over ``total_envs`` environments."""
    base, rem = divmod(int(total_envs), int(world_size))
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, n_local


def global_moments(x: torch.Tensor) -> Tuple[float, float, float]:
    """(count, sum, sum of squares) of x over all ranks."""
    t = torch.stack([torch.tensor(float(x.numel()), device=x.device, dtype=torch.float64), x.double().sum(),
                     (x.double() ** 2).sum()])
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    c, s1, s2 = t.tolist()
    return c, s1, s2


def reduce_max(value: float, device) -> float:
    """Max over ranks (multi-GPU timings are reported as the slowest rank)."""
    t = torch.tensor([value], device=device, dtype=torch.float64)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, device) -> float:
    t = torch.tensor([value], device=device, dtype=torch.float64)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def wrap_ddp(model: torch.nn.Module, device: torch.device) -> torch.nn.Module:
    """DistributedDataParallel when a process group exists (gradient all-reduce over NCCL / NVLink), else the
    model itself.  The reference declares a ``ddp`` flag (config_schema.py:81) but never wires it."""
    if world()[1] > 1:
        ids = [device.index] if device.type == "cuda" else None
        return torch.nn.parallel.DistributedDataParallel(model, device_ids=ids)
    return model


def broadcast_module(module: torch.nn.Module, src: int = 0) -> None:
    """Every rank starts from rank ``src``'s parameters and buffers (what the DDP constructor does)."""
    if world()[1] > 1:
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t, src)


def all_reduce_grads(params) -> None:
    """Sum the gradients over ranks, one all-reduce per parameter (NCCL over NVLink on GPUs).  Issued on the
    current stream, so a CUDA-graph capture of the update records them like kernels."""
    if world()[1] > 1:
        for p in params:
            if p.grad is not None:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
