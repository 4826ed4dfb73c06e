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


class GradReducer:
    """Gradient averaging of the fused-minibatch models, overlapped with the rest of backward.

    The default CNN's gradient is one 70 MB tensor (policy_head.weight) plus five small ones.  The big one is the first
    gradient backward produces (the policy head is the last layer), so its all-reduce is started from inside the
    policy head's backward node (``nn_ops.policy_head_evaluate`` calls ``early(tensor)``) as an asynchronous NCCL
    operation and runs over NVLink while the remaining backward kernels (input-gradient GEMM, value head, input-layer
    weight gradient) execute; ``finish`` then reduces the small gradients as ONE flattened buffer and waits for the big
    one before the optimizer reads it.  Everything is issued on / joined into the current stream, so a CUDA-graph
    capture of the update records the collectives and their cross-stream edges like kernels.  With ``gloo`` (CPU tests)
    the same calls run synchronously."""

    def __init__(self):
        self._pending = []   # (work handle, tensor) of reductions started early
        self._early_ids = set()

    def early(self, grad: torch.Tensor) -> torch.Tensor:
        """Start summing ``grad`` (a freshly computed, contiguous gradient tensor) over the ranks; returns it."""
        if world()[1] > 1:
            work = dist.all_reduce(grad, op=dist.ReduceOp.SUM, async_op=True)
            self._pending.append(work)
            self._early_ids.add(grad.data_ptr())
        return grad

    def finish(self, params) -> None:
        """Reduce every gradient that ``early`` did not take (flattened into one buffer) and wait for the early ones."""
        if world()[1] <= 1:
            return
        rest = [p.grad for p in params if p.grad is not None and p.grad.data_ptr() not in self._early_ids]
        if rest:
            flat = torch.cat([g.reshape(-1) for g in rest])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            off = 0
            for g in rest:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        for work in self._pending:
            work.wait()  # the current stream waits for the NCCL stream
        self._pending.clear()
        self._early_ids.clear()


def all_reduce_grads(params) -> None:
    """Sum the gradients over ranks, one all-reduce per parameter (NCCL over NVLink on GPUs).  Issued on the
    current stream, so a CUDA-graph capture of the update records them like kernels."""
    if world()[1] > 1:
        for p in params:
            if p.grad is not None:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
