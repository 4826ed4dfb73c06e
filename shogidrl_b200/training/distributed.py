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

    The default CNN's gradient is one big tensor (policy_head.weight: 17.5 M of the 17.55 M parameters) plus five small
    ones.  The big one is the first gradient backward produces (the policy head is the last layer) and it leaves its GEMM
    in bf16, so ``early`` -- called from inside the policy head's backward node (``nn_ops.policy_head_evaluate``) --
    forks a side stream that all-reduces those 35 MB of bf16 over NCCL / NVLink and then widens the sum into the
    parameter's fp32 ``.grad``, while the main stream goes on with the input-gradient GEMM, the value head and the input
    layer's weight gradient.  ``finish`` reduces the small gradients as ONE flattened fp32 buffer and joins the side
    stream before the optimizer reads anything.  All work is enqueued on torch streams (fork / join by events), so a
    CUDA-graph capture of the update records the collectives and the cross-stream edges like kernels.  With ``gloo`` (CPU
    tests) the same calls run synchronously.  ``KZ_GRAD_REDUCE_DTYPE=fp32`` widens before the all-reduce instead
    (twice the bytes on the wire, no bf16 rounding of the cross-rank sum)."""

    def __init__(self):
        import os
        self._events = []
        self._early = set()   # id() of the parameters whose gradient `early` has taken care of
        self._stream = None
        self._wide = os.environ.get("KZ_GRAD_REDUCE_DTYPE", "bf16").lower() in ("fp32", "float32")

    def early(self, param: torch.nn.Parameter, grad_lowp: torch.Tensor) -> None:
        """``param.grad`` := sum over ranks of ``grad_lowp`` (a freshly computed contiguous gradient, any float dtype),
        computed off the main stream; ``param.grad`` must be unset (zero_grad(set_to_none=True))."""
        assert param.grad is None, "GradReducer.early expects gradients cleared with set_to_none=True"
        out = torch.empty(grad_lowp.shape, dtype=param.dtype, device=grad_lowp.device)
        if grad_lowp.is_cuda:
            cur = torch.cuda.current_stream(grad_lowp.device)
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=grad_lowp.device)
            side = self._stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                if self._wide:
                    out.copy_(grad_lowp)
                    dist.all_reduce(out, op=dist.ReduceOp.SUM)
                else:
                    dist.all_reduce(grad_lowp, op=dist.ReduceOp.SUM)
                    out.copy_(grad_lowp)
                ev = torch.cuda.Event()
                ev.record(side)
            out.record_stream(side)
            grad_lowp.record_stream(side)
            self._events.append(ev)
        else:
            dist.all_reduce(grad_lowp, op=dist.ReduceOp.SUM)
            out.copy_(grad_lowp)
        with torch.no_grad():
            param.grad = out
        self._early.add(id(param))

    def finish(self, params) -> None:
        """Reduce every gradient that ``early`` did not take (flattened into one buffer) and join the early ones."""
        if world()[1] <= 1:
            return
        rest = [p.grad for p in params if p.grad is not None and id(p) not in self._early]
        if rest:
            flat = torch.cat([g.reshape(-1) for g in rest])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            off = 0
            for g in rest:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        for ev in self._events:
            torch.cuda.current_stream().wait_event(ev)
        self._events.clear()
        self._early.clear()


def all_reduce_grads(params) -> None:
    """Sum the gradients over ranks, one all-reduce per parameter (NCCL over NVLink on GPUs).  Issued on the
    current stream, so a CUDA-graph capture of the update records them like kernels."""
    if world()[1] > 1:
        for p in params:
            if p.grad is not None:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
