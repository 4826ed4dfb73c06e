"""Device-side agent/buffer primitives: masked categorical sampling, GAE, masked evaluation, the PPO loss and the
clip + Adam tail of an update (C ABI: kz_sample_masked, kz_gae, kz_eval_masked_*, kz_ppo_loss, kz_adam_clip_step)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _native as nv


def is_bitmap(mask: torch.Tensor) -> bool:
    """A legal set given as the engine's bitmap rows (int32 [..., 448], bit i of a row = action i legal; what
    ``VecShogiEnv.step_rollout`` writes) rather than as PolicyOutputMapper.get_legal_mask's byte rows."""
    return mask.dtype == torch.int32 and mask.shape[-1] == nv.BITMAP_WORDS


def _check_legal_rows(mask: torch.Tensor, what: str) -> None:
    if is_bitmap(mask):
        nv.require(mask.dim() == 2 and mask.stride(1) == 1 and mask.stride(0) >= nv.BITMAP_WORDS and mask.data_ptr() % 4 == 0,
                   f"{what}: bitmap rows must be contiguous int32 [B, {nv.BITMAP_WORDS}]")
    else:
        nv.require(mask.dim() == 2 and mask.shape[1] == nv.NUM_ACTIONS and mask.stride(1) == 1
                   and mask.dtype in (torch.uint8, torch.bool),
                   f"{what}: expected bool/uint8 [B, {nv.NUM_ACTIONS}] rows or an int32 [B, {nv.BITMAP_WORDS}] legal bitmap")


def sample_masked(logits: torch.Tensor, mask: torch.Tensor, seed: int = 0, offset: int = 0,
                  deterministic: bool = False, want_entropy: bool = False,
                  out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, offset_tensor: Optional[torch.Tensor] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """Masked softmax -> Categorical sample (argmax if deterministic) -> log_prob, one warp per row.

    Mirrors BaseActorCriticModel.get_action_and_value after forward() (base_actor_critic.py:64-116):
    illegal logits -> -inf, softmax, NaN rows -> uniform, Categorical(probs) with its eps clamp.  ``mask`` is the
    byte mask [n, 13527] or the engine's legal bitmap [n, 448] int32 (kz_sample_bitmap: identical results, 7.5x fewer
    mask bytes).  ``out`` = (actions int64 [n], log_probs fp32 [n]) writes straight into rollout storage.
    ``offset_tensor`` (int64 [1] on the device) is added to ``offset`` by the kernel when it runs: a draw counter that a
    CUDA-graph replay can advance."""
    dev = nv.require_cuda(logits.device)
    nv.require(logits.dim() == 2 and logits.shape[1] == nv.NUM_ACTIONS and logits.stride(1) == 1, "logits.dim() == 2 and logits.shape[1] == nv.NUM_ACTIONS and logits.stride(1) == 1")
    nv.require(logits.dtype in (torch.float32, torch.bfloat16), "logits.dtype in (torch.float32, torch.bfloat16)")
    _check_legal_rows(mask, "mask")
    n = logits.shape[0]
    nv.require(mask.shape[0] == n, "mask.shape[0] == logits.shape[0]")
    if out is not None:
        actions, logp = out
        nv.require(actions.dtype == torch.int64 and logp.dtype == torch.float32 and actions.is_contiguous()
                   and logp.is_contiguous() and actions.numel() >= n and logp.numel() >= n and actions.device == dev
                   and logp.device == dev, "out: (int64 [n], float32 [n]) contiguous tensors on the logits' device")
    else:
        actions = torch.empty(n, dtype=torch.int64, device=dev)
        logp = torch.empty(n, dtype=torch.float32, device=dev)
    ent = torch.empty(n, dtype=torch.float32, device=dev) if want_entropy else None
    if offset_tensor is not None:
        nv.require(offset_tensor.dtype == torch.int64 and offset_tensor.numel() >= 1 and offset_tensor.device == dev,
                   "offset_tensor: int64 [1] on the logits' device")
    fn = nv.lib().kz_sample_bitmap if is_bitmap(mask) else nv.lib().kz_sample_masked
    nv.check(fn(logits.data_ptr(), int(logits.dtype == torch.bfloat16), logits.stride(0),
                mask.data_ptr(), mask.stride(0), n, int(seed), int(offset), nv.ptr(offset_tensor), actions.data_ptr(), 1,
                logp.data_ptr(), nv.ptr(ent), int(deterministic), nv.stream_ptr(dev)),
             "kz_sample_masked")
    return actions, logp, ent


def bitmap_to_mask(bitmap: torch.Tensor, rows: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None
                   ) -> torch.Tensor:
    """bool [B, 13527] legal masks (PolicyOutputMapper.get_legal_mask's layout, utils.py:310-336) from legal bitmap
    rows ``bitmap[rows]`` (or every row): the view API callers of ``ExperienceBuffer.legal_masks`` expect."""
    dev = nv.require_cuda(bitmap.device)
    _check_legal_rows(bitmap, "bitmap")
    nv.require(is_bitmap(bitmap), "bitmap: int32 [B, 448]")
    if rows is not None:
        rows = rows.to(device=dev, dtype=torch.int64).contiguous()
    n = bitmap.shape[0] if rows is None else rows.numel()
    if out is None:
        out = torch.empty((n, nv.NUM_ACTIONS), dtype=torch.bool, device=dev)
    nv.require(out.dim() == 2 and out.shape[0] >= n and out.shape[1] == nv.NUM_ACTIONS and out.stride(1) == 1
               and out.dtype in (torch.bool, torch.uint8) and out.device == dev, "out: bool/uint8 [>= n, 13527] rows")
    if n:
        nv.check(nv.lib().kz_bitmap_expand(bitmap.data_ptr(), bitmap.stride(0), nv.ptr(rows), n, out.data_ptr(), out.stride(0),
                                           nv.stream_ptr(dev)), "kz_bitmap_expand")
    return out


def gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, last_value: torch.Tensor,
        gamma: float, lambda_gae: float, exact: bool = False,
        out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """[T, N] reverse-time GAE (experience_buffer.py:99-145).  gamma*lambda is multiplied in double first and
    both factors are rounded to fp32 exactly as the reference's Python-float x tensor products are."""
    dev = nv.require_cuda(rewards.device)
    T, N = rewards.shape
    r = rewards.contiguous().float()
    v = values.contiguous().float()
    d = dones.contiguous()
    d = d.view(torch.uint8) if d.dtype == torch.bool else d.to(torch.uint8)
    lv = last_value.to(device=dev, dtype=torch.float32).reshape(-1).expand(N).contiguous() if last_value.numel() == 1 \
        else last_value.to(device=dev, dtype=torch.float32).contiguous()
    adv, ret = out if out is not None else (torch.empty_like(r), torch.empty_like(r))
    fn = nv.lib().kz_gae_exact if exact else nv.lib().kz_gae
    nv.check(fn(r.data_ptr(), v.data_ptr(), d.data_ptr(), lv.data_ptr(), T, N, float(gamma),
                float(gamma * lambda_gae), adv.data_ptr(), ret.data_ptr(), nv.stream_ptr(dev)), "kz_gae")
    return adv, ret


class _MaskedCategoricalEval(torch.autograd.Function):
    """log-prob of taken actions + entropy of the masked softmax, fused forward/backward (kz_eval_masked_*)."""

    @staticmethod
    def forward(ctx, logits, mask, actions, mask_rows):
        dev = nv.require_cuda(logits.device)
        n = logits.shape[0]
        logp = torch.empty(n, dtype=torch.float32, device=dev)
        ent = torch.empty(n, dtype=torch.float32, device=dev)
        saved = torch.empty((n, 4), dtype=torch.float32, device=dev)
        actions = actions.contiguous().long()
        fn = nv.lib().kz_eval_bitmap_fwd if is_bitmap(mask) else nv.lib().kz_eval_masked_fwd
        nv.check(fn(logits.data_ptr(), int(logits.dtype == torch.bfloat16), logits.stride(0),
                    mask.data_ptr(), mask.stride(0), nv.ptr(mask_rows), actions.data_ptr(), n,
                    logp.data_ptr(), ent.data_ptr(), saved.data_ptr(), nv.stream_ptr(dev)),
                 "kz_eval_masked_fwd")
        ctx.save_for_backward(logits, mask, actions, saved)
        ctx.mask_rows = mask_rows
        return logp, ent

    @staticmethod
    def backward(ctx, dlogp, dent):
        logits, mask, actions, saved = ctx.saved_tensors
        dev = logits.device
        n = logits.shape[0]
        ldg = (nv.NUM_ACTIONS + 15) // 16 * 16
        store = torch.empty((n, ldg), dtype=logits.dtype, device=dev)  # 16-byte aligned rows; the kernel clears [0, A)
        store[:, nv.NUM_ACTIONS:].zero_()
        args = (logits.data_ptr(), int(logits.dtype == torch.bfloat16), logits.stride(0),
                mask.data_ptr(), mask.stride(0), nv.ptr(ctx.mask_rows), actions.data_ptr(), n,
                dlogp.contiguous().float().data_ptr(), dent.contiguous().float().data_ptr(),
                saved.data_ptr(), store.data_ptr(), ldg)
        if is_bitmap(mask):
            nv.check(nv.lib().kz_eval_bitmap_bwd(*args, None, nv.stream_ptr(dev)), "kz_eval_bitmap_bwd")
        else:
            nv.check(nv.lib().kz_eval_masked_bwd(*args, nv.stream_ptr(dev)), "kz_eval_masked_bwd")
        return store[:, : nv.NUM_ACTIONS], None, None, None


def evaluate_masked(logits: torch.Tensor, mask: torch.Tensor, actions: torch.Tensor,
                    mask_rows: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(log_prob of ``actions``, entropy) of the masked softmax over ``logits`` [B, 13527], differentiable w.r.t.
    the logits.  ``mask`` is [B, 13527] (bool/uint8, any row stride) -- or the engine's legal bitmap rows, int32
    [B, 448] -- or, with ``mask_rows`` (int64 [B]), the whole rollout mask / bitmap storage indexed per row: no
    minibatch gather of masks."""
    nv.require(logits.dim() == 2 and logits.shape[1] == nv.NUM_ACTIONS and logits.stride(1) == 1, "logits.dim() == 2 and logits.shape[1] == nv.NUM_ACTIONS and logits.stride(1) == 1")
    nv.require(logits.dtype in (torch.float32, torch.bfloat16), "logits.dtype in (torch.float32, torch.bfloat16)")
    _check_legal_rows(mask, "mask")
    if mask_rows is not None:
        mask_rows = mask_rows.contiguous().long()
    return _MaskedCategoricalEval.apply(logits, mask, actions, mask_rows)


class _PPOLoss(torch.autograd.Function):
    """Clipped-surrogate loss + metrics + closed-form gradients in one launch (kz_ppo_loss)."""

    @staticmethod
    def forward(ctx, new_lp, entropy, new_v, old_lp, adv, ret, clip_eps, value_coef, entropy_coef, grad_scale):
        dev = nv.require_cuda(new_lp.device)
        n = new_lp.shape[0]
        args = [t.detach().contiguous().float().reshape(-1) for t in (new_lp, entropy, new_v, old_lp, adv, ret)]
        nv.require(all(t.shape[0] == n for t in args), "all(t.shape[0] == n for t in args)")
        out = torch.empty(6, dtype=torch.float32, device=dev)
        grads = torch.empty((3, n), dtype=torch.float32, device=dev)
        nv.check(nv.lib().kz_ppo_loss(*[t.data_ptr() for t in args], n, float(clip_eps), float(value_coef),
                                      float(entropy_coef), float(grad_scale), out.data_ptr(), grads[0].data_ptr(),
                                      grads[1].data_ptr(), grads[2].data_ptr(), nv.stream_ptr(dev)), "kz_ppo_loss")
        ctx.save_for_backward(grads)
        ctx.shapes = (new_lp.shape, entropy.shape, new_v.shape)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        (grads,) = ctx.saved_tensors
        g = grads * g_loss
        s = ctx.shapes
        return g[0].reshape(s[0]), g[1].reshape(s[1]), g[2].reshape(s[2]), None, None, None, None, None, None, None


def ppo_loss(new_lp: torch.Tensor, entropy: torch.Tensor, new_v: torch.Tensor, old_lp: torch.Tensor, adv: torch.Tensor,
             ret: torch.Tensor, clip_eps: float, value_coef: float, entropy_coef: float, grad_scale: float = 1.0):
    """-> (loss, stats6): loss = policy + value_coef * value + entropy_coef * entropy_term as a differentiable
    scalar (gradients w.r.t. new_lp / entropy / new_v, multiplied by ``grad_scale``); stats6 = [loss, policy loss,
    value loss, entropy term, mean(old_lp - new_lp), clip fraction] (ppo_agent.py:332-372, value clipping off)."""
    return _PPOLoss.apply(new_lp, entropy, new_v, old_lp, adv, ret, clip_eps, value_coef, entropy_coef, grad_scale)


def adam_clip_applicable(optimizer: torch.optim.Optimizer) -> bool:
    """The fused tail stands in for ``clip_grad_norm_`` + ``optimizer.step()`` exactly when the optimizer is a plain
    torch.optim.Adam (one parameter group, no amsgrad / maximize, float lr) over dense fp32 CUDA parameters."""
    if type(optimizer) is not torch.optim.Adam or len(optimizer.param_groups) != 1:
        return False
    g = optimizer.param_groups[0]
    if g.get("amsgrad") or g.get("maximize") or g.get("differentiable") or isinstance(g["lr"], torch.Tensor):
        return False
    return all(p.is_cuda and p.dtype == torch.float32 and _dense(p) and not p.is_sparse for p in g["params"])


def _dense(t: torch.Tensor) -> bool:
    """Non-overlapping and dense in either memory format: the update is elementwise over the storage, so any such layout
    works as long as parameter, gradient and moments share it (channels_last convolution weights of the ResNet tower)."""
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def adam_clip_step(optimizer: torch.optim.Adam, max_norm: float, norm_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``clip_grad_norm_(params, max_norm)`` + ``optimizer.step()`` in three launches (kz_adam_clip_step): returns the
    total gradient norm before clipping as a device scalar.  Works on the optimizer's own state tensors (``step``,
    ``exp_avg``, ``exp_avg_sq``, created here the way torch.optim.Adam creates them when capturable), so
    ``optimizer.state_dict()`` / checkpoints stay what the reference writes (ppo_agent.py:462-487).  No host
    synchronisation: safe under CUDA-graph capture once the state exists."""
    group = optimizer.param_groups[0]
    params = [p for p in group["params"] if p.grad is not None]
    if not params:
        return torch.zeros((), device=group["params"][0].device)
    dev = nv.require_cuda(params[0].device)
    steps, grads = [], []
    for p in params:
        st = optimizer.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=dev)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        if not (torch.is_tensor(st["step"]) and st["step"].is_cuda and st["step"].dtype == torch.float32):
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32, device=dev)  # loaded non-capturable state
        steps.append(st["step"])
        g = p.grad
        if g.dtype != torch.float32 or g.stride() != p.stride():
            g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)  # the parameter's layout, fp32
        grads.append(g)
        for k in ("exp_avg", "exp_avg_sq"):
            if st[k].stride() != p.stride():  # e.g. state loaded from a checkpoint written in the other layout
                st[k] = torch.empty_like(p, memory_format=torch.preserve_format).copy_(st[k])
    torch._foreach_add_(steps, 1.0)
    n = len(params)
    numel = (C.c_int64 * n)(*[p.numel() for p in params])
    L = nv.lib()
    need = int(L.kz_adam_clip_workspace(n, numel))
    ws = torch.empty(max(need, 1), dtype=torch.float32, device=dev)
    out = norm_out if norm_out is not None else torch.empty(2, dtype=torch.float32, device=dev)
    nv.require(out.numel() >= 2 and out.dtype == torch.float32 and out.is_contiguous(), "out.numel() >= 2 and out.dtype == torch.float32 and out.is_contiguous()")
    arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
    b1, b2 = group["betas"]
    nv.check(L.kz_adam_clip_step(n, arr(params), arr(grads), arr([optimizer.state[p]["exp_avg"] for p in params]),
                                 arr([optimizer.state[p]["exp_avg_sq"] for p in params]), arr(steps), numel,
                                 float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                 float(group["weight_decay"]), float(max_norm), ws.data_ptr(), ws.numel(), out.data_ptr(),
                                 nv.stream_ptr(dev)), "kz_adam_clip_step")
    optimizer._opt_called = True  # what LRScheduler.step() looks at before warning about the call order
    return out[0]
