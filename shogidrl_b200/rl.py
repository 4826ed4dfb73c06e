"""Device-side agent/buffer primitives: masked categorical sampling and GAE (C ABI: kz_sample_masked, kz_gae)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _native as nv


def sample_masked(logits: torch.Tensor, mask: torch.Tensor, seed: int = 0, offset: int = 0,
                  deterministic: bool = False, want_entropy: bool = False
                  ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """Masked softmax -> Categorical sample (argmax if deterministic) -> log_prob, one warp per row.

    Mirrors BaseActorCriticModel.get_action_and_value after forward() (base_actor_critic.py:64-116):
    illegal logits -> -inf, softmax, NaN rows -> uniform, Categorical(probs) with its eps clamp."""
    dev = nv.require_cuda(logits.device)
    assert logits.dim() == 2 and logits.shape[1] == nv.NUM_ACTIONS and logits.stride(1) == 1
    assert logits.dtype in (torch.float32, torch.bfloat16)
    assert mask.shape == logits.shape and mask.stride(1) == 1 and mask.dtype in (torch.uint8, torch.bool)
    n = logits.shape[0]
    actions = torch.empty(n, dtype=torch.int64, device=dev)
    logp = torch.empty(n, dtype=torch.float32, device=dev)
    ent = torch.empty(n, dtype=torch.float32, device=dev) if want_entropy else None
    nv.check(nv.lib().kz_sample_masked(logits.data_ptr(), int(logits.dtype == torch.bfloat16), logits.stride(0),
                                       mask.data_ptr(), mask.stride(0), n, int(seed), int(offset), actions.data_ptr(), 1,
                                       logp.data_ptr(), nv.ptr(ent), int(deterministic), nv.stream_ptr(dev)),
             "kz_sample_masked")
    return actions, logp, ent


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
        nv.check(nv.lib().kz_eval_masked_fwd(logits.data_ptr(), int(logits.dtype == torch.bfloat16), logits.stride(0),
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
        nv.check(nv.lib().kz_eval_masked_bwd(logits.data_ptr(), int(logits.dtype == torch.bfloat16), logits.stride(0),
                                             mask.data_ptr(), mask.stride(0), nv.ptr(ctx.mask_rows), actions.data_ptr(), n,
                                             dlogp.contiguous().float().data_ptr(), dent.contiguous().float().data_ptr(),
                                             saved.data_ptr(), store.data_ptr(), ldg, nv.stream_ptr(dev)),
                 "kz_eval_masked_bwd")
        return store[:, : nv.NUM_ACTIONS], None, None, None


def evaluate_masked(logits: torch.Tensor, mask: torch.Tensor, actions: torch.Tensor,
                    mask_rows: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(log_prob of ``actions``, entropy) of the masked softmax over ``logits`` [B, 13527], differentiable w.r.t.
    the logits.  ``mask`` is [B, 13527] (bool/uint8, any row stride) or, with ``mask_rows`` (int64 [B]), the whole
    rollout mask storage indexed per row -- no minibatch gather of masks."""
    assert logits.dim() == 2 and logits.shape[1] == nv.NUM_ACTIONS and logits.stride(1) == 1
    assert logits.dtype in (torch.float32, torch.bfloat16)
    assert mask.stride(-1) == 1 and mask.dtype in (torch.uint8, torch.bool) and mask.dim() == 2
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
        assert all(t.shape[0] == n for t in args)
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
