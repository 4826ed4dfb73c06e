"""Network-side operators around the rollout path that run as hand-written kernels (csrc/kz_nn.cu, csrc/kz_rl.cu).

``obs_conv`` is the default model's input layer -- nn.Conv2d(46, 16, 3, padding=1) [+ ReLU] under bf16 autocast
(keisei/core/neural_network.py:14-28, keisei/core/ppo_agent.py:323) -- reading the fp32 observations straight
from the rollout storage (optionally through minibatch row indices): no gather, no cast pass, forward and weight
gradient each one pass over the observations.  The input gets no gradient (observations are data).

``policy_head_evaluate`` is the policy head of a PPO minibatch as one autograd node: the 13,527-wide linear layer
(computed 13,536 wide, see ``padded_linear``) + masked softmax / log-prob of the taken action / entropy
(kz_eval_masked_fwd), and backward kz_eval_masked_bwd straight into the padded dlogits buffer the two gradient GEMMs
read -- the logits never enter the autograd graph, so no slice / pad copies of [B, 13527] tensors."""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as nv

_LD = (nv.NUM_ACTIONS + 15) // 16 * 16


grad_reducer = None  # set by PPOAgent.enable_ddp (training.distributed.GradReducer): early all-reduce of the head's weight gradient


def _is_bitmap(mask: torch.Tensor) -> bool:
    return mask.dtype == torch.int32 and mask.shape[-1] == nv.BITMAP_WORDS


class _ObsConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obs, weight, bias, relu, rows, cobs):
        dev = nv.require_cuda(obs.device)
        nv.require(obs.dtype == torch.float32 and obs.dim() == 4 and tuple(obs.shape[1:]) == (46, 9, 9), "obs.dtype == torch.float32 and obs.dim() == 4 and tuple(obs.shape[1:]) == (46, 9, 9)")
        nv.require(tuple(weight.shape) == (16, 46, 3, 3), "tuple(weight.shape) == (16, 46, 3, 3)")
        obs = obs.contiguous()
        if rows is not None:
            rows = rows.contiguous().long()
        n = obs.shape[0] if rows is None else rows.shape[0]
        w = weight.detach().float().contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        y = torch.empty((n, 16, 9, 9), dtype=torch.bfloat16, device=dev)
        if cobs is not None:
            # the engine's compact observations of the same positions: 160 bytes per board instead of 14,904
            nv.require(cobs.dtype == torch.int32 and cobs.dim() == 2 and cobs.shape[1] == nv.COBS_WORDS and cobs.is_contiguous()
                       and cobs.device == dev and cobs.shape[0] == obs.shape[0],
                       "cobs: contiguous int32 [len(obs), 40] compact observations on the observations' device")
            nv.check(nv.lib().kz_cobs_conv_fwd(cobs.data_ptr(), nv.ptr(rows), w.data_ptr(), nv.ptr(b), 16, n, int(relu),
                                               y.data_ptr(), nv.stream_ptr(dev)), "kz_cobs_conv_fwd")
        else:
            nv.check(nv.lib().kz_obs_conv_fwd(obs.data_ptr(), nv.ptr(rows), w.data_ptr(), nv.ptr(b), 16, n, int(relu),
                                              y.data_ptr(), nv.stream_ptr(dev)), "kz_obs_conv_fwd")
        ctx.save_for_backward(obs, y, rows, cobs)
        ctx.relu, ctx.has_bias = bool(relu), bias is not None
        ctx.wdtype = weight.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        obs, y, rows, cobs = ctx.saved_tensors
        assert not ctx.needs_input_grad[0], "observations are data: the input layer has no input gradient"
        dev = obs.device
        n = y.shape[0]
        if dy.dtype not in (torch.bfloat16, torch.float32):
            dy = dy.float()
        dy = dy.contiguous()
        L = nv.lib()
        ctas = L.kz_cobs_conv_wgrad_ctas(n) if cobs is not None else L.kz_obs_conv_wgrad_ctas(n)
        ws = torch.empty(ctas * 16 * 432, dtype=torch.float32, device=dev)
        dw = torch.empty((16, 46, 3, 3), dtype=torch.float32, device=dev)
        db = torch.empty(16, dtype=torch.float32, device=dev) if ctx.has_bias else None
        args = (nv.ptr(rows), y.data_ptr() if ctx.relu else None, dy.data_ptr(), int(dy.dtype == torch.bfloat16), 16, n,
                ws.data_ptr(), ctas, dw.data_ptr(), nv.ptr(db), nv.stream_ptr(dev))
        if cobs is not None:
            nv.check(L.kz_cobs_conv_wgrad(cobs.data_ptr(), *args), "kz_cobs_conv_wgrad")
        else:
            nv.check(L.kz_obs_conv_wgrad(obs.data_ptr(), *args), "kz_obs_conv_wgrad")
        return None, dw.to(ctx.wdtype), (db.to(ctx.wdtype) if db is not None else None), None, None, None


def obs_conv(obs: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = True,
             rows: Optional[torch.Tensor] = None, cobs: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 [n, 16, 9, 9] = [relu](conv3x3(obs[rows], weight) + bias), operands rounded to bf16, fp32 accumulation.
    ``cobs`` (int32 [len(obs), 40]): the engine's compact observations of the same positions (kz_step_rollout); the layer
    then reads those 160 bytes per board instead of the 14,904-byte tensor rows -- same function of the same positions."""
    return _ObsConv.apply(obs, weight, bias, relu, rows, cobs)


def obs_conv_applicable(conv: torch.nn.Conv2d, x: torch.Tensor) -> bool:
    """The fused input layer stands in for ``relu(conv(x))`` exactly when autocast would run that conv in bf16."""
    return (x.is_cuda and x.dtype == torch.float32 and not x.requires_grad and x.dim() == 4
            and tuple(x.shape[1:]) == (46, 9, 9) and torch.is_autocast_enabled()
            and torch.get_autocast_dtype("cuda") == torch.bfloat16 and tuple(conv.weight.shape) == (16, 46, 3, 3)
            and conv.padding == (1, 1) and conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1
            and conv.padding_mode == "zeros")


class _PolicyHeadEval(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, weight, bias, mask, mask_rows, actions):
        dev = nv.require_cuda(h.device)
        n, k = h.shape
        hb = h.to(torch.bfloat16).contiguous()
        wp = torch.zeros((_LD, k), dtype=torch.bfloat16, device=dev)   # 16-aligned GEMM N (see padded_linear)
        wp[: nv.NUM_ACTIONS].copy_(weight.detach())
        bp = torch.zeros(_LD, dtype=torch.bfloat16, device=dev)
        if bias is not None:
            bp[: nv.NUM_ACTIONS].copy_(bias.detach())
        logits = torch.addmm(bp, hb, wp.t())                            # [n, 13536] bf16
        logp = torch.empty(n, dtype=torch.float32, device=dev)
        ent = torch.empty(n, dtype=torch.float32, device=dev)
        saved = torch.empty((n, 4), dtype=torch.float32, device=dev)
        actions = actions.contiguous().long()
        if mask_rows is not None:
            mask_rows = mask_rows.contiguous().long()
        fwd = nv.lib().kz_eval_bitmap_fwd if _is_bitmap(mask) else nv.lib().kz_eval_masked_fwd
        nv.check(fwd(logits.data_ptr(), 1, _LD, mask.data_ptr(), mask.stride(0), nv.ptr(mask_rows),
                     actions.data_ptr(), n, logp.data_ptr(), ent.data_ptr(), saved.data_ptr(),
                     nv.stream_ptr(dev)), "kz_eval_masked_fwd")
        ctx.save_for_backward(hb, wp, logits, mask, mask_rows, actions, saved)
        ctx.has_bias, ctx.hdtype, ctx.wdtype = bias is not None, h.dtype, weight.dtype
        ctx.weight_param = weight if isinstance(weight, torch.nn.Parameter) else None
        return logp, ent

    @staticmethod
    def backward(ctx, dlogp, dent):
        hb, wp, logits, mask, mask_rows, actions, saved = ctx.saved_tensors
        dev = hb.device
        n = hb.shape[0]
        dlogits = torch.empty((n, _LD), dtype=torch.bfloat16, device=dev)
        dlogits[:, nv.NUM_ACTIONS:].zero_()                             # the kernel clears and fills [0, 13527)
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        # bias gradient = column sums of dlogits: scatter-added by the same kernel (only legal entries are non-zero) as
        # Q20.44 fixed-point integers -- associative, hence deterministic run to run
        dbq = torch.zeros(_LD, dtype=torch.int64, device=dev) if want_db else None
        common = (logits.data_ptr(), 1, _LD, mask.data_ptr(), mask.stride(0), nv.ptr(mask_rows), actions.data_ptr(), n,
                  dlogp.contiguous().float().data_ptr(), dent.contiguous().float().data_ptr(), saved.data_ptr(),
                  dlogits.data_ptr(), _LD)
        if _is_bitmap(mask):
            nv.check(nv.lib().kz_eval_bitmap_bwd(*common, nv.ptr(dbq), nv.stream_ptr(dev)), "kz_eval_bitmap_bwd")
        elif want_db:
            nv.check(nv.lib().kz_eval_masked_bwd_bias(*common, dbq.data_ptr(), nv.stream_ptr(dev)), "kz_eval_masked_bwd_bias")
        else:
            nv.check(nv.lib().kz_eval_masked_bwd(*common, nv.stream_ptr(dev)), "kz_eval_masked_bwd")
        # the weight gradient first: under data parallelism its all-reduce (the bulk of the model's gradient bytes) starts
        # here and overlaps the input-gradient GEMM and everything backward still has to do (distributed.GradReducer)
        dw = None
        if ctx.needs_input_grad[1]:
            dw16 = torch.mm(dlogits.t(), hb)[: nv.NUM_ACTIONS]  # bf16, contiguous (leading rows of the padded product)
            if grad_reducer is not None and ctx.weight_param is not None and ctx.weight_param.grad is None:
                grad_reducer.early(ctx.weight_param, dw16)  # sets weight.grad itself once the cross-rank sum has arrived
            else:
                dw = dw16.to(ctx.wdtype)
        dh = torch.mm(dlogits, wp).to(ctx.hdtype) if ctx.needs_input_grad[0] else None
        db = (dbq[: nv.NUM_ACTIONS].to(torch.float64) * 2.0 ** -44).to(ctx.wdtype) if want_db else None
        return dh, dw, db, None, None, None


def policy_head_evaluate(h: torch.Tensor, linear: torch.nn.Linear, mask: torch.Tensor, actions: torch.Tensor,
                         mask_rows: Optional[torch.Tensor] = None):
    """(log-prob of ``actions``, entropy) of softmax(mask(linear(h))) for a [B, K] feature batch, bf16 GEMMs."""
    nv.require(linear.out_features == nv.NUM_ACTIONS and mask.stride(-1) == 1
               and (mask.dtype in (torch.uint8, torch.bool) or _is_bitmap(mask)),
               "policy_head_evaluate: 13,527-wide head and bool/uint8 mask rows or int32 [B, 448] legal bitmap rows")
    return _PolicyHeadEval.apply(h, linear.weight, linear.bias, mask, mask_rows, actions)
