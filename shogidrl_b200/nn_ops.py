"""Network-side operators around the rollout path that run as hand-written kernels (csrc/kz_nn.cu).

``obs_conv`` is the default model's input layer -- nn.Conv2d(46, 16, 3, padding=1) [+ ReLU] under bf16 autocast
(keisei/core/neural_network.py:14-28, keisei/core/ppo_agent.py:323) -- reading the fp32 observation batch
straight from the rollout storage: no separate cast pass, forward and weight gradient each one pass over the
observations.  The input gets no gradient (observations are data)."""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as nv


class _ObsConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obs, weight, bias, relu):
        dev = nv.require_cuda(obs.device)
        assert obs.dtype == torch.float32 and obs.dim() == 4 and tuple(obs.shape[1:]) == (46, 9, 9)
        assert tuple(weight.shape) == (16, 46, 3, 3)
        obs = obs.contiguous()
        n = obs.shape[0]
        w = weight.detach().float().contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        y = torch.empty((n, 16, 9, 9), dtype=torch.bfloat16, device=dev)
        nv.check(nv.lib().kz_obs_conv_fwd(obs.data_ptr(), w.data_ptr(), nv.ptr(b), 16, n, int(relu), y.data_ptr(),
                                          nv.stream_ptr(dev)), "kz_obs_conv_fwd")
        ctx.save_for_backward(obs, y)
        ctx.relu, ctx.has_bias = bool(relu), bias is not None
        ctx.wdtype = weight.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        obs, y = ctx.saved_tensors
        assert not ctx.needs_input_grad[0], "observations are data: the input layer has no input gradient"
        dev = obs.device
        n = obs.shape[0]
        if dy.dtype not in (torch.bfloat16, torch.float32):
            dy = dy.float()
        dy = dy.contiguous()
        L = nv.lib()
        ctas = L.kz_obs_conv_wgrad_ctas(n)
        ws = torch.empty(ctas * 16 * 432, dtype=torch.float32, device=dev)
        dw = torch.empty((16, 46, 3, 3), dtype=torch.float32, device=dev)
        db = torch.empty(16, dtype=torch.float32, device=dev) if ctx.has_bias else None
        nv.check(L.kz_obs_conv_wgrad(obs.data_ptr(), y.data_ptr() if ctx.relu else None, dy.data_ptr(),
                                     int(dy.dtype == torch.bfloat16), 16, n, ws.data_ptr(), ctas, dw.data_ptr(), nv.ptr(db),
                                     nv.stream_ptr(dev)), "kz_obs_conv_wgrad")
        return None, dw.to(ctx.wdtype), (db.to(ctx.wdtype) if db is not None else None), None


def obs_conv(obs: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = True) -> torch.Tensor:
    """bf16 [n, 16, 9, 9] = [relu](conv3x3(obs, weight) + bias), operands rounded to bf16, fp32 accumulation."""
    return _ObsConv.apply(obs, weight, bias, relu)


def obs_conv_applicable(conv: torch.nn.Conv2d, x: torch.Tensor) -> bool:
    """The fused input layer stands in for ``relu(conv(x))`` exactly when autocast would run that conv in bf16."""
    return (x.is_cuda and x.dtype == torch.float32 and not x.requires_grad and x.dim() == 4
            and tuple(x.shape[1:]) == (46, 9, 9) and torch.is_autocast_enabled()
            and torch.get_autocast_dtype("cuda") == torch.bfloat16 and tuple(conv.weight.shape) == (16, 46, 3, 3)
            and conv.padding == (1, 1) and conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1
            and conv.padding_mode == "zeros")
