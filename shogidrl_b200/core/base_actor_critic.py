"""Actor-critic base class and the two reference model families.

``get_action_and_value`` keeps the signature of keisei/core/base_actor_critic.py:43-116 but, after the PyTorch
forward pass (the only dense contraction on the path), the masked softmax / Categorical sample / log-prob run in
one hand-written kernel (kz_sample_masked).  ``evaluate_actions`` (:118-184) runs the same formulas
(Categorical(probs) clamps probabilities to [eps, 1 - eps] before the log) as a fused forward/backward kernel pair
(kz_eval_masked_fwd/bwd) on CUDA and in plain PyTorch elsewhere.

Models: ``ActorCritic`` (keisei/core/neural_network.py:10-29) and ``ActorCriticResTower`` with optional
squeeze-excitation (keisei/training/models/resnet_tower.py:15-84).  Parameter names match the reference so that
its checkpoints load."""
from __future__ import annotations

import itertools
import sys
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import nn_ops, rl

_EPS = torch.finfo(torch.float32).eps
_sample_counter = itertools.count()


def padded_linear(x: torch.Tensor, linear: nn.Linear) -> torch.Tensor:
    """``linear(x)`` computed with the output width padded to a multiple of 16.  The 13,527-wide policy head is an
    odd GEMM N (and an odd output row stride), which sends cuBLAS to a misaligned bf16 path measured 10x slower on
    B200 (profiles/linear_pad_probe.py: 4.93 ms vs 0.50 ms forward for 16,384 x 1296 x 13527).  Parameters keep the
    reference shapes (checkpoints stay compatible); the result is the [:, :out_features] view of the padded GEMM."""
    out = linear.out_features
    pad = (-out) % 16
    if pad == 0 or not x.is_cuda:
        return linear(x)
    w = F.pad(linear.weight, (0, 0, 0, pad))
    b = F.pad(linear.bias, (0, pad)) if linear.bias is not None else None
    return F.linear(x, w, b)[:, :out]


class BaseActorCriticModel(nn.Module):
    sample_seed: int = 0x5EED

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:  # pragma: no cover - abstract
        raise NotImplementedError("Subclasses must implement forward method")

    def get_action_and_value(self, obs: torch.Tensor, legal_mask: Optional[torch.Tensor] = None,
                             deterministic: bool = False, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                             cobs: Optional[torch.Tensor] = None, draw_counter: Optional[torch.Tensor] = None
                             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """``legal_mask``: bool/uint8 [B, 13527] as in the reference, or the engine's legal bitmap rows (int32
        [B, 448]).  ``out`` = (actions int64 [B], log_probs fp32 [B]): the sampler writes there (rollout storage).
        ``cobs``: the engine's compact observations of ``obs`` (int32 [B, 40]) for models whose input layer reads them."""
        logits, value = self.forward(obs, cobs=cobs) if cobs is not None and getattr(self, "reads_compact_obs", False) \
            else self.forward(obs)
        if legal_mask is None:
            legal_mask = torch.ones_like(logits, dtype=torch.bool)
        elif legal_mask.dim() == 1:
            legal_mask = legal_mask.unsqueeze(0)
        if legal_mask.shape[0] != logits.shape[0]:
            legal_mask = legal_mask.expand(logits.shape[0], -1)
        if logits.dtype not in (torch.float32, torch.bfloat16):
            logits = logits.float()
        if draw_counter is not None:
            # the draw counter lives on the device and is advanced by a device op: the whole call can sit in a CUDA graph
            # and still draw fresh numbers on every replay
            action, log_prob, _ = rl.sample_masked(logits, legal_mask, seed=self.sample_seed, offset=0,
                                                   deterministic=deterministic, out=out, offset_tensor=draw_counter)
            draw_counter.add_(1 << 20)
        else:
            offset = next(_sample_counter) * (1 << 20)
            action, log_prob, _ = rl.sample_masked(logits, legal_mask, seed=self.sample_seed, offset=offset,
                                                   deterministic=deterministic, out=out)
        if value.dim() > 1 and value.shape[-1] == 1:
            value = value.squeeze(-1)
        return action, log_prob, value

    def evaluate_actions(self, obs: torch.Tensor, actions: torch.Tensor, legal_mask: Optional[torch.Tensor] = None
                         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        logits, value = self.forward(obs)
        if legal_mask is not None and not rl.is_bitmap(legal_mask) and legal_mask.dim() == 2 \
                and not bool(legal_mask.bool().any(dim=1).all()):
            # the public method reports rows without a legal action as the reference does (base_actor_critic.py:166-174);
            # the PPO update goes through evaluate_from_logits and never synchronises
            print(f"[{self.__class__.__name__}] ERROR: NaNs in probabilities in evaluate_actions. Check legal_mask and "
                  "logits. Defaulting to uniform for affected rows.", file=sys.stderr)
        return self.evaluate_from_logits(logits, value, actions, legal_mask)

    @staticmethod
    def evaluate_from_logits(logits: torch.Tensor, value: torch.Tensor, actions: torch.Tensor,
                             legal_mask: Optional[torch.Tensor] = None, mask_rows: Optional[torch.Tensor] = None):
        """log-prob / entropy / value from a forward pass that may have gone through a DDP wrapper.  On CUDA the
        masked softmax, log-prob gather, entropy and their backward run in one fused kernel pair
        (rl.evaluate_masked); ``mask_rows`` lets ``legal_mask`` be the whole rollout mask storage."""
        if logits.is_cuda and logits.dtype in (torch.float32, torch.bfloat16):
            if legal_mask is None:
                legal_mask = torch.ones((logits.shape[0], logits.shape[1]), dtype=torch.uint8, device=logits.device)
                mask_rows = None
            log_probs, entropy = rl.evaluate_masked(logits, legal_mask, actions, mask_rows)
            if value.dim() > 1 and value.shape[-1] == 1:
                value = value.squeeze(-1)
            return log_probs, entropy, value
        if mask_rows is not None and legal_mask is not None:
            legal_mask = legal_mask[mask_rows]
        if legal_mask is not None and rl.is_bitmap(legal_mask):  # CPU tensors: unpack the bitmap rows
            bits = (legal_mask.unsqueeze(-1) >> torch.arange(32, device=legal_mask.device, dtype=torch.int32)) & 1
            legal_mask = bits.reshape(legal_mask.shape[0], -1)[:, : logits.shape[1]].bool()
        logits = logits.float()
        if legal_mask is not None:
            logits = torch.where(legal_mask.bool(), logits, torch.full((), float("-inf"), device=logits.device))
        probs = F.softmax(logits, dim=-1)
        nan_rows = torch.isnan(probs).any(dim=1, keepdim=True)
        probs = torch.where(nan_rows, torch.full_like(probs, 1.0 / probs.shape[-1]), probs)
        log_p = torch.log(probs.clamp(min=_EPS, max=1.0 - _EPS))  # Categorical(probs=...).logits
        log_probs = log_p.gather(1, actions.long().unsqueeze(1)).squeeze(1)
        entropy = -(probs * log_p).sum(dim=-1)
        if value.dim() > 1 and value.shape[-1] == 1:
            value = value.squeeze(-1)
        return log_probs, entropy, value


class ActorCritic(BaseActorCriticModel):
    """Conv(46->16, 3x3) + ReLU + Flatten + Linear heads (BASELINE config 3's "default CNN")."""

    def __init__(self, input_channels: int, num_actions_total: int):
        super().__init__()
        self.conv = nn.Conv2d(input_channels, 16, kernel_size=3, padding=1)
        self.relu = nn.ReLU()
        self.flatten = nn.Flatten()
        self.policy_head = nn.Linear(16 * 81, num_actions_total)
        self.value_head = nn.Linear(16 * 81, 1)

    fused_minibatch = True  # forward(obs_store, rows=..., actions=..., legal_mask=...) evaluates a PPO minibatch
    reads_compact_obs = True  # forward(x, cobs=...) feeds the input layer from the engine's 160-byte compact observations

    def forward(self, x, rows=None, actions=None, legal_mask=None, mask_rows=None, cobs=None):
        """``forward(x)`` -> (logits, value) as in the reference.  The PPO update calls (through the DDP wrapper when
        there is one) ``forward(obs_store, rows=mb, actions=a, legal_mask=mask_store, mask_rows=mb)`` and gets
        (log_probs, entropy, value) of the minibatch: observations and masks are read in place through the row
        indices, the input layer and the policy head + masked evaluation each run as one fused node."""
        if nn_ops.obs_conv_applicable(self.conv, x):  # bf16 autocast on CUDA: fused input layer (csrc/kz_nn.cu)
            h = self.flatten(nn_ops.obs_conv(x, self.conv.weight, self.conv.bias, relu=True, rows=rows, cobs=cobs))
        else:
            h = self.flatten(self.relu(self.conv(x if rows is None else x[rows])))
        if actions is None:
            cache = getattr(self, "_head_cache", None) if getattr(self, "_head_cache_live", False) else None
            if cache is not None and h.dtype == torch.bfloat16 and not torch.is_grad_enabled():
                # rollout: the 13,536-wide bf16 copy of the head's weights made once per rollout (cache_inference_weights)
                return F.linear(h, cache[0], cache[1])[:, : self.policy_head.out_features], self.value_head(h)
            return padded_linear(h, self.policy_head), self.value_head(h)
        value = self.value_head(h).squeeze(-1)
        if h.is_cuda and h.dtype == torch.bfloat16 and legal_mask is not None:
            log_probs, entropy = nn_ops.policy_head_evaluate(h, self.policy_head, legal_mask, actions, mask_rows)
            return log_probs, entropy, value
        return self.evaluate_from_logits(padded_linear(h, self.policy_head), value, actions, legal_mask, mask_rows)


    def cache_inference_weights(self, live: bool = True) -> None:
        """Refresh the padded bf16 copy of the policy head that no-grad bf16 forwards use while ``live`` (a rollout: the
        parameters do not change during it, so the per-step pad (70 MB) + autocast cast (105 MB) of ``padded_linear`` is
        paid once per rollout); ``live=False`` switches the copy off again.  The buffers keep their addresses, so a
        captured rollout graph reads the refreshed values."""
        object.__setattr__(self, "_head_cache_live", bool(live))
        if not live:
            return
        lin = self.policy_head
        pad = (-lin.out_features) % 16
        if getattr(self, "_head_cache", None) is None:
            w = torch.zeros((lin.out_features + pad, lin.in_features), dtype=torch.bfloat16, device=lin.weight.device)
            b = torch.zeros(lin.out_features + pad, dtype=torch.bfloat16, device=lin.weight.device)
            object.__setattr__(self, "_head_cache", (w, b))  # plain attribute: not a parameter / buffer, not in state_dict
        w, b = self._head_cache
        with torch.no_grad():
            w[: lin.out_features].copy_(lin.weight)
            if lin.bias is not None:
                b[: lin.out_features].copy_(lin.bias)


class SqueezeExcitation(nn.Module):
    def __init__(self, channels: int, se_ratio: float = 0.25):
        super().__init__()
        hidden = max(1, int(channels * se_ratio))
        self.fc1 = nn.Conv2d(channels, hidden, 1)
        self.fc2 = nn.Conv2d(hidden, channels, 1)

    def forward(self, x):
        s = torch.sigmoid(self.fc2(F.relu(self.fc1(F.adaptive_avg_pool2d(x, 1)))))
        return x * s


class ResidualBlock(nn.Module):
    def __init__(self, channels: int, se_ratio: Optional[float] = None):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(channels)
        self.conv2 = nn.Conv2d(channels, channels, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(channels)
        self.se = SqueezeExcitation(channels, se_ratio) if se_ratio else None

    def forward(self, x):
        out = self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x)))))
        if self.se is not None:
            out = self.se(out)
        return F.relu(out + x)


class ActorCriticResTower(BaseActorCriticModel):
    prefers_channels_last = True  # PPOAgent converts the parameters once, on CUDA

    def __init__(self, input_channels: int, num_actions_total: int, tower_depth: int = 9, tower_width: int = 256,
                 se_ratio: Optional[float] = None):
        super().__init__()
        self.stem = nn.Conv2d(input_channels, tower_width, 3, padding=1)
        self.bn_stem = nn.BatchNorm2d(tower_width)
        self.res_blocks = nn.Sequential(*[ResidualBlock(tower_width, se_ratio) for _ in range(tower_depth)])
        self.policy_head = nn.Sequential(nn.Conv2d(tower_width, 2, 1), nn.BatchNorm2d(2), nn.ReLU(), nn.Flatten(),
                                         nn.Linear(2 * 81, num_actions_total))
        self.value_head = nn.Sequential(nn.Conv2d(tower_width, 2, 1), nn.BatchNorm2d(2), nn.ReLU(), nn.Flatten(),
                                        nn.Linear(2 * 81, 1))

    def forward(self, x):
        if x.is_cuda:
            # cuDNN's bf16 tensor-core convolutions want NHWC: 210 -> 318 TFLOP/s forward, 206 -> 340 forward + backward
            # on B200 for this tower (profiles/tower_format_probe.py).  Logical shapes, state_dict and results unchanged.
            if self.stem.weight.is_contiguous(memory_format=torch.contiguous_format) and self.stem.weight.dim() == 4 \
                    and not self.stem.weight.is_contiguous(memory_format=torch.channels_last):
                self.to(memory_format=torch.channels_last)
            x = x.contiguous(memory_format=torch.channels_last)
        x = self.res_blocks(F.relu(self.bn_stem(self.stem(x))))
        policy = padded_linear(self.policy_head[:-1](x), self.policy_head[-1])
        return policy, self.value_head(x).squeeze(-1)


def model_factory(model_type, obs_shape, num_actions, tower_depth, tower_width, se_ratio, **kwargs):
    """keisei/training/models/__init__.py:6-31."""
    if model_type == "resnet":
        return ActorCriticResTower(obs_shape[0], num_actions, tower_depth, tower_width, se_ratio, **kwargs)
    if model_type in ("dummy", "testmodel", "resumemodel"):
        return ActorCriticResTower(obs_shape[0], num_actions, 1, 16, None, **kwargs)
    raise ValueError(f"Unknown model_type: {model_type}")
