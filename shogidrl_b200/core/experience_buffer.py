"""Experience storage and GAE.

``ExperienceBuffer`` keeps the contract of keisei/core/experience_buffer.py (same tensors, dtypes, messages and
errors) with the reverse-time GAE scan done by kz_gae on the device.  ``RolloutBuffer`` is the batched [T, N]
layout the vectorised engine writes into directly (observations and masks are written by kz_step itself); with
N = 1 it degenerates to the reference's flat buffer."""
from __future__ import annotations

import sys
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from .. import _native as nv
from .. import rl


def _warn(component: str, message: str) -> None:
    print(f"[{component}] WARNING: {message}", file=sys.stderr)


@dataclass
class Experience:
    obs: torch.Tensor
    action: int
    reward: float
    log_prob: float
    value: float
    done: bool
    legal_mask: torch.Tensor


class ExperienceBuffer:
    def __init__(self, buffer_size: int, gamma: float, lambda_gae: float, device: str = "cpu"):
        self.buffer_size = buffer_size
        self.gamma = gamma
        self.lambda_gae = lambda_gae
        self.device = torch.device(device)
        d = self.device
        self.obs = torch.zeros((buffer_size, 46, 9, 9), dtype=torch.float32, device=d)
        self.actions = torch.zeros(buffer_size, dtype=torch.int64, device=d)
        self.rewards = torch.zeros(buffer_size, dtype=torch.float32, device=d)
        self.log_probs = torch.zeros(buffer_size, dtype=torch.float32, device=d)
        self.values = torch.zeros(buffer_size, dtype=torch.float32, device=d)
        self.dones = torch.zeros(buffer_size, dtype=torch.bool, device=d)
        self.legal_masks = torch.zeros((buffer_size, 13527), dtype=torch.bool, device=d)
        self.advantages = torch.zeros(buffer_size, dtype=torch.float32, device=d)
        self.returns = torch.zeros(buffer_size, dtype=torch.float32, device=d)
        self.ptr = 0
        self._advantages_computed = False

    def add(self, obs: torch.Tensor, action: int, reward: float, log_prob: float, value: float, done: bool,
            legal_mask: torch.Tensor):
        if self.ptr >= self.buffer_size:
            _warn("ExperienceBuffer", "Buffer is full. Cannot add new experience.")
            return
        i = self.ptr
        self.obs[i] = obs.to(self.device)
        self.actions[i] = action
        self.rewards[i] = reward
        self.log_probs[i] = log_prob
        self.values[i] = value
        self.dones[i] = done
        self.legal_masks[i] = legal_mask.to(self.device)
        self.ptr += 1

    def compute_advantages_and_returns(self, last_value: float):
        """GAE over the flat sequence (experience_buffer.py:99-145) on the device: the bit-exact column kernel
        with N = 1 (same fp32 operation order as the reference's eager loop)."""
        if self.ptr == 0:
            _warn("ExperienceBuffer", "compute_advantages_and_returns called on an empty buffer.")
            return
        nv.require_cuda(self.device)  # no CPU fallback
        T = self.ptr
        lv = torch.tensor([last_value], dtype=torch.float32, device=self.device)
        rl.gae(self.rewards[:T].view(T, 1), self.values[:T].view(T, 1), self.dones[:T].view(T, 1), lv, self.gamma,
               self.lambda_gae, exact=True, out=(self.advantages[:T].view(T, 1), self.returns[:T].view(T, 1)))
        self._advantages_computed = True

    def get_batch(self) -> dict:
        if self.ptr == 0:
            _warn("ExperienceBuffer", "get_batch called on an empty or not-yet-computed buffer.")
            return {}
        if not self._advantages_computed:
            raise RuntimeError("Cannot get batch: compute_advantages_and_returns() must be called first")
        n = self.ptr
        return {"obs": self.obs[:n], "actions": self.actions[:n], "log_probs": self.log_probs[:n],
                "values": self.values[:n], "rewards": self.rewards[:n], "advantages": self.advantages[:n],
                "returns": self.returns[:n], "dones": self.dones[:n], "legal_masks": self.legal_masks[:n]}

    def clear(self):
        self.ptr = 0
        self._advantages_computed = False

    def __len__(self):
        return self.ptr

    def size(self) -> int:
        return self.ptr

    def capacity(self) -> int:
        return self.buffer_size

    def add_batch(self, experiences: List[Experience]) -> None:
        for e in experiences:
            if self.ptr >= self.buffer_size:
                break
            self.add(e.obs, e.action, e.reward, e.log_prob, e.value, e.done, e.legal_mask)

    def add_from_worker_batch(self, worker_data: Dict[str, torch.Tensor]) -> None:
        """Bulk insert of a worker batch (experience_buffer.py:232-253) with slice copies instead of per-sample
        ``add`` calls; rows beyond the capacity are dropped, as in the reference."""
        k = min(int(worker_data["obs"].shape[0]), self.buffer_size - self.ptr)
        if k <= 0:
            return
        s = slice(self.ptr, self.ptr + k)
        d = self.device
        self.obs[s] = worker_data["obs"][:k].to(d)
        self.actions[s] = worker_data["actions"][:k].to(d)
        self.rewards[s] = worker_data["rewards"][:k].to(d)
        self.log_probs[s] = worker_data["log_probs"][:k].to(d)
        self.values[s] = worker_data["values"][:k].to(d)
        self.dones[s] = worker_data["dones"][:k].to(d)
        self.legal_masks[s] = worker_data["legal_masks"][:k].to(d)
        self.ptr += k

    def get_worker_batch_format(self) -> Optional[Dict[str, torch.Tensor]]:
        if self.ptr == 0:
            return None
        n = self.ptr
        return {"obs": self.obs[:n], "actions": self.actions[:n], "rewards": self.rewards[:n],
                "log_probs": self.log_probs[:n], "values": self.values[:n], "dones": self.dones[:n],
                "legal_masks": self.legal_masks[:n]}

    def merge_from_parallel_buffers(self, parallel_buffers: List["ExperienceBuffer"]) -> None:
        for b in parallel_buffers:
            if b.ptr:
                self.add_from_worker_batch(b.get_worker_batch_format())


class RolloutBuffer:
    """[T, N] rollout storage in HBM.  obs[t] / bitmaps[t] are the positions the action of step t was chosen in;
    kz_step_rollout writes obs[t+1] / bitmaps[t+1] directly (slot T holds the bootstrap observation).

    Legal sets are kept as the engine's 13,527-bit legal bitmaps (int32 [T+1, N, 448], 1,792 B per position) -- the
    sampler and both masked-evaluation kernels read them as they are.  The reference's byte masks
    (``ExperienceBuffer.legal_masks``, bool [B, 13527], experience_buffer.py:52-54) are a 7.5x larger encoding of the
    same sets (13.5 KB per position: 28 GB of a config-3 rollout) and are materialised only for callers that ask:
    ``masks_at(t)``, ``legal_masks(rows)`` or ``get_batch(expand_masks=True)``."""

    def __init__(self, horizon: int, num_envs: int, gamma: float, lambda_gae: float, device="cuda"):
        self.T, self.N = int(horizon), int(num_envs)
        self.gamma, self.lambda_gae = gamma, lambda_gae
        self.device = nv.require_cuda(device)
        d, T, N = self.device, self.T, self.N
        self.obs = torch.zeros((T + 1, N, 46, 9, 9), dtype=torch.float32, device=d)
        self.bitmaps = torch.zeros((T + 1, N, nv.BITMAP_WORDS), dtype=torch.int32, device=d)
        # compact observations (160 B each: piece-plane index per square + the constant-plane values), written by the
        # engine beside obs[t]: what the policy network's input layer reads in the rollout and in every update epoch
        self.cobs = torch.zeros((T + 1, N, nv.COBS_WORDS), dtype=torch.int32, device=d)
        self.actions = torch.zeros((T, N), dtype=torch.int64, device=d)
        self.log_probs = torch.zeros((T, N), dtype=torch.float32, device=d)
        self.values = torch.zeros((T, N), dtype=torch.float32, device=d)
        self.rewards = torch.zeros((T, N), dtype=torch.float32, device=d)
        self.dones = torch.zeros((T, N), dtype=torch.uint8, device=d)
        self.advantages = torch.zeros((T, N), dtype=torch.float32, device=d)
        self.returns = torch.zeros((T, N), dtype=torch.float32, device=d)
        self._advantages_computed = False

    def compute_advantages_and_returns(self, last_values: torch.Tensor):
        """One reverse scan per env column; the bootstrap value is used even if the last transition was terminal
        (its done flag masks it), as in trainer.py:229-232."""
        rl.gae(self.rewards, self.values, self.dones, last_values, self.gamma, self.lambda_gae,
               out=(self.advantages, self.returns))
        self._advantages_computed = True

    def masks_at(self, t: int) -> torch.Tensor:
        """bool [N, 13527] legal masks of slot t (expanded from the bitmaps on demand)."""
        return rl.bitmap_to_mask(self.bitmaps[t])

    def legal_masks(self, rows: Optional[torch.Tensor] = None) -> torch.Tensor:
        """bool [len(rows), 13527] legal masks of flat transitions ``rows`` (index t * N + n; default: all T * N --
        13.5 KB each, so ask for what is needed)."""
        flat = self.bitmaps[: self.T].reshape(self.T * self.N, nv.BITMAP_WORDS)
        return rl.bitmap_to_mask(flat, rows)

    def get_batch(self, expand_masks: bool = False) -> Dict[str, torch.Tensor]:
        """The reference's batch dictionary (experience_buffer.py:147-189) over the T * N transitions, with the legal
        sets as ``legal_bitmaps`` (int32 [B, 448]); ``expand_masks`` adds the reference's ``legal_masks`` bool
        [B, 13527]."""
        if not self._advantages_computed:
            raise RuntimeError("Cannot get batch: compute_advantages_and_returns() must be called first")
        T, N = self.T, self.N
        B = T * N
        batch = {"obs": self.obs[:T].reshape(B, 46, 9, 9), "actions": self.actions.reshape(B),
                 "log_probs": self.log_probs.reshape(B), "values": self.values.reshape(B),
                 "rewards": self.rewards.reshape(B), "advantages": self.advantages.reshape(B),
                 "returns": self.returns.reshape(B), "dones": self.dones.reshape(B).bool(),
                 "legal_bitmaps": self.bitmaps[:T].reshape(B, nv.BITMAP_WORDS),
                 "compact_obs": self.cobs[:T].reshape(B, nv.COBS_WORDS)}
        if expand_masks:
            batch["legal_masks"] = self.legal_masks()
        return batch

    def clear(self):
        """Start the next rollout from the last written state: slot T becomes slot 0."""
        self.obs[0].copy_(self.obs[self.T])
        self.bitmaps[0].copy_(self.bitmaps[self.T])
        self.cobs[0].copy_(self.cobs[self.T])
        self._advantages_computed = False
