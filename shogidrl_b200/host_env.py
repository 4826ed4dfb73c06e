"""Host-facing stepping of a device batch: ``HostPipelinedEnv``.

A host-side consumer (a CPU policy, a remote actor, the reference's StepManager loop) pays, per step, one
host->device copy of the actions, the step kernel, one device->host copy of the packed results (15 bytes per
game, ``VecShogiEnv.results``) and a stream synchronisation.  Run serially those latencies add to the kernel time.
Here the batch is cut into G independent groups (separate engine state, CUDA stream and pinned buffers, the same
per-game RNG streams as one big batch): while one group's kernel runs, the host consumes another group's results
and submits its next actions, so the copies and wake-ups hide behind compute.  Observations and legal masks stay
in HBM (written straight into the caller's rollout storage) for the policy tower."""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _native as nv
from .vec_env import VecShogiEnv


class HostPipelinedEnv:
    def __init__(self, num_envs: int, groups: int = 2, max_moves_per_game: int = 500, device="cuda", seed: int = 1234,
                 env_offset: int = 0, auto_reset: bool = True):
        self.device = nv.require_cuda(device)
        nv.require(groups >= 1 and num_envs % groups == 0, "num_envs must divide into equal groups")
        self.n, self.G, self.ng = int(num_envs), int(groups), int(num_envs) // int(groups)
        self.envs: List[VecShogiEnv] = [
            VecShogiEnv(self.ng, max_moves_per_game, self.device, seed=seed, env_offset=env_offset + g * self.ng,
                        auto_reset=auto_reset, step_streams=1) for g in range(self.G)]  # the groups already overlap
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.G)]
        self.events = [torch.cuda.Event() for _ in range(self.G)]
        self.d_actions = [torch.zeros(self.ng, dtype=torch.int64, device=self.device) for _ in range(self.G)]
        self.h_actions = [torch.zeros(self.ng, dtype=torch.int64).pin_memory() for _ in range(self.G)]
        self.h_results = [torch.zeros(15 * self.ng, dtype=torch.uint8).pin_memory() for _ in range(self.G)]
        torch.cuda.synchronize(self.device)  # construction ran on the current stream

    def h2d_bytes_per_step(self) -> int:
        return 8 * self.n

    def d2h_bytes_per_step(self) -> int:
        return 15 * self.n

    def prime(self, random_actions: bool = True) -> None:
        """Recompute obs / masks (and, with ``random_actions``, a uniform-random legal action per game) for the
        current positions and bring the packed results to the host, as after a step."""
        for g, env in enumerate(self.envs):
            with torch.cuda.stream(self.streams[g]):
                env.refresh(random_actions=random_actions)
                self.h_results[g].copy_(env.results, non_blocking=True)
                self.events[g].record()

    def submit(self, g: int, obs: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
               random_actions: bool = False) -> None:
        """Queue one step of group ``g`` with the actions in ``h_actions[g]`` (pinned host memory); returns at once."""
        env = self.envs[g]
        with torch.cuda.stream(self.streams[g]):
            self.d_actions[g].copy_(self.h_actions[g], non_blocking=True)
            env.step(self.d_actions[g], obs=obs, mask=mask, random_actions=random_actions)
            self.h_results[g].copy_(env.results, non_blocking=True)
            self.events[g].record()

    def wait(self, g: int):
        """Block until group ``g``'s last submitted step has landed in host memory; returns host views
        (next_actions int64, reward f32, done u8, reason u8, winner i8) of its packed results."""
        self.events[g].synchronize()
        r, n = self.h_results[g], self.ng
        return (r[: 8 * n].view(torch.int64), r[8 * n: 12 * n].view(torch.float32), r[12 * n: 13 * n], r[13 * n: 14 * n],
                r[14 * n:].view(torch.int8))

    def synchronize(self) -> None:
        for s in self.streams:
            s.synchronize()
