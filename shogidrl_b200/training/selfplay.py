"""SelfPlayTrainer -- the batched counterpart of one TrainingLoopManager epoch + Trainer.perform_ppo_update
(keisei/training/training_loop_manager.py:325-454, trainer.py:214-269): T steps of N device-resident games
(tower forward in bf16, fused masked sampling, engine step writing straight into the rollout buffer), bootstrap
values, GAE on the device, PPO update, buffer hand-over.  One process per GPU; under torch.distributed each rank
owns its own games and only gradients (DDP) and three normalisation scalars cross NVLink."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ..core.experience_buffer import RolloutBuffer
from ..core.ppo_agent import PPOAgent
from ..vec_env import VecShogiEnv
from . import distributed as kd
from .step_manager import VecStepManager


class SelfPlayTrainer:
    def __init__(self, model, config, num_envs: int, horizon: int, device="cuda", use_mixed_precision: bool = True,
                 total_envs: Optional[int] = None):
        rank, world = kd.world()
        self.device = torch.device(device)
        offset = kd.shard_envs(total_envs, rank, world)[0] if total_envs else rank * num_envs
        seed = getattr(config.env, "seed", 0) or 0
        self.env = VecShogiEnv(num_envs, max_moves_per_game=config.env.max_moves_per_game, device=self.device,
                               seed=int(seed), env_offset=offset, auto_reset=True)
        self.agent = PPOAgent(model, config, self.device, use_mixed_precision=use_mixed_precision)
        self.agent.model.sample_seed = int(seed) * 7919 + rank
        self.agent.enable_ddp()
        self.buffer = RolloutBuffer(horizon, num_envs, config.training.gamma, config.training.lambda_gae, self.device)
        self.driver = VecStepManager(self.env, self.agent, self.buffer)
        self.global_timestep = 0

    def collect(self) -> Dict[str, int]:
        self.driver.collect()
        stats = self.driver.finish()
        self.global_timestep += self.buffer.T * self.buffer.N
        return stats

    def update(self) -> Dict[str, float]:
        metrics = self.agent.learn(self.buffer)
        self.buffer.clear()
        return metrics

    def run_epoch(self) -> Dict[str, float]:
        stats = self.collect()
        metrics = self.update()
        metrics.update({f"episodes/{k}": float(v) for k, v in stats.items()})
        return metrics
