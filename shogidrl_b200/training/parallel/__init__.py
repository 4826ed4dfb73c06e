"""keisei.training.parallel on a B200: the reference spreads self-play over CPU worker processes that pickle
tensors through mp.Queue and receive gzip-compressed weights (parallel_manager.py:23-345, self_play_worker.py,
communication.py, model_sync.py).  Here the games live on the device next to the model, so "workers" are rows of
one VecShogiEnv batch: ``ParallelManager.collect_experiences(buffer)`` keeps its signature and fills the buffer
from a batched rollout; model synchronisation is a no-op (same process, same weights)."""
from __future__ import annotations

import gzip
from typing import Any, Dict, Optional

import numpy as np
import torch

from ..step_manager import VecStepManager
from ...core.experience_buffer import RolloutBuffer


def compress_array(array: np.ndarray, compression_level: int = 6) -> Dict[str, Any]:
    """parallel/utils.py:8-37 (kept for callers that still ship weights between hosts)."""
    raw = np.ascontiguousarray(array).tobytes()
    comp = gzip.compress(raw, compresslevel=compression_level)
    return {"data": comp, "shape": array.shape, "dtype": str(array.dtype), "compressed": True,
            "original_size": len(raw), "compressed_size": len(comp), "compression_ratio": len(raw) / max(1, len(comp))}


def decompress_array(data: Dict[str, Any]) -> np.ndarray:
    raw = gzip.decompress(data["data"]) if data.get("compressed") else data["data"]
    return np.frombuffer(raw, dtype=np.dtype(data["dtype"])).reshape(data["shape"]).copy()


class ModelSynchronizer:
    """No-op shim: rollout and update share one model object on one device."""

    def __init__(self, sync_interval: int = 100, compression_enabled: bool = True):
        self.sync_interval, self.compression_enabled = sync_interval, compression_enabled
        self.last_sync_step, self.sync_count = 0, 0

    def should_sync(self, current_step: int) -> bool:
        return current_step - self.last_sync_step >= self.sync_interval

    def mark_sync_completed(self, current_step: int) -> None:
        self.last_sync_step, self.sync_count = current_step, self.sync_count + 1

    def get_sync_stats(self) -> Dict[str, Any]:
        return {"sync_count": self.sync_count, "last_sync_step": self.last_sync_step,
                "sync_interval": self.sync_interval}


class ParallelManager:
    """collect_experiences(buffer) fed by ``num_workers`` device-resident games instead of worker processes."""

    def __init__(self, env_config: Dict[str, Any], model_config: Dict[str, Any], parallel_config: Dict[str, Any],
                 device: str = "cuda"):
        self.env_config, self.model_config, self.parallel_config = env_config, model_config, parallel_config
        self.device = torch.device(device)
        self.num_workers = int(parallel_config.get("num_workers", 4))
        self.batch_size = int(parallel_config.get("batch_size", 32))
        self.model_sync = ModelSynchronizer(parallel_config.get("sync_interval", 100),
                                            parallel_config.get("compression_enabled", True))
        self.total_steps_collected = 0
        self.total_batches_received = 0
        self.is_running = False
        self._driver: Optional[VecStepManager] = None

    def start_workers(self, agent, gamma: float = 0.99, lambda_gae: float = 0.95, worker_semantics: bool = True) -> bool:
        """``worker_semantics`` keeps what distinguishes the reference's worker processes from its main loop
        (self_play_worker.py:105, 130, 190-214): every worker game is a ``ShogiGame()`` with the DEFAULT 500-move limit
        whatever ``env.max_moves_per_game`` says, and actions are sampled from the network in ``eval()`` mode.  (Its
        masking formula, softmax(all logits) * mask / (sum + 1e-8) handed to ``Categorical(probs=...)``, is the masked
        softmax the sampler draws from: Categorical renormalises, so the 1e-8 cancels.)"""
        from ...vec_env import VecShogiEnv
        max_moves = 500 if worker_semantics else int(self.env_config.get("max_moves_per_game", 500))
        env = VecShogiEnv(self.num_workers, max_moves_per_game=max_moves,
                          device=self.device, seed=int(self.env_config.get("seed", 0) or 0))
        buf = RolloutBuffer(self.batch_size, self.num_workers, gamma, lambda_gae, self.device)
        self._driver = VecStepManager(env, agent, buf)
        self._driver.model_eval_mode = bool(worker_semantics)
        self._driver.start()
        self.is_running = True
        return True

    def collect_experiences(self, experience_buffer) -> int:
        """One batched rollout of ``batch_size`` steps x ``num_workers`` games appended to ``experience_buffer``
        (an ExperienceBuffer); returns the number of transitions added."""
        if not self.is_running or self._driver is None:
            return 0
        d = self._driver
        d.collect()
        b = d.buffer
        T, N = b.T, b.N
        before = experience_buffer.size()
        experience_buffer.add_from_worker_batch({
            "obs": b.obs[:T].transpose(0, 1).reshape(T * N, 46, 9, 9), "actions": b.actions.t().reshape(-1),
            "rewards": b.rewards.t().reshape(-1), "log_probs": b.log_probs.t().reshape(-1),
            "values": b.values.t().reshape(-1), "dones": b.dones.t().reshape(-1).bool(),
            # worker-major order like the reference's per-worker batches: flat row of (n, t) is t * N + n
            "legal_masks": b.legal_masks((torch.arange(T, device=b.device)[None, :] * N
                                          + torch.arange(N, device=b.device)[:, None]).reshape(-1))})
        b.clear()
        added = experience_buffer.size() - before
        self.total_steps_collected += added
        self.total_batches_received += 1
        return added

    def sync_model_if_needed(self, model, current_step: int) -> bool:
        if self.model_sync.should_sync(current_step):
            self.model_sync.mark_sync_completed(current_step)
            return True
        return False

    def stop_workers(self) -> None:
        self.is_running = False
        self._driver = None

    def is_healthy(self) -> bool:
        return self.is_running and self._driver is not None

    def get_parallel_stats(self) -> Dict[str, Any]:
        return {"num_workers": self.num_workers, "total_steps_collected": self.total_steps_collected,
                "total_batches_received": self.total_batches_received, "is_running": self.is_running,
                "sync_stats": self.model_sync.get_sync_stats()}


class WorkerCommunicator:
    """Name-compatible stand-in for the mp.Queue plumbing of keisei/training/parallel/communication.py:22-236.
    There are no worker processes to talk to: control commands and model pushes are accepted and counted, and
    ``collect_experiences`` has nothing queued because rollouts are produced in-process by ParallelManager."""

    def __init__(self, num_workers: int, max_queue_size: int = 1000, timeout: float = 10.0):
        self.num_workers, self.max_queue_size, self.timeout = num_workers, max_queue_size, timeout
        self.commands_sent = 0
        self.model_updates_sent = 0

    def send_control_command(self, command: str, data: Optional[Dict] = None, worker_ids=None) -> None:
        self.commands_sent += 1

    def send_model_weights(self, model_state_dict: Dict[str, torch.Tensor], worker_ids=None,
                           compression_enabled: bool = True) -> None:
        self.model_updates_sent += 1  # same process, same weights: nothing to ship

    def collect_experiences(self):
        return []

    def get_queue_info(self) -> Dict[str, Any]:
        return {"num_workers": self.num_workers, "queued_batches": 0}

    def cleanup(self) -> None:
        pass


class SelfPlayWorker:
    """Name-compatible stand-in for keisei/training/parallel/self_play_worker.py:27-463.  A "worker" here is a
    contiguous slice of envs of one VecShogiEnv batch, not an OS process; ``run`` is therefore not available."""

    def __init__(self, worker_id: int, env_slice: slice):
        self.worker_id, self.env_slice = worker_id, env_slice

    def run(self) -> None:  # pragma: no cover - documented non-feature
        raise RuntimeError("self-play runs on the device inside ParallelManager.collect_experiences; "
                           "there are no worker processes to start")


__all__ = ["ParallelManager", "ModelSynchronizer", "WorkerCommunicator", "SelfPlayWorker", "compress_array",
           "decompress_array"]
