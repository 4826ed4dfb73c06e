"""PPOAgent -- API of keisei/core/ppo_agent.py (select_action :134-223, get_value :225-241, learn :243-460,
save/load :462-534) on top of the device kernels.

* ``select_action`` / ``select_actions``: PyTorch forward (bf16 autocast optional) + kz_sample_masked.
* ``learn``: PPO-clip in PyTorch (the dense path stays PyTorch by design), whole-buffer advantage normalisation,
  numpy-seeded minibatch shuffling exactly as the reference; under torch.distributed the normalisation statistics
  are all-reduced so that sharded rollouts normalise like one big buffer."""
from __future__ import annotations

import copy
import os
import sys
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .. import rl

from ..utils.policy_mapper import PolicyOutputMapper


def _log(level: str, msg: str) -> None:
    print(f"[PPOAgent] {level}: {msg}", file=sys.stderr)


def _make_scheduler(optimizer, schedule_type, total_steps, kwargs):
    """Minimal counterpart of keisei/core/scheduler_factory.py:11-110."""
    if not schedule_type:
        return None
    kwargs = dict(kwargs or {})
    total_steps = max(1, int(total_steps))
    if schedule_type == "linear":
        final = kwargs.get("final_lr_fraction", 0.1)
        return torch.optim.lr_scheduler.LambdaLR(optimizer, lambda s: 1.0 - (1.0 - final) * min(s, total_steps) / total_steps)
    if schedule_type == "cosine":
        base = optimizer.param_groups[0]["lr"]
        return torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=total_steps,
                                                          eta_min=base * kwargs.get("eta_min_fraction", 0.0))
    if schedule_type == "exponential":
        return torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=kwargs.get("gamma", 0.995))
    if schedule_type == "step":
        return torch.optim.lr_scheduler.StepLR(optimizer, step_size=kwargs.get("step_size", max(1, total_steps // 3)),
                                               gamma=kwargs.get("gamma", 0.5))
    raise ValueError(f"Unsupported scheduler type: {schedule_type}")


class PPOAgent:
    def __init__(self, model, config, device: torch.device, name: str = "PPOAgent", scaler=None,
                 use_mixed_precision: bool = False):
        self.config = config.model_copy(deep=True) if hasattr(config, "model_copy") else copy.deepcopy(config)
        self.device = torch.device(device)
        self.name = name
        self.scaler = scaler
        self.use_mixed_precision = use_mixed_precision
        self.model = model.to(self.device)
        if self.device.type == "cuda" and getattr(self.model, "prefers_channels_last", False):
            self.model.to(memory_format=torch.channels_last)  # before the optimizer / DDP wrapper look at the parameters
            for p in self.model.parameters():
                # 1x1 kernels are the same bytes in both layouts: keep the default strides, which is how autograd lays out
                # their gradients (DDP's bucket views compare strides)
                if p.dim() == 4 and tuple(p.shape[2:]) == (1, 1):
                    p.data = torch.empty_like(p, memory_format=torch.contiguous_format).copy_(p)
        self.policy_output_mapper = PolicyOutputMapper()
        self.num_actions_total = self.policy_output_mapper.get_total_actions()
        tr = config.training
        weight_decay = getattr(tr, "weight_decay", 0.0)
        # capturable: step counters live on the device, so a whole minibatch update can sit in one CUDA graph
        adam = dict(weight_decay=weight_decay, capturable=self.device.type == "cuda")
        try:
            self.optimizer = torch.optim.Adam(self.model.parameters(), lr=tr.learning_rate, **adam)
        except Exception as e:  # same fallback as the reference (ppo_agent.py:72-80)
            _log("ERROR", f"Could not initialize optimizer with lr={tr.learning_rate}, using default lr=1e-3: {e}")
            self.optimizer = torch.optim.Adam(self.model.parameters(), lr=1e-3, **adam)
        self.cuda_graph_update = bool(getattr(tr, "cuda_graph_update", True))
        self._graph = None          # captured minibatch update (forward, losses, backward, clip, Adam)
        self._graph_key = None
        self._graph_warm = 0
        self.gamma = tr.gamma
        self.clip_epsilon = tr.clip_epsilon
        self.value_loss_coeff = tr.value_loss_coeff
        self.entropy_coef = tr.entropy_coef
        self.ppo_epochs = tr.ppo_epochs
        self.minibatch_size = tr.minibatch_size
        self.normalize_advantages = getattr(tr, "normalize_advantages", True)
        self.enable_value_clipping = getattr(tr, "enable_value_clipping", False)
        self.gradient_clip_max_norm = tr.gradient_clip_max_norm
        self.last_kl_div = 0.0
        self.last_gradient_norm = 0.0
        self._rng = np.random.default_rng(getattr(config.env, "seed", None))
        self.lr_schedule_type = getattr(tr, "lr_schedule_type", None)
        self.lr_schedule_step_on = getattr(tr, "lr_schedule_step_on", "epoch")
        epochs = max(1, getattr(tr, "total_timesteps", tr.steps_per_epoch) // tr.steps_per_epoch)
        if self.lr_schedule_step_on == "epoch":
            total = epochs * tr.ppo_epochs
        else:
            total = epochs * (tr.steps_per_epoch // tr.minibatch_size) * tr.ppo_epochs
        self.scheduler = _make_scheduler(self.optimizer, self.lr_schedule_type, total,
                                         getattr(tr, "lr_schedule_kwargs", None))

    # ------------------------------------------------------------------ acting
    def _is_obs_scaler(self) -> bool:
        return self.scaler is not None and not isinstance(self.scaler, torch.amp.GradScaler)

    def _scale(self, obs: torch.Tensor) -> torch.Tensor:
        if self._is_obs_scaler():
            return self.scaler.transform(obs) if hasattr(self.scaler, "transform") else self.scaler(obs)
        return obs

    def _autocast(self):
        return torch.autocast("cuda", dtype=torch.bfloat16, enabled=bool(self.use_mixed_precision and self.device.type == "cuda"))

    def select_actions(self, obs: torch.Tensor, legal_mask: torch.Tensor, *, is_training: bool = True,
                       out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, cobs: Optional[torch.Tensor] = None,
                       draw_counter: Optional[torch.Tensor] = None, eval_mode: bool = False
                       ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Batched select_action: obs [N,46,9,9] and legal_mask [N,13527] (or the engine's legal bitmap rows, int32
        [N,448]) stay on the device; returns (actions int64 [N], log_probs fp32 [N], values fp32 [N]) without any
        host synchronisation.  ``out`` = (actions, log_probs) rows of the rollout storage to write into."""
        # eval_mode: sample (is_training) from a network in eval() -- what the reference's self-play workers do
        # (training/parallel/self_play_worker.py:130, 190-214)
        self.model.train(is_training and not eval_mode)
        kw = {"out": out} if out is not None else {}
        if cobs is not None and not self._is_obs_scaler():
            kw["cobs"] = cobs  # the engine's compact observations of `obs` (input layers that read them skip the 14.9 KB rows)
        if draw_counter is not None:
            kw["draw_counter"] = draw_counter  # device-side sampling counter (CUDA-graph rollouts)
        with torch.no_grad(), self._autocast():
            action, log_prob, value = self.model.get_action_and_value(self._scale(obs), legal_mask=legal_mask,
                                                                      deterministic=not is_training, **kw)
        return action, log_prob, value.float()

    def select_action(self, obs: np.ndarray, legal_mask: torch.Tensor, *, is_training: bool = True):
        """(MoveTuple, policy index, log_prob, value) for one observation (ppo_agent.py:134-223)."""
        obs_tensor = torch.as_tensor(obs, dtype=torch.float32, device=self.device).unsqueeze(0)
        if not bool(legal_mask.any()):
            _log("ERROR", "select_action called with no legal moves (based on input legal_mask)")
            # ... and the line the reference's model prints when every logit is masked (base_actor_critic.py:92-101); the
            # sampling kernel falls back to the same uniform distribution without a host round trip, so it is said here
            print(f"[{type(getattr(self.model, 'module', self.model)).__name__}] ERROR: NaNs in probabilities in "
                  "get_action_and_value. Check legal_mask and logits. Defaulting to uniform.", file=sys.stderr)
        action, log_prob, value = self.select_actions(obs_tensor, legal_mask.to(self.device), is_training=is_training)
        idx, lp, v = int(action.item()), float(log_prob.item()), float(value.item())
        try:
            move = self.policy_output_mapper.policy_index_to_shogi_move(idx)
        except IndexError as e:
            _log("ERROR", f"Policy index {idx} out of bounds in select_action: {e}")
            return None, -1, 0.0, v
        return move, idx, lp, v

    def get_values(self, obs: torch.Tensor, cobs: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.model.eval()
        use_cobs = cobs is not None and not self._is_obs_scaler() and getattr(self.model, "reads_compact_obs", False)
        with torch.no_grad(), self._autocast():
            _, value = self.model(obs, cobs=cobs) if use_cobs else self.model(self._scale(obs))
        if value.dim() > 1 and value.shape[-1] == 1:
            value = value.squeeze(-1)
        return value.float()

    def get_value(self, obs_np: np.ndarray) -> float:
        obs = torch.as_tensor(obs_np, dtype=torch.float32, device=self.device).unsqueeze(0)
        return float(self.get_values(obs).item())

    # ------------------------------------------------------------------ learning
    def learn(self, experience_buffer) -> Dict[str, float]:
        self.model.train()
        batch = experience_buffer.get_batch()
        lr = self.optimizer.param_groups[0]["lr"]
        if not batch or batch["obs"].shape[0] == 0:
            _log("ERROR", "learn called with empty batch_data")
            return {"ppo/policy_loss": 0.0, "ppo/value_loss": 0.0, "ppo/entropy": 0.0,
                    "ppo/kl_divergence_approx": self.last_kl_div, "ppo/learning_rate": lr}
        d = self.device
        obs_b, act_b = batch["obs"].to(d), batch["actions"].to(d)
        oldlp_b, oldv_b = batch["log_probs"].to(d), batch["values"].to(d)
        # legal sets: the reference's byte masks, or the rollout buffer's legal bitmaps (read in place by the kernels)
        mask_b = (batch["legal_bitmaps"] if "legal_bitmaps" in batch else batch["legal_masks"]).to(d)
        adv_b, ret_b = batch["advantages"].to(d), batch["returns"].to(d)
        if self.normalize_advantages:
            adv_b = self._normalize(adv_b)
        n = obs_b.shape[0]
        self._sums = torch.zeros(5, device=d) if getattr(self, "_sums", None) is None else self._sums.zero_()
        if getattr(self, "_gn", None) is None:
            self._gn = torch.zeros((), device=d)
            self._gn2 = torch.zeros(2, device=d)  # {gradient norm, clip coefficient} of the fused optimizer tail
        self._fused_optimizer = (d.type == "cuda" and bool(getattr(self.config.training, "fused_optimizer", True))
                                 and rl.adam_clip_applicable(self.optimizer))
        updates = 0
        in_place = (getattr(getattr(self.model, "module", self.model), "fused_minibatch", False)
                    and not self._is_obs_scaler() and d.type == "cuda")
        S = {"obs": obs_b, "act": act_b, "oldlp": oldlp_b, "oldv": oldv_b, "adv": adv_b, "ret": ret_b, "mask": mask_b}
        if "compact_obs" in batch and getattr(getattr(self.model, "module", self.model), "reads_compact_obs", False):
            S["cobs"] = batch["compact_obs"].to(d)
        graphed = (in_place and self.cuda_graph_update and getattr(self, "_ddp", None) is None and self.scheduler is None
                   and n >= self.minibatch_size)
        if graphed:
            S = self._static_batch(S)
        for _ in range(self.ppo_epochs):
            # minibatch order drawn on the device (the reference shuffles a numpy index array, ppo_agent.py:298)
            indices = torch.randperm(n, device=d, generator=self._device_generator())
            for start in range(0, n, self.minibatch_size):
                mb = indices[start:start + self.minibatch_size]
                if graphed and mb.shape[0] == self.minibatch_size:
                    self._graphed_update(S, mb)
                else:
                    self._minibatch_update(S, mb, in_place)
                if self.scheduler is not None and self.lr_schedule_step_on == "update":
                    self.scheduler.step()
                updates += 1
        if self.scheduler is not None and self.lr_schedule_step_on == "epoch":
            self.scheduler.step()
        avg = (self._sums / max(1, updates)).tolist()  # one host synchronisation per learn() instead of 5 per minibatch
        # the reference reports the norm of the gradients AFTER clipping (ppo_agent.py:395-401 / 409-415)
        gn = float(self._gn) if updates else 0.0
        self.last_gradient_norm = gn * min(1.0, self.gradient_clip_max_norm / (gn + 1e-6))
        self.last_kl_div = avg[3]
        return {"ppo/policy_loss": avg[0], "ppo/value_loss": avg[1], "ppo/entropy": avg[2],
                "ppo/kl_divergence_approx": avg[3], "ppo/clip_fraction": avg[4],
                "ppo/learning_rate": self.optimizer.param_groups[0]["lr"]}

    def _minibatch_update(self, S: Dict[str, torch.Tensor], mb: torch.Tensor, in_place: bool) -> None:
        """One clipped-surrogate update (ppo_agent.py:300-420) entirely on the device: no host synchronisation, so
        the same code runs eagerly or under CUDA-graph capture."""
        with self._autocast():
            if in_place:
                # observations and masks are read in place from the rollout storage through the minibatch indices;
                # input layer and policy head + masked evaluation are fused nodes (nn_ops.py)
                extra = {"cobs": S["cobs"]} if "cobs" in S else {}
                new_lp, entropy, new_v = self._train_forward(S["obs"], rows=mb, actions=S["act"][mb], legal_mask=S["mask"],
                                                             mask_rows=mb, **extra)
            else:
                logits, values = self._train_forward(self._scale(S["obs"][mb]))
                new_lp, entropy, new_v = type(self.model).evaluate_from_logits(logits, values, S["act"][mb], S["mask"],
                                                                               mask_rows=mb)
        new_v = new_v.float()
        old_lp, adv, ret = S["oldlp"][mb], S["adv"][mb], S["ret"][mb]
        grad_world = getattr(self, "_grad_world", 1)
        fused_loss = self.device.type == "cuda" and not self.enable_value_clipping
        if fused_loss:
            # clipped surrogate, value and entropy terms, their gradients and the five metrics in one launch
            loss, stats = rl.ppo_loss(new_lp, entropy, new_v.squeeze(-1) if new_v.dim() > 1 else new_v, old_lp, adv, ret,
                                      self.clip_epsilon, self.value_loss_coeff, self.entropy_coef, 1.0 / grad_world)
        else:
            ratio = torch.exp(new_lp - old_lp)
            policy_loss = -torch.min(ratio * adv, torch.clamp(ratio, 1 - self.clip_epsilon, 1 + self.clip_epsilon) * adv).mean()
            if self.enable_value_clipping:
                old_v = S["oldv"][mb]
                clipped = old_v + torch.clamp(new_v - old_v, -self.clip_epsilon, self.clip_epsilon)
                value_loss = torch.max(F.mse_loss(new_v.squeeze(), ret.squeeze()), F.mse_loss(clipped.squeeze(), ret.squeeze()))
            else:
                value_loss = F.mse_loss(new_v.squeeze(), ret.squeeze())
            entropy_loss = -entropy.mean()
            loss = policy_loss + self.value_loss_coeff * value_loss + self.entropy_coef * entropy_loss
            if grad_world > 1:
                loss = loss / grad_world  # the all-reduce below sums: mean gradient over ranks
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()  # under DistributedDataParallel the gradient all-reduce (NCCL) fires here
        if grad_world > 1:
            # sum over ranks (the loss was scaled by 1 / world): the policy head's weight gradient has been on its way since
            # its GEMM finished inside backward; the small gradients go as one flattened buffer
            self._reducer.finish(self.model.parameters())
        if self._fused_optimizer:
            # global-norm clip + Adam on the optimizer's own state tensors in three launches (csrc/kz_opt.cu)
            gn = rl.adam_clip_step(self.optimizer, self.gradient_clip_max_norm, self._gn2)
        else:
            gn = torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.gradient_clip_max_norm)
            self.optimizer.step()
        with torch.no_grad():
            if fused_loss:
                self._sums += stats[1:]
            else:
                self._sums += torch.stack([policy_loss, value_loss, entropy_loss, (old_lp - new_lp).mean(),
                                           ((ratio - 1.0).abs() > self.clip_epsilon).float().mean()]).detach()
            self._gn.copy_(gn)

    def _static_batch(self, S: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Pin the batch to fixed addresses for graph replay: the rollout storage already is; the normalised
        advantages are a fresh tensor per learn() and are copied into a persistent one.  A different storage (or
        size) invalidates the captured graph."""
        if getattr(self, "_static_adv", None) is None or self._static_adv.shape != S["adv"].shape:
            self._static_adv = torch.empty_like(S["adv"])
        self._static_adv.copy_(S["adv"])
        S = dict(S, adv=self._static_adv)
        key = self._key_of(S)
        if self._graph is not None and key != self._graph_key:
            self._drop_graph()
        if getattr(self, "_static_mb", None) is None or self._static_mb.shape[0] != self.minibatch_size:
            self._static_mb = torch.empty(self.minibatch_size, dtype=torch.int64, device=self.device)
        return S

    def _key_of(self, S: Dict[str, torch.Tensor]):
        # everything the captured graph bakes in: batch storage, optimizer state tensors (load_state_dict swaps them) and
        # the hyper-parameters that travel as kernel scalars (a callback may anneal them)
        g = self.optimizer.param_groups[0]
        opt_state = tuple(t.data_ptr() for p in g["params"] for t in self.optimizer.state.get(p, {}).values()
                          if torch.is_tensor(t))
        scalars = (float(g["lr"]), tuple(g["betas"]), float(g["eps"]), float(g["weight_decay"]), float(self.clip_epsilon),
                   float(self.value_loss_coeff), float(self.entropy_coef), float(self.gradient_clip_max_norm),
                   bool(self.enable_value_clipping), getattr(self, "_grad_world", 1))
        key = (tuple((k, v.data_ptr(), tuple(v.shape), tuple(v.stride())) for k, v in sorted(S.items()))
               + (self.minibatch_size, opt_state, scalars))
        return key

    def _drop_graph(self) -> None:
        """Forget the captured update: its kernels hold raw pointers to the batch, the optimizer state and baked
        scalars, so anything that replaces one of them (load_model, an annealed coefficient, another buffer) must
        come through here."""
        self._graph, self._graph_key, self._graph_warm = None, None, 0

    def _graphed_update(self, S: Dict[str, torch.Tensor], mb: torch.Tensor) -> None:
        """Launch-bound inner loop -> one CUDA graph per minibatch update (~60 kernels, 2.4 ms of device time
        against 6.6 ms of eager launch overhead on B200).  The first three minibatches run eagerly on a side
        stream (allocator / cuBLAS warm-up, as torch.cuda.graphs requires), the fourth is captured, the rest replay."""
        self._static_mb.copy_(mb)
        if self._graph is not None:
            self._graph.replay()
            return
        self._graph_key = self._key_of(S)  # optimizer state tensors come into being during the eager warm-up
        if self._graph_warm < 3:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._minibatch_update(S, self._static_mb, True)
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graph_warm += 1
            return
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            self._minibatch_update(S, self._static_mb, True)
        self._graph = graph
        graph.replay()

    def enable_ddp(self) -> None:
        """Data-parallel update when torch.distributed is initialised (the reference declares a ``ddp`` flag,
        config_schema.py:81, but never wires it).  Models with the fused minibatch path average their gradients
        with explicit all-reduces inside the update -- no wrapper hooks, so the update stays CUDA-graph capturable
        (NCCL collectives are captured like kernels); other models are wrapped in DistributedDataParallel."""
        from ..training import distributed as kd
        world = kd.world()[1]
        if world == 1:
            return
        if getattr(self.model, "fused_minibatch", False):
            from .. import nn_ops
            kd.broadcast_module(self.model)
            self._grad_world = world
            self._reducer = kd.GradReducer()
            nn_ops.grad_reducer = self._reducer
        else:
            self._ddp = kd.wrap_ddp(self.model, self.device)

    def _train_forward(self, obs: torch.Tensor, **kwargs):
        ddp = getattr(self, "_ddp", None)
        return (ddp if ddp is not None else self.model)(obs, **kwargs)

    def _device_generator(self) -> torch.Generator:
        if getattr(self, "_gen", None) is None:
            self._gen = torch.Generator(device=self.device)
            self._gen.manual_seed(int(self._rng.integers(0, 2 ** 62)))
        return self._gen

    def _normalize(self, adv: torch.Tensor) -> torch.Tensor:
        """Whole-buffer normalisation (ppo_agent.py:276-282: unbiased std, skipped for tiny std / single sample).
        When rollouts are sharded over ranks, (count, sum, sum of squares) are all-reduced first."""
        from ..training import distributed as kd
        cnt, s1, s2 = kd.global_moments(adv)
        if cnt <= 1:
            return adv
        mean = s1 / cnt
        var = max(0.0, (s2 - cnt * mean * mean) / (cnt - 1))
        std = var ** 0.5
        return (adv - mean) / std if std > 1e-8 else adv

    # ------------------------------------------------------------------ checkpoints (format kept as is)
    def save_model(self, file_path: str, global_timestep: int = 0, total_episodes_completed: int = 0,
                   stats_to_save: Optional[Dict[str, int]] = None) -> None:
        model = getattr(self.model, "module", self.model)  # unwrap DistributedDataParallel
        # the reference's Adam is not capturable and keeps `step` as a CPU fp32 tensor: write the state that way, so a
        # reference run (CPU included) resumes from this file (torch's Adam asserts capturable state lives on CUDA)
        opt_sd = self.optimizer.state_dict()
        opt_sd = {"state": {k: {n: (t.detach().to("cpu", torch.float32) if n == "step" and torch.is_tensor(t) else t)
                                for n, t in st.items()} for k, st in opt_sd["state"].items()},
                  "param_groups": [dict(g, capturable=False) for g in opt_sd["param_groups"]]}
        data = {"model_state_dict": model.state_dict(), "optimizer_state_dict": opt_sd,
                "global_timestep": global_timestep, "total_episodes_completed": total_episodes_completed}
        if stats_to_save:
            data.update(stats_to_save)
        if self.scheduler is not None:
            data.update(scheduler_state_dict=self.scheduler.state_dict(), lr_schedule_type=self.lr_schedule_type,
                        lr_schedule_step_on=self.lr_schedule_step_on)
        torch.save(data, file_path)

    def load_model(self, file_path: str) -> Dict[str, Any]:
        empty = {"global_timestep": 0, "total_episodes_completed": 0, "black_wins": 0, "white_wins": 0, "draws": 0}
        if not os.path.exists(file_path):
            _log("ERROR", f"Checkpoint file {file_path} not found")
            return {**empty, "error": "File not found"}
        try:
            ck = torch.load(file_path, map_location=self.device, weights_only=False)
            getattr(self.model, "module", self.model).load_state_dict(ck["model_state_dict"])
            self.optimizer.load_state_dict(ck["optimizer_state_dict"])
            for g in self.optimizer.param_groups:  # saved as the reference writes it (capturable False); see save_model
                g["capturable"] = self.device.type == "cuda"
            self._drop_graph()  # a captured update holds pointers to the replaced optimizer state tensors
            if self.scheduler is not None and "scheduler_state_dict" in ck:
                self.scheduler.load_state_dict(ck["scheduler_state_dict"])
            return {k: ck.get(k, v) for k, v in {**empty, "lr_schedule_type": None, "lr_schedule_step_on": "epoch"}.items()}
        except (KeyError, RuntimeError, EOFError) as e:
            _log("ERROR", f"Error loading checkpoint from {file_path}: {e}")
            return {**empty, "error": str(e)}

    def get_name(self) -> str:
        return self.name
