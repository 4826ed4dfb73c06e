"""Golden vectors for the agent side of the path, produced by IMPORTING the Python
reference on CPU in fp32:

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python oracle/gen_golden_agent.py

The reference's ActorCritic gets deterministic weights (an integer hash of the element index -- no RNG stream to match),
plays one game with PPOAgent.select_action into an ExperienceBuffer, computes GAE and runs PPOAgent.learn with one
minibatch per epoch (so the minibatch order cannot matter).  Stored: the buffer contents (the inputs), the
deterministic-mode actions / log-probs / values, evaluate_actions outputs, advantages / returns, learn()'s metrics and
the parameter changes.  Test infrastructure only; writes tests/golden/agent_golden.npz."""
from __future__ import annotations

import copy
import os
import zlib
from types import SimpleNamespace

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
N_STEPS = 64
SHAPES = {"conv.weight": (16, 46, 3, 3), "conv.bias": (16,), "policy_head.weight": (13527, 1296),
          "policy_head.bias": (13527,), "value_head.weight": (1, 1296), "value_head.bias": (1,)}
SCALES = {"conv.weight": 0.05, "conv.bias": 0.01, "policy_head.weight": 0.028, "policy_head.bias": 0.01,
          "value_head.weight": 0.028, "value_head.bias": 0.01}
TRAINING = dict(learning_rate=3e-4, gamma=0.99, lambda_gae=0.95, clip_epsilon=0.2, value_loss_coeff=0.5,
                entropy_coef=0.01, ppo_epochs=2, minibatch_size=N_STEPS, steps_per_epoch=N_STEPS, total_timesteps=1024,
                gradient_clip_max_norm=0.5, normalize_advantages=True, enable_value_clipping=False, weight_decay=0.0,
                lr_schedule_type=None, lr_schedule_step_on="epoch", lr_schedule_kwargs=None)


def det_tensor(name: str, shape, scale: float) -> np.ndarray:
    """Deterministic pseudo-random fp32 tensor in [-scale, scale): exact integer arithmetic, then one rounding."""
    n = int(np.prod(shape))
    x = (np.arange(n, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(zlib.crc32(name.encode()))) % np.uint64(1 << 32)
    x = ((x ^ (x >> np.uint64(15))) * np.uint64(2246822519)) % np.uint64(1 << 32)
    x = x ^ (x >> np.uint64(13))
    return ((x.astype(np.float64) / float(1 << 32) - 0.5) * 2.0 * scale).astype(np.float32).reshape(shape)


def det_state_dict():
    import torch
    return {k: torch.from_numpy(det_tensor(k, s, SCALES[k])) for k, s in SHAPES.items()}


class Cfg(SimpleNamespace):
    def model_copy(self, deep=True):
        return copy.deepcopy(self)


def main():
    import torch
    from keisei.core.experience_buffer import ExperienceBuffer
    from keisei.core.neural_network import ActorCritic
    from keisei.core.ppo_agent import PPOAgent
    from keisei.shogi import ShogiGame
    from keisei.utils import PolicyOutputMapper

    torch.manual_seed(7)
    torch.set_num_threads(1)
    cfg = Cfg(env=Cfg(device="cpu", seed=42, input_channels=46, num_actions_total=13527, max_moves_per_game=500),
              training=Cfg(**TRAINING))
    model = ActorCritic(46, 13527)
    model.load_state_dict(det_state_dict())
    agent = PPOAgent(model, cfg, torch.device("cpu"))
    mapper = PolicyOutputMapper()
    buf = ExperienceBuffer(N_STEPS, TRAINING["gamma"], TRAINING["lambda_gae"], "cpu")
    game = ShogiGame(max_moves_per_game=500)
    obs = game.reset()
    legal_lists = []
    for t in range(N_STEPS):
        moves = game.get_legal_moves()
        mask = mapper.get_legal_mask(moves, torch.device("cpu"))
        move, idx, lp, v = agent.select_action(obs, mask, is_training=True)
        next_obs, reward, done, _ = game.make_move(move)
        if t == 20:
            reward, done = 1.0, True  # a synthetic episode boundary inside the buffer for the GAE scan
        buf.add(torch.from_numpy(obs), idx, reward, lp, v, done, mask)
        legal_lists.append(np.nonzero(mask.numpy())[0].astype(np.uint16))
        obs = game.reset() if done else next_obs
    last_value = agent.get_value(obs)
    buf.compute_advantages_and_returns(last_value)
    batch = buf.get_batch()
    model.eval()
    with torch.no_grad():
        det_a, det_lp, det_v = model.get_action_and_value(batch["obs"], batch["legal_masks"], deterministic=True)
        ev_lp, ev_ent, ev_v = model.evaluate_actions(batch["obs"], batch["actions"], batch["legal_masks"])
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    metrics = agent.learn(buf)
    after = model.state_dict()
    out = dict(n=np.int32(N_STEPS), obs=batch["obs"].numpy(), actions=batch["actions"].numpy().astype(np.int64),
               rewards=batch["rewards"].numpy() if "rewards" in batch else buf.rewards[:N_STEPS].numpy(),
               log_probs=batch["log_probs"].numpy(), values=batch["values"].numpy(),
               dones=batch["dones"].numpy().astype(np.uint8), last_value=np.float32(last_value),
               legal=np.concatenate(legal_lists), legal_off=np.concatenate([[0], np.cumsum([len(x) for x in legal_lists])]),
               advantages=batch["advantages"].numpy(), returns=batch["returns"].numpy(),
               det_actions=det_a.numpy().astype(np.int64), det_log_probs=det_lp.numpy(), det_values=det_v.reshape(-1).numpy(),
               ev_log_probs=ev_lp.numpy(), ev_entropy=ev_ent.numpy(), ev_values=ev_v.reshape(-1).numpy(),
               metric_names=np.asarray(sorted(metrics)), metric_values=np.asarray([metrics[k] for k in sorted(metrics)], np.float64),
               last_gradient_norm=np.float64(agent.last_gradient_norm))
    rng = np.random.default_rng(3)
    for k in SHAPES:
        delta = (after[k] - before[k]).numpy().reshape(-1)
        nz = np.nonzero(delta)[0]
        pick = np.sort(rng.choice(nz, size=min(2048, len(nz)), replace=False)) if len(nz) else np.zeros(0, np.int64)
        out[f"delta_idx/{k}"] = pick.astype(np.int64)
        out[f"delta_val/{k}"] = delta[pick]
        out[f"delta_stats/{k}"] = np.asarray([len(nz), np.abs(delta).sum(dtype=np.float64), delta.sum(dtype=np.float64)])
    np.savez_compressed(os.path.join(GOLD, "agent_golden.npz"), **out)
    print({k: float(v) for k, v in metrics.items()}, "grad norm", agent.last_gradient_norm,
          {k: out[f"delta_stats/{k}"].tolist() for k in SHAPES})


if __name__ == "__main__":
    main()
