"""Times the UNMODIFIED reference's sequential PPO self-play loop on the host: the reference's
own ShogiGame, PolicyOutputMapper, PPOAgent.select_action, ExperienceBuffer.add / compute_advantages_and_returns and
PPOAgent.learn around keisei.core.neural_network.ActorCritic(46, 13527) -- the composition of StepManager.execute_step
(keisei/training/step_manager.py:98-348) and Trainer.perform_ppo_update (trainer.py:214-269), built directly as
tests/conftest.py:501-525 does because model_factory cannot build ActorCritic.

    python baseline/ref_ppo_loop.py <path to baseline/_ref> <timesteps> <ppo_epochs> <minibatch> <torch threads>

Prints one JSON object.  Run by bench.py (cpu_baseline legs) only; imports nothing from this repository."""
import copy
import json
import sys
import time
from types import SimpleNamespace


class Cfg(SimpleNamespace):
    def model_copy(self, deep=True):
        return copy.deepcopy(self)


def main():
    ref, S, epochs, mb, threads = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    sys.path.insert(0, ref)
    import torch
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    from keisei.core.experience_buffer import ExperienceBuffer
    from keisei.core.neural_network import ActorCritic
    from keisei.core.ppo_agent import PPOAgent
    from keisei.shogi import ShogiGame
    from keisei.utils import PolicyOutputMapper

    cfg = Cfg(env=Cfg(device="cpu", seed=42, input_channels=46, num_actions_total=13527, max_moves_per_game=500),
              training=Cfg(learning_rate=3e-4, gamma=0.99, lambda_gae=0.95, clip_epsilon=0.2, value_loss_coeff=0.5,
                           entropy_coef=0.01, ppo_epochs=epochs, minibatch_size=mb, steps_per_epoch=S, total_timesteps=S * 4,
                           gradient_clip_max_norm=0.5, normalize_advantages=True, enable_value_clipping=False,
                           weight_decay=0.0, lr_schedule_type=None, lr_schedule_step_on="epoch", lr_schedule_kwargs=None))
    dev = torch.device("cpu")
    agent = PPOAgent(ActorCritic(46, 13527), cfg, dev)
    mapper = PolicyOutputMapper()
    buf = ExperienceBuffer(S, 0.99, 0.95, "cpu")
    game = ShogiGame(max_moves_per_game=500)
    obs = game.reset()
    t0 = time.perf_counter()
    for _ in range(S):
        moves = game.get_legal_moves()
        mask = mapper.get_legal_mask(moves, dev)
        move, idx, lp, v = agent.select_action(obs, mask, is_training=True)
        next_obs, reward, done, _ = game.make_move(move)
        buf.add(torch.from_numpy(obs), idx, reward, lp, v, done, mask)
        obs = game.reset() if done else next_obs
    t1 = time.perf_counter()
    buf.compute_advantages_and_returns(agent.get_value(obs))
    metrics = agent.learn(buf)
    t2 = time.perf_counter()
    print(json.dumps({"timesteps": S, "collect_s": t1 - t0, "update_s": t2 - t1, "samples_per_s": S / (t2 - t0),
                      "kl": float(metrics.get("ppo/kl_divergence_approx", 0.0))}))


if __name__ == "__main__":
    main()
