"""Golden trace of the reference's per-timestep composition (StepManager.execute_step /
handle_episode_end / update_episode_state around ShogiGame, PolicyOutputMapper and ExperienceBuffer), produced by
IMPORTING the reference:

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python oracle/gen_golden_stepmanager.py

A scripted agent (k-th legal action by the counter RNG, log-prob -ln(n), value n/100) drives the reference's real
objects for STEPS timesteps with max_moves_per_game = MAX_MOVES, crossing several episode ends.  Recorded: every
StepResult, the per-episode counters, EpisodeState after each update, the episode-end log line and W&B payload, and
the buffer contents.  Test infrastructure only; writes tests/golden/stepmanager_golden.json."""
from __future__ import annotations

import json
import math
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import rand32  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED, STEPS, MAX_MOVES = 777, 150, 40
COUNTERS = ("sente_capture_count", "gote_capture_count", "sente_drop_count", "gote_drop_count", "sente_promo_count",
            "gote_promo_count", "sente_best_capture", "gote_best_capture")


class ScriptedAgent:
    """select_action of the PPOAgent protocol without a network: deterministic in (seed, call index, legal mask)."""

    def __init__(self, mapper, seed: int = SEED):
        self.mapper, self.seed, self.calls = mapper, seed, 0

    def select_action(self, obs, legal_mask, *, is_training=True):
        legal = legal_mask.nonzero().flatten().tolist()
        idx = legal[(rand32(self.seed, 0, self.calls) * len(legal)) >> 32]
        self.calls += 1
        return self.mapper.policy_index_to_shogi_move(idx), idx, -math.log(len(legal)), len(legal) / 100.0


def make_config(device: str):
    return SimpleNamespace(env=SimpleNamespace(device=device, seed=42, input_channels=46, num_actions_total=13527,
                                               max_moves_per_game=MAX_MOVES),
                           training=SimpleNamespace(gamma=0.99, lambda_gae=0.95),
                           display=SimpleNamespace(display_moves=False, turn_tick=0.0))


def drive(StepManager, ShogiGame, PolicyOutputMapper, ExperienceBuffer, device: str):
    """The loop of TrainingLoopManager._process_step_and_handle_episode (training_loop_manager.py:522-526) over the
    given classes; returns a JSON-able record.  Used by the generator (reference classes) and by the test (ours)."""
    cfg = make_config(device)
    game = ShogiGame(max_moves_per_game=MAX_MOVES)
    mapper = PolicyOutputMapper()
    buf = ExperienceBuffer(STEPS, 0.99, 0.95, device)
    sm = StepManager(cfg, game, ScriptedAgent(mapper), mapper, buf)
    logs = []

    def logger(msg, also_to_wandb=False, wandb_data=None, log_level="info"):
        logs.append({"msg": msg, "wandb": wandb_data, "level": log_level})

    state = sm.reset_episode()
    stats = {"black_wins": 0, "white_wins": 0, "draws": 0}
    episodes = 0
    steps, ends = [], []
    for t in range(STEPS):
        r = sm.execute_step(state, t, logger)
        rec = {"policy_index": int(r.policy_index), "reward": float(r.reward), "done": bool(r.done), "success": bool(r.success),
               "info": {k: (v if isinstance(v, (str, int, float, bool, type(None))) else str(v)) for k, v in r.info.items()},
               "move": [None if x is None else (bool(x) if isinstance(x, bool) else int(getattr(x, "value", x)))
                        for x in r.selected_move] if r.selected_move is not None else None,
               "obs_sum": float(r.next_obs.sum(dtype="float64")), "log_prob": float(r.log_prob), "value": float(r.value_pred),
               "counters": {c: getattr(sm, c) for c in COUNTERS}}
        if r.done:
            state, winner = sm.handle_episode_end(state, r, stats, episodes, logger)
            if winner == "black":
                stats["black_wins"] += 1
            elif winner == "white":
                stats["white_wins"] += 1
            else:
                stats["draws"] += 1
            episodes += 1
            ends.append({"t": t, "winner": winner, "stats": dict(stats), "log": logs[-1]})
        else:
            state = sm.update_episode_state(state, r)
        rec["episode"] = [float(state.episode_reward), int(state.episode_length)]
        steps.append(rec)
    n = len(buf) if hasattr(buf, "__len__") else buf.ptr
    return {"steps": steps, "ends": ends, "n_logs": len(logs),
            "buffer": {"n": int(n), "actions": [int(x) for x in buf.actions[:n].tolist()],
                       "rewards": [float(x) for x in buf.rewards[:n].tolist()],
                       "dones": [bool(x) for x in buf.dones[:n].tolist()],
                       "log_probs": [float(x) for x in buf.log_probs[:n].tolist()],
                       "values": [float(x) for x in buf.values[:n].tolist()],
                       "mask_sums": [int(x) for x in buf.legal_masks[:n].sum(1).tolist()],
                       "obs_sums": [float(x) for x in buf.obs[:n].double().sum((1, 2, 3)).tolist()]}}


def main():
    from keisei.core.experience_buffer import ExperienceBuffer
    from keisei.shogi import ShogiGame
    from keisei.training.step_manager import StepManager
    from keisei.utils import PolicyOutputMapper

    rec = drive(StepManager, ShogiGame, PolicyOutputMapper, ExperienceBuffer, "cpu")
    with open(os.path.join(GOLD, "stepmanager_golden.json"), "w") as f:
        json.dump(rec, f, indent=0)
    print(len(rec["steps"]), "steps,", len(rec["ends"]), "episode ends:", [(e["t"], e["winner"], e["log"]["msg"]) for e in rec["ends"]])
    print("final counters", rec["steps"][-1]["counters"], "drops", sum(1 for s in rec["steps"] if s["move"] and s["move"][0] is None))


if __name__ == "__main__":
    main()
