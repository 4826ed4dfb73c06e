"""Batched evaluation games on the device engine.

The reference's evaluation strategies each run a scalar ``get_legal_moves -> get_legal_mask -> select_action ->
make_move`` loop per game (keisei/evaluation/strategies/single_opponent.py:162-221).  Here N games are played at
once on a VecShogiEnv: the agent moves for one colour, the opponent (uniform-random legal, the engine's fused
policy, or a second agent) for the other; games that finish are reset and keep counting until ``num_games`` have
been completed."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from .vec_env import VecShogiEnv


@dataclass
class EvaluationResult:
    games: int
    agent_wins: int
    opponent_wins: int
    draws: int
    mean_length: float

    @property
    def win_rate(self) -> float:
        return self.agent_wins / max(1, self.games)


def evaluate_vs_opponent(agent, num_games: int, *, opponent=None, num_envs: int = 1024, max_moves_per_game: int = 500,
                         device="cuda", seed: int = 0, deterministic: bool = True,
                         max_steps: Optional[int] = None) -> EvaluationResult:
    """Play ``num_games`` games of ``agent`` against ``opponent`` (None = uniform-random legal moves).  The agent
    plays Black in even-numbered envs and White in odd-numbered ones.  Every tensor stays on the device; the host
    reads four counters per step."""
    n = min(num_envs, max(2, num_games))
    env = VecShogiEnv(n, max_moves_per_game=max_moves_per_game, device=device, seed=seed, auto_reset=True)
    dev = env.device
    agent_is_black = (torch.arange(n, device=dev) % 2) == 0
    env.refresh(random_actions=True)
    # side to move per env: read from plane 42 of the observation (1.0 = Black to move)
    done_games = agent_w = opp_w = draws = 0
    length_sum = 0
    steps = 0
    limit = max_steps if max_steps is not None else 4 * max_moves_per_game * (1 + num_games // n)
    while done_games < num_games and steps < limit:
        black_to_move = env.obs[:, 42, 0, 0] > 0.5
        agent_turn = black_to_move == agent_is_black
        a_agent, _, _ = agent.select_actions(env.obs, env.mask, is_training=not deterministic)
        if opponent is None:
            a_opp = env.next_actions
        else:
            a_opp, _, _ = opponent.select_actions(env.obs, env.mask, is_training=not deterministic)
        actions = torch.where(agent_turn, a_agent, a_opp).contiguous()
        mover_black = black_to_move.clone()
        out = env.step(actions, random_actions=opponent is None)
        steps += 1
        d = out["done"] != 0
        if bool(d.any()):
            w = out["winner"]
            agent_won = d & (((w == 0) & agent_is_black) | ((w == 1) & ~agent_is_black))
            opp_won = d & (w >= 0) & ~agent_won
            stats = torch.stack([d.sum(), agent_won.sum(), opp_won.sum(), (out["ep_len"] * d).sum()]).tolist()
            done_games += stats[0]; agent_w += stats[1]; opp_w += stats[2]; length_sum += stats[3]
            draws += stats[0] - stats[1] - stats[2]
        del mover_black
    return EvaluationResult(done_games, agent_w, opp_w, draws, length_sum / max(1, done_games))
