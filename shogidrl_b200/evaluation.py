"""Batched evaluation games on the device engine.

The reference's evaluation strategies each run a scalar ``get_legal_moves -> get_legal_mask -> select_action ->
make_move`` loop per game (keisei/evaluation/strategies/single_opponent.py:162-221).  Here N games are played at
once on a VecShogiEnv: the agent moves for one colour, the opponent (uniform-random legal, the engine's fused
policy, or a second agent) for the other; games that finish are reset and keep counting until ``num_games`` have
been completed."""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, List, Mapping, Optional, Tuple

import torch

from .vec_env import VecShogiEnv

OUTCOMES = ("draw", "agent_win", "opponent_win")  # the reference's result strings (elo_registry.py:86)


@dataclass
class EvaluationResult:
    games: int
    agent_wins: int
    opponent_wins: int
    draws: int
    mean_length: float
    outcomes: List[str] = field(default_factory=list)  # one of OUTCOMES per finished game, in completion order

    @property
    def win_rate(self) -> float:
        return self.agent_wins / max(1, self.games)


def evaluate_vs_opponent(agent, num_games: int, *, opponent=None, num_envs: int = 1024, max_moves_per_game: int = 500,
                         device="cuda", seed: int = 0, deterministic: bool = True,
                         max_steps: Optional[int] = None) -> EvaluationResult:
    """Play ``num_games`` games of ``agent`` against ``opponent`` (None = uniform-random legal moves).  The agent
    plays Black in even-numbered envs and White in odd-numbered ones.  Every env has a fixed QUOTA of games
    (``num_games`` spread evenly over the envs): the first ``quota`` games an env finishes are counted and later ones
    are ignored, so the sample is ``num_games`` independent games exactly as the reference's sequential loop plays them
    -- counting "the first num_games completions of the batch" would over-sample short (decisive) games, because a
    fast env restarts and finishes again while the long draws are still running.  Every tensor stays on the device;
    the host reads four counters per step."""
    n = min(num_envs, max(2, num_games))
    env = VecShogiEnv(n, max_moves_per_game=max_moves_per_game, device=device, seed=seed, auto_reset=True)
    dev = env.device
    agent_is_black = (torch.arange(n, device=dev) % 2) == 0
    quota = torch.full((n,), num_games // n, dtype=torch.int64, device=dev)
    quota[: num_games % n] += 1
    completed = torch.zeros(n, dtype=torch.int64, device=dev)
    env.refresh(random_actions=True)
    # side to move per env: read from plane 42 of the observation (1.0 = Black to move)
    done_games = agent_w = opp_w = draws = 0
    length_sum = 0
    outcomes: List[str] = []
    steps = 0
    limit = max_steps if max_steps is not None else (max_moves_per_game + 1) * (1 + num_games // n) + 8
    while done_games < num_games and steps < limit:
        black_to_move = env.obs[:, 42, 0, 0] > 0.5
        agent_turn = black_to_move == agent_is_black
        a_agent, _, _ = agent.select_actions(env.obs, env.mask, is_training=not deterministic)
        if opponent is None:
            a_opp = env.next_actions
        else:
            a_opp, _, _ = opponent.select_actions(env.obs, env.mask, is_training=not deterministic)
        actions = torch.where(agent_turn, a_agent, a_opp).contiguous()
        out = env.step(actions, random_actions=opponent is None)
        steps += 1
        d = (out["done"] != 0) & (completed < quota)  # finished games that still count towards their env's quota
        completed += d
        if bool(d.any()):
            w = out["winner"]
            agent_won = d & (((w == 0) & agent_is_black) | ((w == 1) & ~agent_is_black))
            opp_won = d & (w >= 0) & ~agent_won
            codes = (agent_won.long() + 2 * opp_won.long())[d]  # index into OUTCOMES, env order
            stats = torch.cat([torch.stack([d.sum(), agent_won.sum(), opp_won.sum(), (out["ep_len"] * d).sum()]),
                               codes]).tolist()  # one device-to-host read per step that finished a game
            done_games += stats[0]; agent_w += stats[1]; opp_w += stats[2]; length_sum += stats[3]
            draws += stats[0] - stats[1] - stats[2]
            outcomes.extend(OUTCOMES[c] for c in stats[4:])
    return EvaluationResult(done_games, agent_w, opp_w, draws, length_sum / max(1, done_games), outcomes)


# ---------------------------------------------------------------------------------------------------------------
# Tournament and ladder evaluation on top of evaluate_vs_opponent (keisei/evaluation/strategies/tournament.py,
# ladder.py): the scheduling and the rating arithmetic are host-side and follow the reference; the games run batched.

class EloRegistry:
    """keisei/evaluation/opponents/elo_registry.py:15-135: match-level Elo update, ratings persisted as JSON in the
    reference's layout ({"ratings": {...}, "metadata": {"initial_rating", "k_factor"}})."""

    def __init__(self, file_path: Optional[Path] = None, initial_rating: float = 1500.0, k_factor: float = 32.0):
        self.file_path = Path(file_path) if file_path is not None else None
        self.initial_rating = initial_rating
        self.k_factor = k_factor
        self.ratings: Dict[str, float] = {}
        self.load()

    def load(self) -> None:
        if self.file_path is not None and self.file_path.exists():
            try:
                with open(self.file_path, "r", encoding="utf-8") as f:
                    self.ratings = {k: float(v) for k, v in json.load(f).get("ratings", {}).items()}
            except (OSError, ValueError):
                self.ratings = {}

    def save(self) -> None:
        if self.file_path is None:
            return
        self.file_path.parent.mkdir(parents=True, exist_ok=True)
        with open(self.file_path, "w", encoding="utf-8") as f:
            json.dump({"ratings": self.ratings,
                       "metadata": {"initial_rating": self.initial_rating, "k_factor": self.k_factor}}, f, indent=2)

    def get_rating(self, player_id: str) -> float:
        return self.ratings.setdefault(player_id, self.initial_rating)

    def set_rating(self, player_id: str, rating: float) -> None:
        self.ratings[player_id] = rating

    def update_ratings(self, player1_id: str, player2_id: str, results: List[str]) -> None:
        """One update per match from the mean score of player 1 (elo_registry.py:77-121)."""
        if not results:
            return
        r1, r2 = self.get_rating(player1_id), self.get_rating(player2_id)
        score1 = 0.0
        for res in results:
            if res == "agent_win":
                score1 += 1.0
            elif res == "draw":
                score1 += 0.5
        score2 = len(results) - score1
        expected1 = 1.0 / (1.0 + 10.0 ** ((r2 - r1) / 400.0))
        expected2 = 1.0 - expected1
        self.set_rating(player1_id, r1 + self.k_factor * (score1 / len(results) - expected1))
        self.set_rating(player2_id, r2 + self.k_factor * (score2 / len(results) - expected2))

    def get_all_ratings(self) -> Dict[str, float]:
        return self.ratings.copy()

    def get_top_players(self, limit: int = 10) -> List[Tuple[str, float]]:
        return sorted(self.ratings.items(), key=lambda kv: kv[1], reverse=True)[:limit]


class EloTracker:
    """keisei/evaluation/strategies/ladder.py:54-97: game-by-game Elo update inside one match (the expectation is
    recomputed after every game, so the order of the results matters)."""

    def __init__(self, default_rating: float = 1500.0, k_factor: float = 32):
        self.ratings: Dict[str, float] = {}
        self.default_rating = default_rating
        self.k_factor = k_factor

    def get_agent_rating(self, agent_id: str) -> float:
        return self.ratings.get(agent_id, self.default_rating)

    def update_ratings(self, agent_id: str, opponent_id: str, results: List[str]) -> None:
        a, o = self.get_agent_rating(agent_id), self.get_agent_rating(opponent_id)
        for res in results:
            expected = 1 / (1 + 10 ** ((o - a) / 400))
            actual = 1.0 if res == "agent_win" else 0.5 if res == "draw" else 0.0
            da = self.k_factor * (actual - expected)
            do = self.k_factor * ((1 - actual) - (1 - expected))
            a += da
            o += do
        self.ratings[agent_id] = a
        self.ratings[opponent_id] = o

    def get_elo_snapshot(self) -> Dict[str, float]:
        return self.ratings.copy()


def tournament_standings(results: Mapping[str, EvaluationResult]) -> Dict[str, Any]:
    """The reference's tournament analytics (tournament.py:631-696) from per-opponent match results."""
    overall = {"total_games": 0, "agent_total_wins": 0, "agent_total_losses": 0, "agent_total_draws": 0,
               "agent_overall_win_rate": 0.0}
    per: Dict[str, Dict[str, Any]] = {}
    for name, r in results.items():
        per[name] = {"played": r.games, "wins": r.agent_wins, "losses": r.opponent_wins, "draws": r.draws,
                     "win_rate": r.agent_wins / r.games if r.games > 0 else 0.0}
        overall["total_games"] += r.games
        overall["agent_total_wins"] += r.agent_wins
        overall["agent_total_losses"] += r.opponent_wins
        overall["agent_total_draws"] += r.draws
    if overall["total_games"] > 0:
        overall["agent_overall_win_rate"] = overall["agent_total_wins"] / overall["total_games"]
    return {"overall_tournament_stats": overall, "per_opponent_results": per}


def evaluate_tournament(agent, opponents: Mapping[str, Any], num_games_per_opponent: int,
                        **kwargs) -> Tuple[Dict[str, Any], Dict[str, EvaluationResult]]:
    """Round of matches against every opponent of the pool (name -> agent-like object, or None for the uniform-random
    player), colours balanced inside each match; returns (standings, per-opponent results)."""
    results = {name: evaluate_vs_opponent(agent, num_games_per_opponent, opponent=opp, **kwargs)
               for name, opp in opponents.items()}
    return tournament_standings(results), results


def benchmark_performance(results: Mapping[str, EvaluationResult]) -> Dict[str, Any]:
    """The reference's benchmark analytics (keisei/evaluation/strategies/benchmark.py:639-686) from per-case match results:
    a case is "passed" by a game the agent wins; cases without games report the reference's empty record."""
    per: Dict[str, Dict[str, Any]] = {}
    for name, r in results.items():
        if r.games == 0:
            per[name] = {"played": 0, "wins_or_passes": 0, "pass_rate": 0, "details": "No games played"}
        else:
            per[name] = {"played": r.games, "wins_or_passes": r.agent_wins, "pass_rate": r.agent_wins / r.games}
    played = sum(p["played"] for p in per.values())
    overall = sum(p["wins_or_passes"] for p in per.values()) / played if played else 0
    return {"per_benchmark_case_results": per, "overall_benchmark_pass_rate": overall}


def evaluate_benchmark(agent, suite: Mapping[str, Any], num_games_per_case: int,
                       **kwargs) -> Tuple[Dict[str, Any], Dict[str, EvaluationResult]]:
    """The benchmark strategy (benchmark.py:330-426, 556-637) on the vectorised engine: ``num_games_per_case`` games against
    every case of the suite (name -> opponent agent, or None for the uniform-random baseline the reference's default suite
    starts with), colours balanced; returns (the reference's performance dictionary, per-case results)."""
    results = {name: evaluate_vs_opponent(agent, num_games_per_case, opponent=opp, **kwargs) for name, opp in suite.items()}
    return benchmark_performance(results), results


def select_ladder_opponents(agent_rating: float, pool_ratings: Mapping[str, float], num_opponents_to_select: int = 5,
                            window: float = 400.0) -> List[str]:
    """ladder.py:700-732: opponents rated within +-400 of the agent, ascending by rating, the first N."""
    near = [(r, i, name) for i, (name, r) in enumerate(pool_ratings.items())
            if agent_rating - window <= r <= agent_rating + window]
    near.sort(key=lambda t: (t[0], t[1]))  # stable: pool order breaks rating ties, as list.sort does in the reference
    return [name for _, _, name in near[:num_opponents_to_select]]


def evaluate_ladder(agent, agent_id: str, pool: Mapping[str, Any], tracker: Optional[EloTracker] = None,
                    num_games_per_match: int = 2, num_opponents_to_select: int = 5,
                    **kwargs) -> Tuple[Dict[str, float], Dict[str, EvaluationResult]]:
    """One ladder pass (ladder.py:490-558): select opponents near the agent's rating, play a colour-balanced match
    against each, update the ratings game by game; returns (rating snapshot, per-opponent results)."""
    tracker = tracker if tracker is not None else EloTracker()
    ratings = {name: tracker.get_agent_rating(name) for name in pool}
    chosen = select_ladder_opponents(tracker.get_agent_rating(agent_id), ratings, num_opponents_to_select)
    results: Dict[str, EvaluationResult] = {}
    for name in chosen:
        r = evaluate_vs_opponent(agent, num_games_per_match, opponent=pool[name], **kwargs)
        tracker.update_ratings(agent_id, name, r.outcomes)
        results[name] = r
    return tracker.get_elo_snapshot(), results
