"""Host logic of the tournament / ladder evaluation against golden vectors produced by importing the
reference (oracle/gen_golden_eval.py -> tests/golden/eval_golden.json).  Rating arithmetic is float64 and bit-exact."""
import json
import os

import pytest

from shogidrl_b200.evaluation import (EloRegistry, EloTracker, EvaluationResult, benchmark_performance,
                                      select_ladder_opponents, tournament_standings)


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(os.path.dirname(__file__), "golden", "eval_golden.json")) as f:
        return json.load(f)


def test_elo_registry_matches_reference(golden, tmp_path):
    for i, case in enumerate(golden["registry"]):
        path = tmp_path / f"elo{i}.json"
        reg = EloRegistry(path, initial_rating=case["initial_rating"], k_factor=case["k_factor"])
        for p1, p2, results in case["matches"]:
            reg.update_ratings(p1, p2, results)
        assert reg.get_all_ratings() == case["ratings"]  # exact float64 equality
        assert [list(t) for t in reg.get_top_players(2)] == case["top2"]
        reg.save()
        assert json.loads(path.read_text()) == case["file"]
        again = EloRegistry(path)  # round trip through the reference's JSON layout
        assert again.get_all_ratings() == case["ratings"]


def test_elo_tracker_matches_reference(golden):
    for case in golden["tracker"]:
        tr = EloTracker()
        for opp, results in case["matches"]:
            tr.update_ratings("agent", opp, results)
        assert tr.get_elo_snapshot() == case["ratings"]


def test_tournament_standings_match_reference(golden):
    for case in golden["standings"]:
        results = {name: EvaluationResult(w + l + d, w, l, d, 0.0) for name, (w, l, d) in case["per_opponent"].items()}
        assert tournament_standings(results) == case["standings"]


def test_ladder_selection_matches_reference(golden):
    for case in golden["ladder"]:
        pool = {name: rating for name, rating in case["pool"]}
        assert select_ladder_opponents(case["agent_rating"], pool, case["num"]) == case["selected"]
    assert select_ladder_opponents(1500.0, {}) == []


def test_benchmark_performance_schema():
    """benchmark.py:639-686: per-case played / wins_or_passes / pass_rate, the empty-case record, the overall pass rate."""
    res = {"benchmark_random": EvaluationResult(10, 7, 2, 1, 0.0), "benchmark_heuristic": EvaluationResult(4, 1, 3, 0, 0.0),
           "unplayed": EvaluationResult(0, 0, 0, 0, 0.0)}
    perf = benchmark_performance(res)
    assert perf["per_benchmark_case_results"]["benchmark_random"] == {"played": 10, "wins_or_passes": 7, "pass_rate": 0.7}
    assert perf["per_benchmark_case_results"]["unplayed"] == {"played": 0, "wins_or_passes": 0, "pass_rate": 0,
                                                               "details": "No games played"}
    assert perf["overall_benchmark_pass_rate"] == 8 / 14
    assert benchmark_performance({})["overall_benchmark_pass_rate"] == 0
