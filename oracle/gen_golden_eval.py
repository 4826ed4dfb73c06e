"""Golden vectors for the evaluation host logic (Elo registry / tracker, tournament standings, ladder opponent
selection), produced by IMPORTING the Python reference in the build container:

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python oracle/gen_golden_eval.py

Test infrastructure only; writes tests/golden/eval_golden.json."""
import json
import logging
import os
import random
import tempfile
from pathlib import Path
from types import SimpleNamespace

from keisei.evaluation.opponents.elo_registry import EloRegistry
from keisei.evaluation.strategies.ladder import EloTracker, LadderEvaluator
from keisei.evaluation.strategies.tournament import TournamentEvaluator

rng = random.Random(20261018)
RES = ["agent_win", "opponent_win", "draw"]
out = {"registry": [], "tracker": [], "standings": [], "ladder": []}

# EloRegistry.update_ratings: chains of matches over a few players
for case in range(6):
    tmp = Path(tempfile.mkdtemp()) / "elo.json"
    reg = EloRegistry(tmp, initial_rating=rng.choice([1500.0, 1200.0]), k_factor=rng.choice([32.0, 16.0, 24.5]))
    matches = []
    for _ in range(rng.randint(1, 8)):
        p1, p2 = rng.sample(["agent", "a", "b", "c"], 2)
        results = [rng.choice(RES) for _ in range(rng.randint(0, 9))]
        reg.update_ratings(p1, p2, results)
        matches.append([p1, p2, results])
    reg.save()
    out["registry"].append({"initial_rating": reg.initial_rating, "k_factor": reg.k_factor, "matches": matches,
                            "ratings": reg.get_all_ratings(), "top2": reg.get_top_players(2),
                            "file": json.loads(tmp.read_text())})

# EloTracker.update_ratings: game-by-game inside a match
for case in range(6):
    tr = EloTracker()
    matches = []
    for _ in range(rng.randint(1, 6)):
        opp = rng.choice(["a", "b", "c"])
        results = [rng.choice(RES) for _ in range(rng.randint(1, 12))]
        games = [SimpleNamespace(winner=0 if r == "agent_win" else 1 if r == "opponent_win" else None) for r in results]
        tr.update_ratings("agent", opp, games)
        matches.append([opp, results])
    out["tracker"].append({"matches": matches, "ratings": tr.get_elo_snapshot()})

# TournamentEvaluator._calculate_tournament_standings (uses no instance state)
for case in range(4):
    names = ["opp%d" % i for i in range(rng.randint(1, 4))]
    per = {}
    games = []
    for nm in names:
        w, l, d = rng.randint(0, 7), rng.randint(0, 7), rng.randint(0, 3)
        per[nm] = [w, l, d]
        info = SimpleNamespace(name=nm)
        games += [SimpleNamespace(winner=0, opponent_info=info)] * w + [SimpleNamespace(winner=1, opponent_info=info)] * l \
            + [SimpleNamespace(winner=None, opponent_info=info)] * d
    st = TournamentEvaluator._calculate_tournament_standings(None, games, [SimpleNamespace(name=n) for n in names], None)
    out["standings"].append({"per_opponent": per, "standings": st})

# LadderEvaluator._select_ladder_opponents
for case in range(6):
    pool = [SimpleNamespace(name="p%d" % i, metadata={"initial_rating": rng.choice([900, 1100, 1250, 1500, 1500, 1700, 1900, 2100])})
            for i in range(rng.randint(0, 9))]
    num = rng.choice([1, 3, 5])
    fake = SimpleNamespace(opponent_pool=pool, config=SimpleNamespace(num_opponents_to_select=num),
                           logger=logging.getLogger("gen"))
    rating = rng.choice([1500.0, 1320.5, 1710.0])
    sel = LadderEvaluator._select_ladder_opponents(fake, rating, None)
    out["ladder"].append({"agent_rating": rating, "num": num, "pool": [[p.name, p.metadata["initial_rating"]] for p in pool],
                          "selected": [o.name for o in sel]})

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "eval_golden.json")
with open(dst, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", dst, {k: len(v) for k, v in out.items()})
