"""Soak differential (test infrastructure, not shipped): the C oracle against the IMPORTED Python reference, ply by ply,
on fresh random games -- from the start position and from drop-heavy endgames loaded through SFEN -- for as long as asked.
Compared at every ply: the legal action set, and after the move reward / done / winner / reason, the 46-plane observation
and the position (SFEN of the reference vs the oracle's export).  Needs the reference checkout:

    python oracle/soak_vs_reference.py --minutes 20 --procs 8 [--ref /root/reference] [--seed 1] [--sparse]

Prints one summary line per process and exits non-zero on the first divergence (with the seed that reproduces it)."""
import argparse
import multiprocessing as mp
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REASONS = {"Tsumi": 1, "stalemate": 2, "Max moves reached": 3, "Sennichite": 4}  # KZ_* codes of include/keisei_b200.h


def _sparse_positions(n, seed):
    """Kings and at most three non-pawn pieces, empty hands: random play from these repeats positions (sennichite)."""
    import numpy as np
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    boards, hands, sides = [], [], []
    while len(boards) < n:
        b = np.zeros(81, np.int8)
        k0, k1 = rng.choice(81, 2, replace=False)
        if max(abs(k0 // 9 - k1 // 9), abs(k0 % 9 - k1 % 9)) < 2:
            continue
        b[k0], b[k1] = 8, 22
        for _ in range(int(rng.integers(0, 4))):
            sq, t, color = int(rng.integers(0, 81)), int(rng.choice([3, 4, 5, 6, 12, 13])), int(rng.integers(0, 2))
            if b[sq] == 0:
                b[sq] = 1 + t + 14 * color
        side = int(rng.integers(0, 2))
        h = np.zeros(14, np.uint8)
        g = orc.OracleGame.from_arrays(b, h, side, 0, 500, evaluate_termination=False)
        if g.in_check(1 - side) or len(g.legal_indices()) == 0:
            continue
        boards.append(b); hands.append(h); sides.append(side)
    return np.stack(boards), np.stack(hands), np.asarray(sides, np.uint8)


def _worker(args):
    ref_dir, seed, deadline, max_moves, sparse = args
    sys.dont_write_bytecode = True
    sys.path.insert(0, ROOT)
    sys.path.append(ref_dir)
    import numpy as np
    from tests.helpers import random_endgames  # before the reference's imports: its checkout has a `tests` package too
    import keisei.shogi as rshogi
    from keisei.utils import PolicyOutputMapper
    from oracle import oracle as orc
    from shogidrl_b200.shogi.sfen import HostPosition
    mapper = PolicyOutputMapper()
    rng = random.Random(seed)
    plies = games = 0
    ends = {}
    endgames = _sparse_positions(64, seed) if sparse else random_endgames(64, seed)
    gi = 0
    while time.time() < deadline:
        game_seed = rng.randrange(1 << 30)
        grng = random.Random(game_seed)
        mm = grng.choice([max_moves, 40, 120])
        if sparse or gi % 2 == 1:
            k = (gi // 2) % 64
            sfen = HostPosition(endgames[0][k], endgames[1][k], int(endgames[2][k]), 0).to_sfen_string()
            g = rshogi.ShogiGame.from_sfen(sfen, mm)
            o = orc.OracleGame.from_sfen(sfen, mm)
            what = f"sfen {sfen!r}"
        else:
            g = rshogi.ShogiGame(max_moves_per_game=mm)
            o = orc.OracleGame(mm)
            what = "hirate"
        gi += 1
        tag = (seed, game_seed, what, mm)
        if g.game_over != bool(o.meta[3]):
            return ("DIVERGED", tag, "game_over at load", g.game_over, o.meta.tolist())
        ply = 0
        while not g.game_over and time.time() < deadline + 120:
            want = sorted(mapper.shogi_move_to_policy_index(m) for m in g.get_legal_moves())
            got = o.legal_indices().tolist()
            if got != want:
                return ("DIVERGED", tag, f"legal set at ply {ply}", sorted(set(want) ^ set(got)))
            if not want:
                return ("DIVERGED", tag, f"no legal move but game not over at ply {ply}")
            a = want[grng.randrange(len(want))]
            obs, reward, done, info = g.make_move(mapper.policy_index_to_shogi_move(a))
            r2, d2, reason, winner = o.make_move(a)
            ref_reason = REASONS.get(g.termination_reason, 0) if done else 0
            ref_winner = -1 if g.winner is None else g.winner.value
            if (float(reward), bool(done)) != (r2, d2) or (done and (ref_reason, ref_winner) != (reason, winner)):
                return ("DIVERGED", tag, f"outcome at ply {ply} action {a}", (reward, done, g.termination_reason, ref_winner),
                        (r2, d2, reason, winner))
            if not np.array_equal(obs, o.observation()):
                return ("DIVERGED", tag, f"observation at ply {ply} action {a}")
            if orc.parse_sfen(g.to_sfen_string())[0].tolist() != o.export()[0].tolist():
                return ("DIVERGED", tag, f"board at ply {ply} action {a}")
            ply += 1
            plies += 1
        games += 1
        ends[g.termination_reason] = ends.get(g.termination_reason, 0) + 1
    return ("OK", seed, games, plies, ends)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=5.0)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--seed", type=int, default=int.from_bytes(os.urandom(3), "little"))
    ap.add_argument("--max-moves", type=int, default=500)
    ap.add_argument("--sparse", action="store_true", help="start every game from a few-piece position without hands (repetitions)")
    a = ap.parse_args()
    deadline = time.time() + 60 * a.minutes
    with mp.get_context("spawn").Pool(a.procs) as pool:
        results = pool.map(_worker, [(a.ref, a.seed + i, deadline, a.max_moves, a.sparse) for i in range(a.procs)])
    bad = [r for r in results if r[0] != "OK"]
    for r in results:
        print(r)
    games = sum(r[2] for r in results if r[0] == "OK")
    plies = sum(r[3] for r in results if r[0] == "OK")
    print(f"seed {a.seed}: {games} games, {plies} plies compared, {len(bad)} divergence(s)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
