"""Golden traces from DROP-HEAVY ENDGAMES, produced by importing the Python reference itself:

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python oracle/gen_golden_endgame.py

Random positions with two kings, a few pieces and full hands (BASELINE config 5's "drops-heavy endgames") are loaded
through the reference's ShogiGame.from_sfen and played with the config-1 action rule (k-th legal action in ascending
policy-index order, k = mulhi(rand32(seed, env, ply), n)) until the game ends or PLIES plies have been made.  This is
where nifu, last-rank drop limits, drops that answer a check and uchifuzume decide the legal set.  Test
infrastructure only; writes tests/golden/traces_endgame.npz."""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import REASON_CODE, encode_state, obs_digest, rand32  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED, N_POS, PLIES = 4242, 128, 48


def random_position(rng):
    """-> SFEN of a random endgame that the reference accepts as ongoing with the opponent not in check, or None."""
    from keisei.shogi import ShogiGame
    from keisei.shogi.shogi_core_definitions import Color, Piece, PieceType

    types = [0, 0, 0, 1, 2, 3, 4, 4, 5, 6, 8, 9, 10, 11, 12, 13]
    g = ShogiGame()
    g.board = [[None] * 9 for _ in range(9)]
    k0, k1 = (int(x) for x in rng.choice(81, 2, replace=False))
    if max(abs(k0 // 9 - k1 // 9), abs(k0 % 9 - k1 % 9)) < 2:
        return None
    g.board[k0 // 9][k0 % 9] = Piece(PieceType.KING, Color.BLACK)
    g.board[k1 // 9][k1 % 9] = Piece(PieceType.KING, Color.WHITE)
    for color in (Color.BLACK, Color.WHITE):
        for _ in range(int(rng.integers(0, 6))):
            sq, t = int(rng.integers(0, 81)), int(rng.choice(types))
            r, c = divmod(sq, 9)
            if g.board[r][c] is not None:
                continue
            last, second = (0, 1) if color == Color.BLACK else (8, 7)
            if (t in (0, 1) and r == last) or (t == 2 and r in (last, second)):
                continue
            if t == 0 and any(g.board[rr][c] is not None and g.board[rr][c].type == PieceType.PAWN
                              and g.board[rr][c].color == color for rr in range(9)):
                continue
            g.board[r][c] = Piece(PieceType(t), color)
    for color in (0, 1):
        for pt in g.hands[color]:
            g.hands[color][pt] = 0
        g.hands[color][PieceType.PAWN] = int(rng.integers(0, 5))
        for t in range(1, 7):
            g.hands[color][PieceType(t)] = int(rng.integers(0, 3)) if rng.random() < 0.5 else 0
    g.current_player = Color.BLACK if int(rng.integers(0, 2)) == 0 else Color.WHITE
    g.move_count = 0
    if g.is_in_check(g.current_player.opponent()):
        return None
    sfen = g.to_sfen_string()
    loaded = ShogiGame.from_sfen(sfen)
    if loaded.game_over or not loaded.get_legal_moves():
        return None
    return sfen


def play(args):
    env, sfen = args
    from keisei.shogi import ShogiGame
    from keisei.utils import PolicyOutputMapper

    mapper = PolicyOutputMapper()
    g = ShogiGame.from_sfen(sfen)
    rec = dict(actions=[], rewards=[], dones=[], reasons=[], winners=[], digests=[], boards=[], hands=[], sides=[],
               move_counts=[], legal=[], legal_n=[])
    b0, h0 = encode_state(g)
    for t in range(PLIES):
        idx = sorted(mapper.shogi_move_to_policy_index(m) for m in g.get_legal_moves())
        a = idx[(rand32(SEED, env, t) * len(idx)) >> 32]
        obs, reward, done, info = g.make_move(mapper.policy_index_to_shogi_move(a))
        b, h = encode_state(g)
        rec["legal"].append(np.asarray(idx, np.uint16)); rec["legal_n"].append(len(idx))
        rec["actions"].append(a); rec["rewards"].append(reward); rec["dones"].append(done)
        rec["reasons"].append(REASON_CODE[info["reason"]])
        rec["winners"].append({"BLACK": 0, "WHITE": 1}.get(info.get("winner"), -1))
        rec["digests"].append(obs_digest(obs)); rec["boards"].append(b); rec["hands"].append(h)
        rec["sides"].append(g.current_player.value); rec["move_counts"].append(g.move_count)
        if done:
            break
    return env, sfen, b0, h0, rec


def main():
    rng = np.random.default_rng(SEED)
    sfens = []
    while len(sfens) < N_POS:
        s = random_position(rng)
        if s is not None:
            sfens.append(s)
    with mp.get_context("spawn").Pool(min(8, os.cpu_count() or 1)) as pool:
        games = pool.map(play, list(enumerate(sfens)), chunksize=1)
    cat = lambda k, dt: np.concatenate([np.asarray(g[4][k], dt) for g in games])
    T = np.asarray([len(g[4]["actions"]) for g in games], np.int32)
    legal_n = cat("legal_n", np.int64)
    out = dict(seed=np.int64(SEED), sfens=np.asarray([g[1] for g in games]), T=T,
               start_boards=np.stack([g[2] for g in games]), start_hands=np.stack([g[3] for g in games]),
               actions=cat("actions", np.int32), rewards=cat("rewards", np.float32), dones=cat("dones", np.uint8),
               reasons=cat("reasons", np.uint8), winners=cat("winners", np.int8), digests=cat("digests", np.uint64),
               boards=np.concatenate([np.stack(g[4]["boards"]) for g in games]),
               hands=np.concatenate([np.stack(g[4]["hands"]) for g in games]),
               sides=cat("sides", np.uint8), move_counts=cat("move_counts", np.int32),
               legal_off=np.concatenate([[0], np.cumsum(legal_n)]).astype(np.int64),
               legal=np.concatenate([np.concatenate(g[4]["legal"]) for g in games]))
    np.savez_compressed(os.path.join(GOLD, "traces_endgame.npz"), **out)
    drops = int((out["actions"] >= 12960).sum())
    print(f"{len(games)} games, {int(T.sum())} plies, {drops} drops, mean legal {legal_n.mean():.1f}, max {legal_n.max()}, "
          f"finished {int(out['dones'].sum())} (reasons {np.bincount(out['reasons'][out['dones'] != 0], minlength=5).tolist()})")


if __name__ == "__main__":
    main()
