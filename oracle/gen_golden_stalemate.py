"""Golden fixture for STALEMATE (reason "stalemate", a draw; keisei/shogi/shogi_game.py:431-435), produced by importing
the Python reference itself:

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python oracle/gen_golden_stalemate.py

Stalemate never occurs in the random-play and drop-heavy fixtures (their finished games are Tsumi / max-moves), so this
generator looks for it: (1) the two positions of the reference's own tests -- stalemate at load
(tests/shogi/test_shogi_engine_integration.py:793) and stalemate by a move (test_shogi_game_core_logic.py:957);
(2) random "bare king" endgames -- one side has only its king and an empty hand, the other a few step pieces -- played
with the config-1 action rule (k-th legal action in ascending policy-index order).  The C oracle port searches candidate
positions quickly for games that END in stalemate; the ones it finds are then replayed ply by ply through the reference,
and only what the reference computes is written.  Test infrastructure only; writes tests/golden/traces_stalemate.npz in
the layout of traces_endgame.npz (+ kat_* arrays for the two known-answer positions)."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from gen_golden import REASON_CODE, encode_state, obs_digest, rand32  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED, PLIES, WANT = 777, 40, 12
KAT_LOAD = "k8/2G6/1G7/9/9/9/9/9/8K w - 1"          # White to move, not in check, no legal move
KAT_MOVE = ("k8/9/1GG6/9/9/9/9/9/8K b - 1", (2, 2, 1, 2, False))  # Black's G c3-c2 stalemates White
SYM = "PLNSGBRK"


def sfen_of(board, hands, side, move_count):
    rows = []
    for r in range(9):
        row, run = "", 0
        for c in range(9):
            code = int(board[r * 9 + c])
            if code == 0:
                run += 1
                continue
            if run:
                row += str(run); run = 0
            t, col = (code - 1) % 14, (code - 1) // 14
            ch = ("+" + SYM[{8: 0, 9: 1, 10: 2, 11: 3, 12: 5, 13: 6}[t]]) if t >= 8 else SYM[t]
            row += ch.lower() if col else ch
        if run:
            row += str(run)
        rows.append(row)
    hs = ""
    for col in (0, 1):
        for t in (6, 5, 4, 3, 2, 1, 0):
            n = int(hands[col * 7 + t])
            if n:
                ch = SYM[t].lower() if col else SYM[t]
                hs += (str(n) if n > 1 else "") + ch
    return "/".join(rows) + (" b " if side == 0 else " w ") + (hs or "-") + f" {move_count + 1}"


def candidates(rng, count):
    """Bare-king endgames: the strong side has its king and 2-4 gold / silver / promoted-pawn movers near the bare king."""
    out = []
    while len(out) < count:
        bare = int(rng.integers(0, 2))
        kb = int(rng.choice([0, 8, 72, 80, 1, 7, 9, 17, 63, 71, 73, 79, 4, 36, 44, 76]))
        board = np.zeros(81, np.int8)
        board[kb] = 1 + 7 + 14 * bare
        kr, kc = divmod(kb, 9)
        ks = int(rng.integers(0, 81))
        if max(abs(ks // 9 - kr), abs(ks % 9 - kc)) < 3:
            continue
        board[ks] = 1 + 7 + 14 * (1 - bare)
        ok = True
        for _ in range(int(rng.integers(2, 5))):
            r, c = kr + int(rng.integers(-3, 4)), kc + int(rng.integers(-3, 4))
            if not (0 <= r < 9 and 0 <= c < 9) or board[r * 9 + c] != 0:
                ok = False
                break
            board[r * 9 + c] = 1 + int(rng.choice([4, 4, 3, 8, 11])) + 14 * (1 - bare)
        if not ok:
            continue
        side = int(rng.integers(0, 2))
        out.append((board, np.zeros(14, np.uint8), side, int(rng.integers(0, 20))))
    return out


def search(n_candidates=60000):
    """Play every candidate with the C oracle; keep the ones whose game ends in stalemate within PLIES plies."""
    from oracle import oracle as orc
    rng = np.random.default_rng(SEED)
    found = []
    for env, (board, hands, side, mc) in enumerate(candidates(rng, n_candidates)):
        g = orc.OracleGame.from_arrays(board, hands, side, mc, 500, True)
        if g.meta[3]:
            continue
        if g.in_check(1 - side):
            continue
        for t in range(PLIES):
            a = g.pick_action(SEED, env, t)
            if a < 0:
                break
            _, done, reason, _ = g.make_move(a)
            if done:
                if reason == 2:
                    found.append((env, sfen_of(board, hands, side, mc)))
                break
        if len(found) >= WANT:
            break
    return found


def replay_on_reference(env, sfen):
    from keisei.shogi import ShogiGame
    from keisei.utils import PolicyOutputMapper
    mapper = PolicyOutputMapper()
    g = ShogiGame.from_sfen(sfen)
    assert not g.game_over, sfen
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


def known_answers():
    from keisei.shogi import ShogiGame
    from keisei.utils import PolicyOutputMapper
    mapper = PolicyOutputMapper()
    out = {}
    g = ShogiGame.from_sfen(KAT_LOAD)
    out["kat_load_sfen"] = KAT_LOAD
    out["kat_load_game_over"] = np.uint8(g.game_over)
    out["kat_load_reason"] = np.uint8(REASON_CODE[g.termination_reason])
    out["kat_load_winner"] = np.int8(-1 if g.winner is None else g.winner.value)
    out["kat_load_legal_n"] = np.int32(len(g.get_legal_moves()))
    out["kat_load_obs"] = g.get_observation()
    sfen, mv = KAT_MOVE
    g = ShogiGame.from_sfen(sfen)
    assert not g.game_over
    idx = sorted(mapper.shogi_move_to_policy_index(m) for m in g.get_legal_moves())
    a = mapper.shogi_move_to_policy_index(mv)
    assert a in idx
    obs, reward, done, info = g.make_move(mv)
    out["kat_move_sfen"] = sfen
    out["kat_move_legal"] = np.asarray(idx, np.uint16)
    out["kat_move_action"] = np.int32(a)
    out["kat_move_reward"] = np.float32(reward)
    out["kat_move_done"] = np.uint8(done)
    out["kat_move_reason"] = np.uint8(REASON_CODE[info["reason"]])
    out["kat_move_winner"] = np.int8({"BLACK": 0, "WHITE": 1}.get(info.get("winner"), -1))
    out["kat_move_obs"] = obs
    out["kat_move_legal_after"] = np.int32(len(g.get_legal_moves()))
    return out


def main():
    found = search()
    print(f"oracle search: {len(found)} candidate games end in stalemate", flush=True)
    games = [replay_on_reference(env, sfen) for env, sfen in found]
    games = [g for g in games if g[4]["reasons"][-1] == 2]  # only what the REFERENCE calls a stalemate
    assert games, "no stalemate game found"
    cat = lambda k, dt: np.concatenate([np.asarray(g[4][k], dt) for g in games])
    T = np.asarray([len(g[4]["actions"]) for g in games], np.int32)
    legal_n = cat("legal_n", np.int64)
    out = dict(seed=np.int64(SEED), envs=np.asarray([g[0] for g in games], np.int64),
               sfens=np.asarray([g[1] for g in games]), T=T,
               start_boards=np.stack([g[2] for g in games]), start_hands=np.stack([g[3] for g in games]),
               actions=cat("actions", np.int32), rewards=cat("rewards", np.float32), dones=cat("dones", np.uint8),
               reasons=cat("reasons", np.uint8), winners=cat("winners", np.int8), digests=cat("digests", np.uint64),
               boards=np.concatenate([np.stack(g[4]["boards"]) for g in games]),
               hands=np.concatenate([np.stack(g[4]["hands"]) for g in games]),
               sides=cat("sides", np.uint8), move_counts=cat("move_counts", np.int32),
               legal_off=np.concatenate([[0], np.cumsum(legal_n)]).astype(np.int64),
               legal=np.concatenate([np.concatenate(g[4]["legal"]) for g in games]))
    out.update(known_answers())
    np.savez_compressed(os.path.join(GOLD, "traces_stalemate.npz"), **out)
    print(f"{len(games)} reference games ending in stalemate, {int(T.sum())} plies; KAT load reason "
          f"{int(out['kat_load_reason'])}, KAT move reason {int(out['kat_move_reason'])} reward {float(out['kat_move_reward'])}")


if __name__ == "__main__":
    main()
