"""Generate the golden fixtures under tests/golden/ by importing the PYTHON REFERENCE itself.

Runs only in the build container (needs /root/reference on PYTHONPATH); the fixtures it writes
are committed so that the GPU box -- which has no reference checkout -- can check both the C
oracle and the CUDA path against real reference outputs.

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python oracle/gen_golden.py [--quick]

Fixtures
  traces_random.npz   random-legal self-play traces (BASELINE config 1 action rule: k-th legal
                      action in ascending policy-index order, k = mulhi(rand32(seed, env, step), n))
                      with auto-reset on game over, several max_moves settings
  kat_positions.npz   SFEN positions (the reference test-suite's known-answer positions plus
                      drop-heavy / pinned / uchifuzume / weird ones) -> legal sets, termination, obs
  scripted.npz        scripted traces (sennichite 4-ply cycle, max-moves) with outcomes
  gae_golden.npz      ExperienceBuffer.compute_advantages_and_returns outputs
"""
from __future__ import annotations

import argparse
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
M64 = (1 << 64) - 1
REASON_CODE = {None: 0, "Game ongoing": 0, "Tsumi": 1, "stalemate": 2, "Max moves reached": 3, "Sennichite": 4}


def rand32(seed: int, env: int, step: int) -> int:
    x = (seed ^ ((env * 0x9E3779B97F4A7C15) & M64) ^ ((step * 0xBF58476D1CE4E5B9) & M64)) & M64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & M64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & M64
    x ^= x >> 31
    return x >> 32


def obs_digest(obs: np.ndarray) -> int:
    return int.from_bytes(hashlib.blake2b(np.ascontiguousarray(obs, np.float32).tobytes(), digest_size=8).digest(), "little")


def encode_state(game):
    b = np.zeros(81, np.int8)
    for r in range(9):
        for c in range(9):
            p = game.board[r][c]
            if p is not None:
                b[r * 9 + c] = 1 + p.type.value + 14 * p.color.value
    h = np.zeros(14, np.uint8)
    for color in (0, 1):
        for pt, cnt in game.hands[color].items():
            if pt.value < 7:
                h[color * 7 + pt.value] = cnt
    return b, h


def _trace_worker(args):
    env, seed, T, max_moves = args
    from keisei.shogi import ShogiGame
    from keisei.utils import PolicyOutputMapper

    mapper = PolicyOutputMapper()
    g = ShogiGame(max_moves_per_game=max_moves)
    actions = np.zeros(T, np.int32)
    rewards = np.zeros(T, np.float32)
    dones = np.zeros(T, np.uint8)
    reasons = np.zeros(T, np.uint8)
    winners = np.full(T, -1, np.int8)
    digests = np.zeros(T, np.uint64)
    boards = np.zeros((T, 81), np.int8)
    hands = np.zeros((T, 14), np.uint8)
    sides = np.zeros(T, np.uint8)
    move_counts = np.zeros(T, np.int32)
    legal_off = np.zeros(T + 1, np.int64)
    legal = []
    full_obs = {}
    for t in range(T):
        lm = g.get_legal_moves()
        idx = sorted(mapper.shogi_move_to_policy_index(m) for m in lm)
        legal.append(np.asarray(idx, np.uint16))
        legal_off[t + 1] = legal_off[t] + len(idx)
        a = idx[(rand32(seed, env, t) * len(idx)) >> 32]
        obs, reward, done, info = g.make_move(mapper.policy_index_to_shogi_move(a))
        actions[t] = a
        rewards[t] = reward
        dones[t] = done
        reasons[t] = REASON_CODE[info["reason"]]
        winners[t] = {"BLACK": 0, "WHITE": 1}.get(info.get("winner"), -1)
        digests[t] = obs_digest(obs)
        boards[t], hands[t] = encode_state(g)
        sides[t] = g.current_player.value
        move_counts[t] = g.move_count
        if t % 97 == 0 or done:
            full_obs[t] = obs.copy()
        if done:
            g.reset()
    return dict(env=env, seed=seed, T=T, max_moves=max_moves, actions=actions, rewards=rewards, dones=dones,
                reasons=reasons, winners=winners, digests=digests, boards=boards, hands=hands, sides=sides,
                move_counts=move_counts, legal_off=legal_off, legal=np.concatenate(legal), full_obs=full_obs)


KAT_SFENS = [
    # reference test-suite known-answer positions
    "lnsgkgsnl/1r5b1/ppppppppp/9/9/9/PPPPPPPPP/1B5R1/LNSGKGSNL b - 1",
    "4k4/4r4/9/9/9/9/9/9/4K4 b - 1",
    "9/9/9/9/9/4G4/4r4/4g4/4K4 b - 1",
    "K8/9/9/9/9/9/9/9/r8 w - 1",
    "9/9/9/9/4K4/9/9/9/4k4 b P 1",
    "P8/9/9/9/4k4/9/9/9/4K4 b P 1",
    "4k4/9/9/9/9/9/9/9/4K4 b - 1",
    "4k4/9/9/9/9/R8/9/9/4K4 b - 1",
    # pins / check evasion / promotions / drops / uchifuzume / oddities (ours)
    "4k4/9/9/9/4r4/9/4G4/9/4K4 b - 1",
    "4k4/9/9/9/9/9/2b6/3S5/4K4 b - 1",
    "4k4/9/9/9/9/9/9/4r4/4K4 b G 1",
    "8k/9/8P/9/9/9/9/9/K8 b P 1",
    "7gk/9/7GP/9/9/9/9/9/K8 b P 1",
    "6R1k/9/7G1/9/9/9/9/9/K8 b P 1",
    "7nk/7g1/9/8L/9/9/9/9/K8 b P 1",
    "k8/9/1G7/9/9/9/9/9/K7R b P 1",
    "4k4/9/4P4/9/9/9/9/9/4K4 b 2P 5",
    "4k4/1R7/4P4/9/9/9/9/9/4K4 b - 9",
    "4k4/9/9/9/9/9/9/1pp6/K8 w 2p 8",
    "3+R1k3/9/4+B4/9/9/9/9/9/4K4 w 2g2s 20",
    "ln1g1g1nl/1ks2r3/1pppp1bpp/p4pp2/9/2P1P4/PPBP1PPPP/2G2S1R1/LN2KG1NL b Ss 17",
    "l6nl/5+P1gk/2np1S3/p1p4Pp/3P2Sp1/1PPb2P1P/P5GS1/R8/LN4bKL w RGgsn5p 40",
    "9/9/9/9/9/9/9/9/9 b - 1",
    "4k4/9/9/9/9/9/9/9/9 b 18P4L4N4S4G2B2R 1",
    "+P+L+N+Sk4/9/9/9/9/9/9/9/4K4 w rb 3",
    "4k4/4P4/9/9/9/9/9/9/4KL3 b N 1",
    "k1K6/9/9/9/9/9/9/9/9 b P 1",
    "kl7/9/9/9/9/9/9/9/K8 b P 1",
    "4k4/9/9/4B4/9/9/9/9/4K3r b 2R 1",
    "1k7/9/1P7/9/9/9/9/9/1L2K4 b GP 1",
    "3rk4/9/9/9/9/9/9/3L5/3K5 b - 1",
    "3rk4/9/9/9/9/9/3N5/9/3K5 b - 1",
    "2b1k4/9/9/9/9/9/6B2/9/8K w - 1",
    "4k4/9/9/9/9/9/9/3+r5/4K4 b - 1",
    "4k4/9/9/9/9/9/9/4+b4/4K4 b - 1",
    "4k4/9/9/9/9/9/4n4/9/3K5 b - 1",
    "4k4/9/9/9/9/9/3nn4/9/4K4 b G 1",
    "lnsgk1snl/6gb1/p1pppp2p/6R2/9/1rP6/P2PPPP1P/1BG6/LNS1KGSNL w 3P2p 16",
]


def _kat_worker(sfen):
    from keisei.shogi import ShogiGame
    from keisei.shogi.shogi_core_definitions import Color
    from keisei.utils import PolicyOutputMapper

    mapper = PolicyOutputMapper()
    try:
        g = ShogiGame.from_sfen(sfen)
    except Exception as e:  # the reference's SFEN parser rejects it: not a usable fixture
        print("SKIP", sfen, repr(e), flush=True)
        return None
    over, winner, reason = g.game_over, g.winner, g.termination_reason
    obs = g.get_observation()
    chk = [bool(g.is_in_check(Color.BLACK)), bool(g.is_in_check(Color.WHITE))]
    idx = sorted(mapper.shogi_move_to_policy_index(m) for m in g.get_legal_moves())
    return dict(sfen=sfen, game_over=bool(over), winner=-1 if winner is None else winner.value,
                reason=REASON_CODE[reason], legal=np.asarray(idx, np.uint16), obs=obs, in_check=chk)


def scripted():
    """Sennichite (reference tests/shogi/test_shogi_game_core_logic.py:1126-1179) and max-moves traces."""
    from keisei.shogi import ShogiGame
    from keisei.utils import PolicyOutputMapper

    mapper = PolicyOutputMapper()
    out = {}
    g = ShogiGame.from_sfen("4k4/9/9/9/9/R8/9/9/4K4 b - 1")
    cycle = [(5, 0, 5, 1, False), (0, 4, 0, 3, False), (5, 1, 5, 0, False), (0, 3, 0, 4, False)]
    acts, dones, reasons = [], [], []
    for i in range(16):
        mv = cycle[i % 4]
        obs, r, d, info = g.make_move(mv)
        acts.append(mapper.shogi_move_to_policy_index(mv))
        dones.append(d)
        reasons.append(REASON_CODE[info["reason"]])
        if d:
            break
    out["senn_sfen"] = "4k4/9/9/9/9/R8/9/9/4K4 b - 1"
    out["senn_actions"] = np.asarray(acts, np.int32)
    out["senn_dones"] = np.asarray(dones, np.uint8)
    out["senn_reasons"] = np.asarray(reasons, np.uint8)
    # from the start position: king shuffle, repetition must NOT count the initial position
    g = ShogiGame()
    cycle = [(8, 4, 7, 4, False), (0, 4, 1, 4, False), (7, 4, 8, 4, False), (1, 4, 0, 4, False)]
    acts, dones, reasons = [], [], []
    for i in range(24):
        mv = cycle[i % 4]
        obs, r, d, info = g.make_move(mv)
        acts.append(mapper.shogi_move_to_policy_index(mv))
        dones.append(d)
        reasons.append(REASON_CODE[info["reason"]])
        if d:
            break
    out["senn2_actions"] = np.asarray(acts, np.int32)
    out["senn2_dones"] = np.asarray(dones, np.uint8)
    out["senn2_reasons"] = np.asarray(reasons, np.uint8)
    return out


def gae_golden():
    import torch
    from keisei.core.experience_buffer import ExperienceBuffer

    out = {}
    rng = np.random.default_rng(7)
    cases = [(3, 0.99, 0.95), (64, 0.99, 0.95), (257, 0.997, 0.9), (2048, 0.99, 0.95)]
    for i, (T, gamma, lam) in enumerate(cases):
        buf = ExperienceBuffer(T, gamma, lam, "cpu")
        if T == 3:  # tests/conftest.py:543-581 known answer
            r = np.array([1, 2, 3], np.float32); v = np.array([.5, 1, 1.5], np.float32); d = np.array([0, 0, 1], bool)
        else:
            r = (rng.standard_normal(T) * (rng.random(T) < 0.2)).astype(np.float32)
            v = rng.standard_normal(T).astype(np.float32)
            d = rng.random(T) < 0.02
        obs = torch.zeros(46, 9, 9)
        mask = torch.zeros(13527, dtype=torch.bool)
        for t in range(T):
            buf.add(obs, 0, float(r[t]), 0.0, float(v[t]), bool(d[t]), mask)
        last = float(np.float32(rng.standard_normal()))
        buf.compute_advantages_and_returns(last)
        out[f"c{i}_r"] = r; out[f"c{i}_v"] = v; out[f"c{i}_d"] = d.astype(np.uint8)
        out[f"c{i}_last"] = np.float32(last); out[f"c{i}_gamma"] = gamma; out[f"c{i}_lam"] = lam
        out[f"c{i}_adv"] = buf.advantages.numpy().copy(); out[f"c{i}_ret"] = buf.returns.numpy().copy()
    out["n_cases"] = len(cases)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    t0 = time.time()

    np.savez_compressed(os.path.join(GOLD, "gae_golden.npz"), **gae_golden())
    np.savez_compressed(os.path.join(GOLD, "scripted.npz"), **scripted())
    print("gae + scripted done", time.time() - t0, flush=True)

    with mp.Pool(args.procs) as pool:
        kats = [x for x in pool.map(_kat_worker, KAT_SFENS) if x is not None]
        k = {"sfens": np.asarray([x["sfen"] for x in kats]),
             "game_over": np.asarray([x["game_over"] for x in kats], np.uint8),
             "winner": np.asarray([x["winner"] for x in kats], np.int8),
             "reason": np.asarray([x["reason"] for x in kats], np.uint8),
             "in_check": np.asarray([x["in_check"] for x in kats], np.uint8),
             "obs": np.stack([x["obs"] for x in kats]),
             "legal": np.concatenate([x["legal"] for x in kats]),
             "legal_off": np.cumsum([0] + [len(x["legal"]) for x in kats]).astype(np.int64)}
        np.savez_compressed(os.path.join(GOLD, "kat_positions.npz"), **k)
        print("kat done", time.time() - t0, flush=True)

        seed = 1234
        if args.quick:
            jobs = [(e, seed, 60, 500) for e in range(4)] + [(100 + e, seed, 60, 30) for e in range(4)]
        else:
            jobs = ([(e, seed, 520, 500) for e in range(40)] + [(100 + e, seed, 300, 60) for e in range(12)]
                    + [(200 + e, seed, 200, 14) for e in range(4)])
        res = pool.map(_trace_worker, jobs, chunksize=1)
    T_tot = sum(r["T"] for r in res)
    cat = lambda key: np.concatenate([r[key] for r in res])
    legal_off = [0]
    for r in res:
        base = legal_off[-1]
        legal_off.extend((r["legal_off"][1:] + base).tolist())
    fo_idx, fo = [], []
    base = 0
    for r in res:
        for t, o in sorted(r["full_obs"].items()):
            fo_idx.append(base + t)
            fo.append(o)
        base += r["T"]
    out = dict(env=np.asarray([r["env"] for r in res], np.int32), seed=np.int64(seed),
               T=np.asarray([r["T"] for r in res], np.int32), max_moves=np.asarray([r["max_moves"] for r in res], np.int32),
               actions=cat("actions"), rewards=cat("rewards"), dones=cat("dones"), reasons=cat("reasons"),
               winners=cat("winners"), digests=cat("digests"), boards=cat("boards"), hands=cat("hands"),
               sides=cat("sides"), move_counts=cat("move_counts"), legal=cat("legal"),
               legal_off=np.asarray(legal_off, np.int64), full_obs_idx=np.asarray(fo_idx, np.int64),
               full_obs=np.stack(fo).astype(np.float32))
    np.savez_compressed(os.path.join(GOLD, "traces_random.npz"), **out)
    print(f"traces done: {T_tot} plies, {int(out['dones'].sum())} finished games, {time.time() - t0:.0f}s", flush=True)


if __name__ == "__main__":
    sys.exit(main())
