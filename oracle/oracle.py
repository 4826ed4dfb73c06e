"""ctypes front-end of the CPU ORACLE (``oracle/keisei_oracle.c``).

TEST INFRASTRUCTURE ONLY.  Importable from ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs; never from ``shogidrl_b200``.
The library is the checker, not the product: the product path (``shogidrl_b200``) has no
CPU fallback and fails loudly when its CUDA extension is missing.

Parity status: pinned against the Python reference by ``tests/test_oracle_golden.py``
(fixtures from ``oracle/gen_golden.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkeisei_oracle.so")

NSQ = 81
NACT = 13527
OBS_FLOATS = 46 * 81
REASONS = {0: None, 1: "Tsumi", 2: "stalemate", 3: "Max moves reached", 4: "Sennichite"}


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, no reference build system)."""
    src = os.path.join(_HERE, "keisei_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "CC=gcc"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
        L.orc_new.restype = vp
        L.orc_free.argtypes = [vp]
        L.orc_reset.argtypes = [vp, i32]
        L.orc_load.argtypes = [vp, vp, vp, i32, i32, i32, i32]
        L.orc_export.argtypes = [vp, vp, vp, vp]
        L.orc_legal_moves.argtypes = [vp, vp]
        L.orc_legal_moves.restype = i32
        L.orc_legal_mask.argtypes = [vp, vp]
        L.orc_legal_mask.restype = i32
        L.orc_in_check.argtypes = [vp, i32]
        L.orc_in_check.restype = i32
        L.orc_piece_targets.argtypes = [vp, i32, vp]
        L.orc_piece_targets.restype = i32
        L.orc_uchi_fu_zume.argtypes = [vp, i32, i32]
        L.orc_uchi_fu_zume.restype = i32
        L.orc_can_drop.argtypes = [vp, i32, i32, i32]
        L.orc_can_drop.restype = i32
        L.orc_king_legal_moves.argtypes = [vp, i32]
        L.orc_king_legal_moves.restype = i32
        L.orc_observation.argtypes = [vp, vp]
        L.orc_make_move.argtypes = [vp, i32, vp]
        L.orc_make_move.restype = i32
        L.orc_rand32.argtypes = [u64, u64, u64]
        L.orc_rand32.restype = C.c_uint32
        L.orc_pick_action.argtypes = [vp, u64, u64, u64]
        L.orc_pick_action.restype = i32
        L.orc_selfplay.argtypes = [i32, i32, i32, i32, i32, u64] + [vp] * 10 + [i32] + [vp] * 4
        L.orc_selfplay.restype = C.c_long
        L.orc_batch_new.argtypes = [i32, i32]
        L.orc_batch_new.restype = vp
        L.orc_batch_free.argtypes = [vp]
        L.orc_batch_run.argtypes = [vp, i32, i32, i32, u64, i32, C.POINTER(C.c_double)]
        L.orc_batch_run.restype = C.c_long
        L.orc_gae.argtypes = [vp, vp, vp, vp, i32, i32, C.c_double, C.c_double, vp, vp]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ SFEN (host-side helper)
_SFEN_TYPES = {"P": 0, "L": 1, "N": 2, "S": 3, "G": 4, "B": 5, "R": 6, "K": 7}


def parse_sfen(sfen: str) -> Tuple[np.ndarray, np.ndarray, int, int]:
    """Minimal SFEN parser (board codes, hands14, side, move_count) following the reference's
    conventions: shogi_game_io.py:169-306 and shogi_game.py:293-310 (move_count = n - 1)."""
    board_s, turn, hands_s, num = sfen.strip().split()
    board = np.zeros(81, dtype=np.int8)
    rows = board_s.split("/")
    assert len(rows) == 9, sfen
    for r, row in enumerate(rows):
        c = 0
        promo = False
        for ch in row:
            if ch == "+":
                promo = True
            elif ch.isdigit():
                c += int(ch)
            else:
                t = _SFEN_TYPES[ch.upper()]
                color = 0 if ch.isupper() else 1
                if promo:
                    t = {0: 8, 1: 9, 2: 10, 3: 11, 5: 12, 6: 13}[t]
                    promo = False
                board[r * 9 + c] = 1 + t + 14 * color
                c += 1
        assert c == 9, sfen
    hands = np.zeros(14, dtype=np.uint8)
    if hands_s != "-":
        n = ""
        for ch in hands_s:
            if ch.isdigit():
                n += ch
            else:
                cnt = int(n) if n else 1
                n = ""
                hands[(0 if ch.isupper() else 7) + _SFEN_TYPES[ch.upper()]] += cnt
    return board, hands, 0 if turn == "b" else 1, int(num) - 1


class OracleGame:
    """One reference-semantics game on the CPU (mirrors ``keisei.shogi.ShogiGame`` behaviour)."""

    def __init__(self, max_moves: int = 500):
        self._L = lib()
        self._g = C.c_void_p(self._L.orc_new())
        self.max_moves = max_moves
        self._L.orc_reset(self._g, max_moves)

    def __del__(self):
        try:
            self._L.orc_free(self._g)
        except Exception:
            pass

    def reset(self):
        self._L.orc_reset(self._g, self.max_moves)

    @classmethod
    def from_sfen(cls, sfen: str, max_moves: int = 500, evaluate_termination: bool = True) -> "OracleGame":
        b, h, side, mc = parse_sfen(sfen)
        return cls.from_arrays(b, h, side, mc, max_moves, evaluate_termination)

    @classmethod
    def from_arrays(cls, board, hands, side, move_count, max_moves=500, evaluate_termination=True):
        g = cls(max_moves)
        b = np.ascontiguousarray(board, dtype=np.int8)
        h = np.ascontiguousarray(hands, dtype=np.uint8)
        g._L.orc_load(g._g, _p(b), _p(h), int(side), int(move_count), int(max_moves), int(evaluate_termination))
        return g

    def export(self):
        b = np.zeros(81, np.int8)
        h = np.zeros(14, np.uint8)
        m = np.zeros(6, np.int32)
        self._L.orc_export(self._g, _p(b), _p(h), _p(m))
        return b, h, m  # meta = side, move_count, max_moves, game_over, winner, reason

    @property
    def meta(self):
        return self.export()[2]

    def legal_indices(self) -> np.ndarray:
        out = np.zeros(1024, np.uint16)
        n = self._L.orc_legal_moves(self._g, _p(out))
        return np.sort(out[:n].astype(np.int32))

    def legal_mask(self) -> np.ndarray:
        m = np.zeros(NACT, np.uint8)
        self._L.orc_legal_mask(self._g, _p(m))
        return m

    def in_check(self, color: int) -> bool:
        return bool(self._L.orc_in_check(self._g, color))

    def piece_targets(self, sq: int) -> np.ndarray:
        """generate_piece_potential_moves (shogi_rules_logic.py:82-208) of the piece on ``sq`` as uint8[81]."""
        out = np.zeros(81, np.uint8)
        self._L.orc_piece_targets(self._g, int(sq), _p(out))
        return out

    def is_uchi_fu_zume(self, sq: int, color: int) -> bool:
        return bool(self._L.orc_uchi_fu_zume(self._g, int(sq), int(color)))

    def can_drop(self, piece_type: int, sq: int, color: int) -> bool:
        return bool(self._L.orc_can_drop(self._g, int(piece_type), int(sq), int(color)))

    def king_legal_moves(self, color: int) -> int:
        return int(self._L.orc_king_legal_moves(self._g, int(color)))

    def observation(self) -> np.ndarray:
        o = np.zeros((46, 9, 9), np.float32)
        self._L.orc_observation(self._g, _p(o))
        return o

    def make_move(self, action: int):
        o4 = np.zeros(4, np.float32)
        rc = self._L.orc_make_move(self._g, int(action), _p(o4))
        if rc != 0:
            raise ValueError({-1: "Invalid move index", -2: "Invalid move: no piece / wrong colour",
                              -3: "Illegal movement pattern"}[rc])
        return float(o4[0]), bool(o4[1]), int(o4[2]), int(o4[3])

    def pick_action(self, seed: int, env: int, step: int) -> int:
        return self._L.orc_pick_action(self._g, seed, env, step)


def rand32(seed: int, env: int, step: int) -> int:
    return int(lib().orc_rand32(seed, env, step))


def selfplay(n_envs: int, T: int, *, env0: int = 0, step0: int = 0, max_moves: int = 500, seed: int = 1234,
             threads: int = 1, want_traces: bool = True, want_final: bool = True, start=None):
    """Random-legal self-play with auto-reset (CPU mirror of BASELINE config 2).  ``start`` = optional
    (boards [n,81] int8, hands [n,14] uint8, sides [n] uint8, move_counts [n] int32) initial positions."""
    L = lib()
    out = {}
    if want_traces:
        out["actions"] = np.zeros((T, n_envs), np.int32)
        out["rewards"] = np.zeros((T, n_envs), np.float32)
        out["dones"] = np.zeros((T, n_envs), np.uint8)
        out["reasons"] = np.zeros((T, n_envs), np.uint8)
        out["legal_counts"] = np.zeros((T, n_envs), np.int32)
    if want_final:
        out["obs"] = np.zeros((n_envs, 46, 9, 9), np.float32)
        out["mask"] = np.zeros((n_envs, NACT), np.uint8)
        out["boards"] = np.zeros((n_envs, 81), np.int8)
        out["hands"] = np.zeros((n_envs, 14), np.uint8)
        out["meta"] = np.zeros((n_envs, 6), np.int32)
    g = out.get
    total = L.orc_selfplay(n_envs, env0, T, step0, max_moves, seed, _p(g("actions")), _p(g("rewards")),
                           _p(g("dones")), _p(g("reasons")), _p(g("legal_counts")), _p(g("obs")),
                           _p(g("mask")), _p(g("boards")), _p(g("hands")), _p(g("meta")), threads,
                           *([None] * 4 if start is None else [
                               _p(np.ascontiguousarray(start[0], np.int8)), _p(np.ascontiguousarray(start[1], np.uint8)),
                               _p(np.ascontiguousarray(start[2], np.uint8)), _p(np.ascontiguousarray(start[3], np.int32))]))
    out["total_steps"] = int(total)
    return out


class OracleBatch:
    """Persistent batch of oracle games advanced stepwise (CPU baseline / reference arm timing)."""

    def __init__(self, n: int, max_moves: int = 500, seed: int = 1234, threads: int = 1, env0: int = 0):
        self._L = lib()
        self._b = C.c_void_p(self._L.orc_batch_new(n, max_moves))
        self.n, self.seed, self.threads, self.env0, self.step = n, seed, threads, env0, 0
        self.mean_ply = 0.0

    def run(self, T: int) -> int:
        mp = C.c_double()
        done = self._L.orc_batch_run(self._b, T, self.step, self.env0, self.seed, self.threads, C.byref(mp))
        self.step += T
        self.mean_ply = mp.value
        return int(done)

    def __del__(self):
        try:
            self._L.orc_batch_free(self._b)
        except Exception:
            pass


def gae(rewards: np.ndarray, values: np.ndarray, dones: np.ndarray, last_value: np.ndarray,
        gamma: float, lam: float):
    """[T, N] column-wise GAE with the reference's fp32 op order (experience_buffer.py:99-145)."""
    r = np.ascontiguousarray(rewards, np.float32)
    v = np.ascontiguousarray(values, np.float32)
    d = np.ascontiguousarray(dones, np.uint8)
    T, N = r.shape
    lv = np.ascontiguousarray(np.broadcast_to(np.asarray(last_value, np.float32), (N,)))
    adv = np.zeros_like(r)
    ret = np.zeros_like(r)
    lib().orc_gae(_p(r), _p(v), _p(d), _p(lv), T, N, float(gamma), float(lam), _p(adv), _p(ret))
    return adv, ret


def gae_numpy(rewards, values, dones, last_value, gamma: float, lam: float):
    """Same computation in numpy fp32 scalars (second, independent restatement)."""
    r = np.asarray(rewards, np.float32)
    v = np.asarray(values, np.float32)
    T, N = r.shape
    m = (np.float32(1.0) - np.asarray(dones).astype(np.float32)).astype(np.float32)
    lv = np.broadcast_to(np.asarray(last_value, np.float32), (N,))
    g32 = np.float32(gamma)
    gl32 = np.float32(gamma * lam)
    adv = np.zeros_like(r)
    ret = np.zeros_like(r)
    gae_v = np.zeros(N, np.float32)
    for t in range(T - 1, -1, -1):
        nv = lv if t == T - 1 else v[t + 1]
        delta = ((r[t] + ((g32 * nv).astype(np.float32) * m[t]).astype(np.float32)).astype(np.float32) - v[t]).astype(np.float32)
        gae_v = (delta + ((gl32 * m[t]).astype(np.float32) * gae_v).astype(np.float32)).astype(np.float32)
        adv[t] = gae_v
        ret[t] = (gae_v + v[t]).astype(np.float32)
    return adv, ret


def index_to_move(idx: int):
    """Closed form of PolicyOutputMapper.idx_to_move (keisei/utils/utils.py:208-266)."""
    if idx < 12960:
        promo = bool(idx & 1)
        pair = idx >> 1
        f, t = divmod(pair, 80)
        to = t + (1 if t >= f else 0)
        return (f // 9, f % 9, to // 9, to % 9, promo)
    k = idx - 12960
    to, pt = divmod(k, 7)
    return (None, None, to // 9, to % 9, pt)
