"""VecShogiEnv -- N device-resident Shogi games stepped by one kernel launch.

Batched counterpart of ``keisei.shogi.ShogiGame`` + the reset-on-done handling of
``keisei.training.step_manager.StepManager`` (step_manager.py:98-348, 437-440).  All arrays live on the
GPU; observations and masks are written by the kernel straight into caller-provided rollout storage."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _native as nv


class VecShogiEnv:
    def __init__(self, num_envs: int, max_moves_per_game: int = 500, device="cuda", seed: int = 1234,
                 env_offset: int = 0, auto_reset: bool = True, hist_cap: Optional[int] = None,
                 step_streams: Optional[int] = None):
        self.device = nv.require_cuda(device)
        self.n = int(num_envs)
        self.max_moves = int(max_moves_per_game)
        self.hist_cap = int(hist_cap if hist_cap is not None else max_moves_per_game)
        self.seed = int(seed)
        self.env_offset = int(env_offset)
        self.auto_reset = bool(auto_reset)
        self._L = nv.lib()
        nv.init_tables(self.device)
        offs = (C.c_int64 * 3)()
        total = C.c_int64()
        nv.check(self._L.kz_state_layout(self.n, self.hist_cap, offs, C.byref(total)), "kz_state_layout")
        self.state = torch.zeros(total.value, dtype=torch.uint8, device=self.device)
        self.state_bytes = total.value
        # persistent per-step outputs
        d = self.device
        self.obs = torch.zeros((self.n, 46, 9, 9), dtype=torch.float32, device=d)
        self._mask_store = torch.zeros((self.n, nv.MASK_PAD_STRIDE), dtype=torch.uint8, device=d)
        self.mask = self._mask_store[:, : nv.NUM_ACTIONS]  # uint8 view, row stride 13536
        # per-step scalar results live in ONE contiguous buffer (15 bytes per env: next_actions i64, reward f32,
        # done u8, reason u8, winner i8) so that a host consumer fetches them with a single device-to-host copy
        n = self.n
        self.results = torch.zeros(15 * n, dtype=torch.uint8, device=d)
        self.next_actions = self.results[: 8 * n].view(torch.int64)
        self.reward = self.results[8 * n: 12 * n].view(torch.float32)
        self.done = self.results[12 * n: 13 * n]
        self.reason = self.results[13 * n: 14 * n]
        self.winner = self.results[14 * n:].view(torch.int8)
        self.ep_len = torch.zeros(self.n, dtype=torch.int32, device=d)
        self.legal_count = torch.zeros(self.n, dtype=torch.int32, device=d)
        self.step_index = 0
        # A step can be launched as `step_streams` ranges of games on as many streams (kz_step_range; results identical to
        # one launch).  Off by default: joined back into the caller's stream after every step, the ranges only share the
        # SMs within one step and the step gets no shorter (measured on B200, 65,536 games: 0.344 ms as one launch, 0.356
        # as two ranges, 0.364 as four).  What does pay is letting ranges run AHEAD of each other across steps, which
        # needs a caller that consumes each range's results separately: HostPipelinedEnv.
        if step_streams is None:
            step_streams = 1
        self.step_streams = max(1, int(step_streams))
        if self.n % (8 * self.step_streams) != 0:
            self.step_streams = 1
        self._side_streams = [torch.cuda.Stream(device=d) for _ in range(self.step_streams)] if self.step_streams > 1 else []
        self.reset()

    # ------------------------------------------------------------------ helpers
    def _sp(self):
        return nv.stream_ptr(self.device)

    def _rows_arg(self, t: Optional[torch.Tensor], what: str, dtypes, row_elems: int):
        """(data pointer, row stride in elements) of caller storage the kernel writes ``n`` rows into; rejects
        anything the kernel would write out of bounds (the C ABI sees raw pointers only)."""
        if t is None:
            return None, 0
        if t.device != self.device or t.dtype not in dtypes:
            raise ValueError(f"{what}: expected a {dtypes[0]} tensor on {self.device}, got {t.dtype} on {t.device}")
        if t.dim() > 1:
            rows, stride = t.shape[0], t.stride(0)
            inner = t[0]
        else:
            rows, stride, inner = 1, row_elems, t
        if rows < self.n or inner.numel() < row_elems or not inner.is_contiguous():
            raise ValueError(f"{what}: need at least {self.n} rows of {row_elems} contiguous elements, got shape "
                             f"{tuple(t.shape)} with strides {tuple(t.stride())}")
        return t.data_ptr(), stride

    def _obs_args(self, obs: Optional[torch.Tensor]):
        return self._rows_arg(obs, "obs", (torch.float32,), nv.OBS_FLOATS)

    def _mask_args(self, mask: Optional[torch.Tensor]):
        return self._rows_arg(mask, "mask", (torch.uint8, torch.bool), nv.NUM_ACTIONS)

    # ------------------------------------------------------------------ API
    def reset(self, env_mask: Optional[torch.Tensor] = None, refresh: bool = True, random_actions: bool = False):
        """ShogiGame.reset for all (or the masked) envs; returns (obs, mask) of the new positions."""
        mp = None
        if env_mask is not None:
            env_mask = env_mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mp = env_mask.data_ptr()
        nv.check(self._L.kz_reset(self.state.data_ptr(), self.n, self.hist_cap, mp, self.max_moves, self._sp()), "kz_reset")
        if env_mask is None:
            self.step_index = 0
        if refresh:
            self.refresh(random_actions=random_actions)
        return self.obs, self.mask

    def refresh(self, obs: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                eval_termination: bool = False, random_actions: bool = False,
                next_out: Optional[torch.Tensor] = None, in_check: Optional[torch.Tensor] = None):
        """Recompute obs / mask / legal_count (and optionally uniform-random legal actions) in place."""
        obs = self.obs if obs is None else obs
        mask = self.mask if mask is None else mask
        op, os_ = self._obs_args(obs)
        mp, ms = self._mask_args(mask)
        nv.check(self._L.kz_refresh(self.state.data_ptr(), self.n, self.hist_cap, op, os_, mp, ms,
                                    self.legal_count.data_ptr(),
                                    (self.next_actions if next_out is None else next_out).data_ptr() if random_actions else None,
                                    1, self.seed, self.step_index, self.env_offset, int(eval_termination),
                                    nv.ptr(in_check), self._sp()),
                 "kz_refresh")
        return obs, mask

    def step(self, actions: torch.Tensor, obs: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
             random_actions: bool = False, write_obs: bool = True, write_mask: bool = True,
             next_out: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """make_move for every env.  ``obs`` / ``mask`` may point into a rollout buffer (e.g. obs_buf[t+1]).
        With ``random_actions`` the kernel also writes a uniform-random legal action for the returned state
        into ``next_out`` (default ``self.next_actions``; must not alias ``actions``)."""
        if (actions.device != self.device or actions.dtype not in (torch.int64, torch.int32)
                or not actions.is_contiguous() or actions.numel() < self.n):
            raise ValueError(f"actions: expected {self.n} contiguous int64/int32 policy indices on {self.device}, got "
                             f"{tuple(actions.shape)} {actions.dtype} on {actions.device}")
        obs = (self.obs if obs is None else obs) if write_obs else None
        mask = (self.mask if mask is None else mask) if write_mask else None
        op, os_ = self._obs_args(obs)
        mp, ms = self._mask_args(mask)
        nxt = None
        if random_actions:
            nxt = self.next_actions if next_out is None else next_out
            if actions.dtype == torch.int32 and nxt.dtype == torch.int64:
                nxt = nxt.view(torch.int32)[: self.n]
            if nxt.data_ptr() == actions.data_ptr():
                raise ValueError("next_out must not alias actions (the kernel reads one while writing the other)")
        self.step_index += 1  # after validation: a rejected call leaves the RNG counter where it was
        if self.step_streams > 1:
            self._step_ranges(actions, op, os_, mp, ms, None, 0, self.reward, self.done, nxt)
        else:
            nv.check(self._L.kz_step(self.state.data_ptr(), self.n, self.hist_cap, actions.data_ptr(),
                                     int(actions.dtype == torch.int64), op, os_, mp, ms, self.reward.data_ptr(),
                                     self.done.data_ptr(), self.reason.data_ptr(), self.winner.data_ptr(),
                                     self.ep_len.data_ptr(), self.legal_count.data_ptr(),
                                     nxt.data_ptr() if nxt is not None else None, self.seed, self.step_index,
                                     self.env_offset, int(self.auto_reset), self._sp()), "kz_step")
        return {"obs": obs, "mask": mask, "reward": self.reward, "done": self.done, "reason": self.reason,
                "winner": self.winner, "ep_len": self.ep_len, "legal_count": self.legal_count}

    def step_rollout(self, actions: torch.Tensor, obs: Optional[torch.Tensor], bitmap: torch.Tensor,
                     reward: Optional[torch.Tensor] = None, done: Optional[torch.Tensor] = None,
                     random_actions: bool = False, next_out: Optional[torch.Tensor] = None,
                     cobs: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """``step`` in its rollout form (kz_step_rollout): the successor's legal set is written as the 13,527-bit legal
        bitmap (int32 [n, 448] rows, e.g. ``RolloutBuffer.bitmaps[t + 1]``) instead of the byte mask -- 1.8 KB instead
        of 13.5 KB per game here and in every later pass of the sampler / PPO update over it.  ``obs`` as in ``step``
        (None: no observation rows); ``reward`` / ``done`` may point into rollout storage (fp32 / uint8 [n]); ``cobs``
        (int32 [n, 40]) receives the compact observations the policy network's input layer reads (kz_cobs_conv_*)."""
        if (actions.device != self.device or actions.dtype not in (torch.int64, torch.int32)
                or not actions.is_contiguous() or actions.numel() < self.n):
            raise ValueError(f"actions: expected {self.n} contiguous int64/int32 policy indices on {self.device}")
        bp, bs = self._bitmap_args(bitmap)
        op, os_ = self._obs_args(obs)
        cp = self._cobs_arg(cobs)
        reward = self.reward if reward is None else reward
        done = self.done if done is None else done
        for t, dt, what in ((reward, torch.float32, "reward"), (done, torch.uint8, "done")):
            if t.device != self.device or t.dtype != dt or not t.is_contiguous() or t.numel() < self.n:
                raise ValueError(f"{what}: expected a contiguous {dt} [{self.n}] tensor on {self.device}")
        nxt = None
        if random_actions:
            nxt = self.next_actions if next_out is None else next_out
            if actions.dtype == torch.int32 and nxt.dtype == torch.int64:
                nxt = nxt.view(torch.int32)[: self.n]
            if nxt.data_ptr() == actions.data_ptr():
                raise ValueError("next_out must not alias actions (the kernel reads one while writing the other)")
        self.step_index += 1
        if self.step_streams > 1:
            self._step_ranges(actions, op, os_, None, 0, bp, bs, reward, done, nxt, cp)
        else:
            nv.check(self._L.kz_step_rollout(self.state.data_ptr(), self.n, self.hist_cap, actions.data_ptr(),
                                             int(actions.dtype == torch.int64), op, os_, bp, bs, cp, reward.data_ptr(),
                                             done.data_ptr(), self.reason.data_ptr(), self.winner.data_ptr(),
                                             self.ep_len.data_ptr(), self.legal_count.data_ptr(),
                                             nxt.data_ptr() if nxt is not None else None, self.seed, self.step_index,
                                             self.env_offset, int(self.auto_reset), self._sp()), "kz_step_rollout")
        return {"obs": obs, "bitmap": bitmap, "reward": reward, "done": done, "reason": self.reason,
                "winner": self.winner, "ep_len": self.ep_len, "legal_count": self.legal_count}

    def _step_ranges(self, actions, op, os_, mp, ms, bp, bs, reward, done, nxt, cp=None) -> None:
        """One step as ``step_streams`` concurrent kz_step_range launches over disjoint ranges of games (each with its own
        work counter), forked from and joined back into the current stream."""
        cur = torch.cuda.current_stream(self.device)
        k, per = self.step_streams, self.n // self.step_streams
        fork = torch.cuda.Event()
        fork.record(cur)
        for i, st in enumerate(self._side_streams):
            st.wait_event(fork)
            nv.check(self._L.kz_step_range(self.state.data_ptr(), self.n, self.hist_cap, i * per, per, i, actions.data_ptr(),
                                           int(actions.dtype == torch.int64), op, os_, mp, ms, bp, bs, cp, reward.data_ptr(),
                                           done.data_ptr(), self.reason.data_ptr(), self.winner.data_ptr(),
                                           self.ep_len.data_ptr(), self.legal_count.data_ptr(),
                                           nxt.data_ptr() if nxt is not None else None, self.seed, self.step_index,
                                           self.env_offset, int(self.auto_reset), st.cuda_stream), "kz_step_range")
            join = torch.cuda.Event()
            join.record(st)
            cur.wait_event(join)

    def legal_bitmap(self, bitmap: torch.Tensor, obs: Optional[torch.Tensor] = None, cobs: Optional[torch.Tensor] = None):
        """Legal bitmap (and optionally the observation rows / compact observations) of the CURRENT positions
        (kz_legal_bitmap)."""
        bp, bs = self._bitmap_args(bitmap)
        op, os_ = self._obs_args(obs)
        nv.check(self._L.kz_legal_bitmap(self.state.data_ptr(), self.n, self.hist_cap, op, os_, bp, bs, self._cobs_arg(cobs),
                                         self.legal_count.data_ptr(), self._sp()), "kz_legal_bitmap")
        return obs, bitmap

    def _cobs_arg(self, cobs: Optional[torch.Tensor]):
        if cobs is None:
            return None
        if (cobs.device != self.device or cobs.dtype != torch.int32 or cobs.dim() != 2 or cobs.shape[0] < self.n
                or cobs.shape[1] != nv.COBS_WORDS or not cobs.is_contiguous()):
            raise ValueError(f"cobs: expected a contiguous int32 [{self.n}, {nv.COBS_WORDS}] tensor on {self.device}")
        return cobs.data_ptr()

    def _bitmap_args(self, bitmap: torch.Tensor):
        if (bitmap.device != self.device or bitmap.dtype != torch.int32 or bitmap.dim() != 2 or bitmap.shape[0] < self.n
                or bitmap.shape[1] != nv.BITMAP_WORDS or bitmap.stride(1) != 1 or bitmap.stride(0) % 4
                or bitmap.data_ptr() % 16):
            raise ValueError(f"bitmap: expected int32 [{self.n}, {nv.BITMAP_WORDS}] rows (16-byte aligned) on {self.device}")
        return bitmap.data_ptr(), bitmap.stride(0)

    def step_compact(self, actions: torch.Tensor, bitmap: Optional[torch.Tensor] = None, random_actions: bool = False,
                     next_out: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """First half of the split pipeline (kz_step_compact): ``step`` without the mask / observation rows; the
        successor's legal bitmap goes to ``bitmap`` (int32 [n, 448], default ``self.bitmap``) for ``expand``."""
        if (actions.device != self.device or actions.dtype not in (torch.int64, torch.int32)
                or not actions.is_contiguous() or actions.numel() < self.n):
            raise ValueError(f"actions: expected {self.n} contiguous int64/int32 policy indices on {self.device}")
        if bitmap is None:
            if getattr(self, "bitmap", None) is None:
                self.bitmap = torch.zeros((self.n, nv.BITMAP_WORDS), dtype=torch.int32, device=self.device)
            bitmap = self.bitmap
        self._check_bitmap(bitmap)
        nxt = None
        if random_actions:
            nxt = self.next_actions if next_out is None else next_out
            if actions.dtype == torch.int32 and nxt.dtype == torch.int64:
                nxt = nxt.view(torch.int32)[: self.n]
            if nxt.data_ptr() == actions.data_ptr():
                raise ValueError("next_out must not alias actions (the kernel reads one while writing the other)")
        self.step_index += 1
        nv.check(self._L.kz_step_compact(self.state.data_ptr(), self.n, self.hist_cap, actions.data_ptr(),
                                         int(actions.dtype == torch.int64), bitmap.data_ptr(), self.reward.data_ptr(),
                                         self.done.data_ptr(), self.reason.data_ptr(), self.winner.data_ptr(),
                                         self.ep_len.data_ptr(), self.legal_count.data_ptr(),
                                         nxt.data_ptr() if nxt is not None else None, self.seed, self.step_index,
                                         self.env_offset, int(self.auto_reset), self._sp()), "kz_step_compact")
        return {"bitmap": bitmap, "reward": self.reward, "done": self.done, "reason": self.reason,
                "winner": self.winner, "ep_len": self.ep_len, "legal_count": self.legal_count}

    def expand(self, bitmap: Optional[torch.Tensor] = None, obs: Optional[torch.Tensor] = None,
               mask: Optional[torch.Tensor] = None):
        """Second half of the split pipeline (kz_expand): mask and observation rows of the current positions from the
        state and the bitmap ``step_compact`` left."""
        bitmap = self.bitmap if bitmap is None else bitmap
        self._check_bitmap(bitmap)
        obs = self.obs if obs is None else obs
        mask = self.mask if mask is None else mask
        op, os_ = self._obs_args(obs)
        mp, ms = self._mask_args(mask)
        nv.check(self._L.kz_expand(self.state.data_ptr(), self.n, self.hist_cap, bitmap.data_ptr(), op, os_, mp, ms,
                                   self._sp()), "kz_expand")
        return obs, mask

    def _check_bitmap(self, bitmap: torch.Tensor) -> None:
        if (bitmap.device != self.device or bitmap.dtype != torch.int32 or not bitmap.is_contiguous()
                or tuple(bitmap.shape) != (self.n, nv.BITMAP_WORDS)):
            raise ValueError(f"bitmap: expected a contiguous int32 [{self.n}, {nv.BITMAP_WORDS}] tensor on {self.device}")

    def load_positions(self, boards, hands, side, move_count, max_moves=None, eval_termination: bool = True):
        """ShogiGame.from_sfen for every env from already-parsed arrays (numpy or torch)."""
        d = self.device
        b = torch.as_tensor(np.ascontiguousarray(boards), dtype=torch.int8).to(d).contiguous()
        h = torch.as_tensor(np.ascontiguousarray(hands), dtype=torch.uint8).to(d).contiguous()
        s = torch.as_tensor(np.ascontiguousarray(side), dtype=torch.uint8).to(d).contiguous()
        mc = torch.as_tensor(np.ascontiguousarray(move_count), dtype=torch.int32).to(d).contiguous()
        if max_moves is None:
            max_moves = np.full(self.n, self.max_moves, np.int32)
        mm = torch.as_tensor(np.ascontiguousarray(max_moves), dtype=torch.int32).to(d).contiguous()
        nv.require(b.shape == (self.n, 81) and h.shape == (self.n, 14), "b.shape == (self.n, 81) and h.shape == (self.n, 14)")
        nv.check(self._L.kz_load_positions(self.state.data_ptr(), self.n, self.hist_cap, b.data_ptr(), h.data_ptr(),
                                           s.data_ptr(), mc.data_ptr(), mm.data_ptr(), self._sp()), "kz_load_positions")
        self.refresh(eval_termination=eval_termination)
        return self.obs, self.mask

    def export(self):
        """-> boards int8 [n,81], hands uint8 [n,14], meta int32 [n,8] (side, move_count, max_moves, status,
        winner, error bits, plies since reset, finished episodes) as torch tensors on the device."""
        d = self.device
        b = torch.empty((self.n, 81), dtype=torch.int8, device=d)
        h = torch.empty((self.n, 14), dtype=torch.uint8, device=d)
        m = torch.empty((self.n, 8), dtype=torch.int32, device=d)
        nv.check(self._L.kz_export_positions(self.state.data_ptr(), self.n, self.hist_cap, b.data_ptr(), h.data_ptr(),
                                             m.data_ptr(), self._sp()), "kz_export_positions")
        return b, h, m

    def to_games(self, env_ids=None):
        """Host snapshots (``shogi.sfen.HostPosition``) of the selected device games: SFEN / board text / KIF
        headers for debugging, parity dumps and displays.  One export launch + one D2H."""
        from .shogi.sfen import HostPosition
        b, h, m = [x.cpu().numpy() for x in self.export()]
        ids = range(self.n) if env_ids is None else [int(i) for i in env_ids]
        return [HostPosition(b[i], h[i], m[i, 0], m[i, 1], status=m[i, 3], winner=m[i, 4]) for i in ids]

    def to_sfen(self, env_ids=None):
        """SFEN strings of the selected device games (shogi_game_io.py:312-379 format)."""
        return [g.to_sfen_string() for g in self.to_games(env_ids)]

    def load_sfens(self, sfens, eval_termination: bool = True):
        """ShogiGame.from_sfen (shogi_game.py:283-345) for a whole batch: one SFEN per env."""
        from .shogi.sfen import pack_sfen
        nv.require(len(sfens) == self.n, "len(sfens) == self.n")
        packed = [pack_sfen(s) for s in sfens]
        return self.load_positions(np.stack([p[0] for p in packed]), np.stack([p[1] for p in packed]),
                                   np.asarray([p[2] for p in packed], np.uint8), np.asarray([p[3] for p in packed], np.int32),
                                   eval_termination=eval_termination)

    def errors(self, clear: bool = False) -> torch.Tensor:
        out = torch.empty(self.n, dtype=torch.int32, device=self.device)
        nv.check(self._L.kz_errors(self.state.data_ptr(), self.n, self.hist_cap, out.data_ptr(), int(clear), self._sp()),
                 "kz_errors")
        return out

    def piece_targets(self, squares) -> torch.Tensor:
        """Pseudo-legal target sets (81-bit, 3 x uint32 per env) of the pieces on ``squares`` [n]."""
        sq = torch.as_tensor(np.ascontiguousarray(squares), dtype=torch.int32).to(self.device).contiguous()
        out = torch.zeros((self.n, 3), dtype=torch.int32, device=self.device)
        nv.check(self._L.kz_piece_targets(self.state.data_ptr(), self.n, self.hist_cap, sq.data_ptr(), out.data_ptr(),
                                          self._sp()), "kz_piece_targets")
        return out
