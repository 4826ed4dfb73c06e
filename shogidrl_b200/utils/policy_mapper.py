"""PolicyOutputMapper -- the 13,527-action enumeration of keisei/utils/utils.py:180-467.

The reference builds two Python tables; the enumeration has a closed form, which is also what the kernels
use: board move (from, to, promote) -> ((from*80 + to - (to > from)) * 2 + promote), drop (to, type) ->
12960 + to*7 + type.  ``idx_to_move`` / ``move_to_idx`` are still exposed for API compatibility."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from ..shogi.definitions import MoveTuple, PieceType

NUM_BOARD_ACTIONS = 81 * 80 * 2
NUM_ACTIONS = NUM_BOARD_ACTIONS + 81 * 7


def move_to_index(move: MoveTuple) -> Optional[int]:
    """Closed-form policy index of a MoveTuple, or None if it is not in the enumeration."""
    if not isinstance(move, tuple) or len(move) != 5:
        return None
    fr, fc, tr, tc, last = move
    if not (isinstance(tr, int) and isinstance(tc, int) and 0 <= tr < 9 and 0 <= tc < 9):
        return None
    to = tr * 9 + tc
    if fr is None and fc is None:
        value = getattr(last, "value", None)  # PieceType from either package compares by value
        if isinstance(last, bool) or value is None or not (0 <= value < 7):
            return None
        return NUM_BOARD_ACTIONS + to * 7 + value
    if not (isinstance(fr, int) and isinstance(fc, int) and 0 <= fr < 9 and 0 <= fc < 9 and isinstance(last, bool)):
        return None
    frm = fr * 9 + fc
    if frm == to:
        return None
    return (frm * 80 + to - (1 if to > frm else 0)) * 2 + int(last)


def index_to_move(idx: int) -> MoveTuple:
    if idx < NUM_BOARD_ACTIONS:
        pair, promo = divmod(idx, 2)
        frm, t = divmod(pair, 80)
        to = t + (1 if t >= frm else 0)
        return (frm // 9, frm % 9, to // 9, to % 9, bool(promo))
    to, pt = divmod(idx - NUM_BOARD_ACTIONS, 7)
    return (None, None, to // 9, to % 9, PieceType(pt))


class PolicyOutputMapper:
    """Maps Shogi moves to/from policy network output indices."""

    def __init__(self) -> None:
        self.idx_to_move: List[MoveTuple] = [index_to_move(i) for i in range(NUM_ACTIONS)]
        self.move_to_idx: Dict[MoveTuple, int] = {m: i for i, m in enumerate(self.idx_to_move)}

    def get_total_actions(self) -> int:
        return NUM_ACTIONS

    def shogi_move_to_policy_index(self, move: MoveTuple) -> int:
        idx = move_to_index(move)
        if idx is None:
            raise ValueError(f"Move {move} (type: {type(move)}, element types: {[type(el) for el in move]}) "
                             f"not found in PolicyOutputMapper's known moves. Known keys example: {self.idx_to_move[0]}")
        return idx

    def policy_index_to_shogi_move(self, idx: int) -> MoveTuple:
        if 0 <= idx < NUM_ACTIONS:
            return self.idx_to_move[idx]
        raise IndexError(f"Policy index {idx} is out of bounds (0-{NUM_ACTIONS - 1}).")

    def action_idx_to_shogi_move(self, action_idx: int) -> MoveTuple:
        if 0 <= action_idx < NUM_ACTIONS:
            return self.idx_to_move[action_idx]
        raise IndexError(f"Action index {action_idx} is out of bounds for idx_to_move (size {NUM_ACTIONS}).")

    def get_legal_mask(self, legal_shogi_moves: Sequence[MoveTuple], device: torch.device) -> torch.Tensor:
        """bool[13527] mask on ``device`` (utils.py:310-336): one scatter instead of per-element writes."""
        idx = []
        for move in legal_shogi_moves:
            i = move_to_index(move)
            if i is None:
                raise ValueError(f"CRITICAL: Legal move {move} could not be mapped to policy index. This indicates "
                                 f"incomplete move coverage in PolicyOutputMapper which will corrupt experiments. "
                                 f"Original error: move not in the 13,527-action enumeration")
            idx.append(i)
        mask = torch.zeros(NUM_ACTIONS, dtype=torch.bool, device=device)
        if idx:
            mask[torch.as_tensor(idx, dtype=torch.int64, device=device)] = True
        return mask

    # ---- USI helpers (utils.py:338-467)
    @staticmethod
    def _usi_sq(r: int, c: int) -> str:
        if not (0 <= r <= 8 and 0 <= c <= 8):
            raise ValueError(f"Invalid square coordinates: ({r}, {c})")
        return f"{9 - c}{chr(ord('a') + r)}"

    @staticmethod
    def _get_usi_char_for_drop(piece_type: PieceType) -> str:
        if getattr(piece_type, "value", 99) > 6:
            raise ValueError(f"Piece type {piece_type} cannot be dropped or is not a recognized droppable piece.")
        return "PLNSGBR"[piece_type.value]

    def shogi_move_to_usi(self, move_tuple: MoveTuple) -> str:
        if len(move_tuple) == 5 and isinstance(move_tuple[4], bool):
            fr, fc, tr, tc, promote = move_tuple
            if not all(isinstance(v, int) for v in (fr, fc, tr, tc)):
                raise ValueError("Invalid coordinates in BoardMoveTuple for USI conversion.")
            return self._usi_sq(fr, fc) + self._usi_sq(tr, tc) + ("+" if promote else "")
        if len(move_tuple) == 5 and hasattr(move_tuple[4], "value") and not isinstance(move_tuple[4], bool):
            _, _, tr, tc, pt = move_tuple
            if not all(isinstance(v, int) for v in (tr, tc)):
                raise ValueError("Invalid coordinates in DropMoveTuple for USI conversion.")
            try:
                ch = self._get_usi_char_for_drop(pt)
            except ValueError as e:
                raise ValueError(f"Invalid piece type for drop in USI conversion: {pt.name}") from e
            return f"{ch}*{self._usi_sq(tr, tc)}"
        raise ValueError(f"Unrecognized move_tuple format for USI conversion: length {len(move_tuple)}, "
                         f"last element type {type(move_tuple[-1]) if move_tuple else 'N/A'}")

    def usi_to_shogi_move(self, usi_move_str: str) -> MoveTuple:
        if not isinstance(usi_move_str, str) or len(usi_move_str) < 4:
            raise ValueError(f"Invalid USI move string format: {usi_move_str}")

        def sq(s: str):
            if not (len(s) == 2 and s[0].isdigit() and s[1].isalpha()):
                raise ValueError(f"Invalid USI square format: {s}")
            r, c = ord(s[1]) - ord("a"), 9 - int(s[0])
            if not (0 <= r <= 8 and 0 <= c <= 8):
                raise ValueError(f"Square coordinates out of bounds: {s} -> ({r}, {c})")
            return r, c

        if usi_move_str[1] == "*":
            if len(usi_move_str) != 4:
                raise ValueError(f"Invalid USI drop move string length: {usi_move_str}")
            k = "PLNSGBR".find(usi_move_str[0])
            if k < 0:
                raise ValueError(f"Invalid piece character for drop: {usi_move_str[0]}")
            r, c = sq(usi_move_str[2:])
            return (None, None, r, c, PieceType(k))
        if len(usi_move_str) > 5:
            raise ValueError(f"Invalid USI board move string length: {usi_move_str}")
        if len(usi_move_str) == 5 and usi_move_str[4] != "+":
            raise ValueError(f"Invalid promotion character in USI move: {usi_move_str}")
        fr, fc = sq(usi_move_str[0:2])
        tr, tc = sq(usi_move_str[2:4])
        return (fr, fc, tr, tc, len(usi_move_str) == 5)

    def action_idx_to_usi_move(self, action_idx: int, _board=None) -> str:
        return self.shogi_move_to_usi(self.action_idx_to_shogi_move(action_idx))
