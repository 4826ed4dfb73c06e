"""ShogiGame -- scalar facade with the API of ``keisei.shogi.ShogiGame`` (keisei/shogi/shogi_game.py) over an
N = 1 device game.

Host attributes (``board``, ``hands``, ``current_player``, ``move_count``, ...) are the authoritative state of the
facade, exactly as in the reference, so code that mutates them directly keeps working (the reference's tests do:
tests/shogi/test_shogi_game_core_logic.py:70,199,210,1069).  Every rules computation -- legal moves, make_move,
termination, observation, in-check, uchifuzume -- is pushed to the CUDA engine through ``VecShogiEnv``; there
is no host implementation of the rules and no CPU fallback.  The repetition rule is evaluated on the host from
``move_history`` state hashes, like check_for_sennichite (shogi_rules_logic.py:638-695), because the facade's
history can be edited by ``undo_move``.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import sfen as _sfen
from .definitions import (BASE_TO_PROMOTED_TYPE, PIECE_TYPE_TO_HAND_TYPE, PROMOTED_TO_BASE_TYPE, Color, MoveTuple,
                          Piece, PieceType, TerminationReason, get_unpromoted_types)
from .. import _native as nv

_REASON_TEXT = {1: TerminationReason.CHECKMATE.value, 2: TerminationReason.STALEMATE.value,
                3: TerminationReason.MAX_MOVES_EXCEEDED.value, 4: TerminationReason.REPETITION.value}
_HAND_TYPES = get_unpromoted_types()


def index_to_move(idx: int):  # late import: utils.policy_mapper itself imports shogi.definitions
    from ..utils.policy_mapper import index_to_move as f
    return f(idx)


def move_to_index(move):
    from ..utils.policy_mapper import move_to_index as f
    return f(move)


class ShogiGame:
    """Shogi game state + operations; rules run on the GPU engine."""

    def __init__(self, max_moves_per_game: int = 500, device: Union[str, torch.device] = "cuda") -> None:
        self.board: List[List[Optional[Piece]]]
        self.hands: Dict[int, Dict[PieceType, int]]
        self.current_player: Color = Color.BLACK
        self.move_count: int = 0
        self.game_over: bool = False
        self.winner: Optional[Color] = None
        self.termination_reason: Optional[str] = None
        self.move_history: List[Dict[str, Any]] = []
        self.board_history: List[Tuple] = []
        self._max_moves_this_game = max_moves_per_game
        self._initial_board_setup_done = False
        self._seed_value: Optional[Any] = None
        self._device = device
        self._env = None  # created on first use so that constructing a game does not need the GPU yet
        self._reset_state()

    # ------------------------------------------------------------------ device plumbing
    def _engine(self):
        if self._env is None:
            from ..vec_env import VecShogiEnv
            self._env = VecShogiEnv(1, max_moves_per_game=min(max(self._max_moves_this_game, 1), 65535),
                                    device=self._device, auto_reset=False, hist_cap=0)
            d = self._env.device
            self._mask1 = torch.zeros((1, nv.NUM_ACTIONS), dtype=torch.uint8, device=d)
            self._chk1 = torch.zeros(1, dtype=torch.uint8, device=d)
        return self._env

    def _encode(self, side: Optional[Color] = None, board=None, hands=None):
        board = self.board if board is None else board
        hands = self.hands if hands is None else hands
        b = np.zeros((1, 81), np.int8)
        for r in range(9):
            row = board[r]
            for c in range(9):
                p = row[c]
                if p is not None:
                    b[0, r * 9 + c] = 1 + p.type.value + 14 * p.color.value
        h = np.zeros((1, 14), np.uint8)
        for color in (0, 1):
            for pt, cnt in hands[color].items():
                if pt.value < 7 and cnt > 0:
                    h[0, color * 7 + pt.value] = min(int(cnt), 255)
        s = np.asarray([(self.current_player if side is None else side).value], np.uint8)
        return b, h, s

    def _push(self, side: Optional[Color] = None, board=None, hands=None, eval_termination: bool = False):
        """Upload the host state (optionally with another side to move / board) and recompute mask + obs."""
        env = self._engine()
        b, h, s = self._encode(side, board, hands)
        env.load_positions(b, h, s, np.asarray([self.move_count], np.int32),
                           np.asarray([min(max(self._max_moves_this_game, 0), 65535)], np.int32),
                           eval_termination=eval_termination)
        return env

    def _pull(self, env) -> np.ndarray:
        b, h, m = [x.cpu().numpy() for x in env.export()]
        self.board = [[Piece.from_code(int(b[0, r * 9 + c])) for c in range(9)] for r in range(9)]
        for color in (0, 1):
            for pt in _HAND_TYPES:
                self.hands[color][pt] = int(h[0, color * 7 + pt.value])
        return m[0]

    # ------------------------------------------------------------------ basics
    def seed(self, seed_value: Optional[Any] = None) -> "ShogiGame":
        self._seed_value = seed_value
        return self

    @property
    def max_moves_per_game(self) -> int:
        return self._max_moves_this_game

    def _setup_initial_board(self) -> None:
        self.board = _sfen.board_from_sfen_segment("lnsgkgsnl/1r5b1/ppppppppp/9/9/9/PPPPPPPPP/1B5R1/LNSGKGSNL")

    def reset(self) -> np.ndarray:
        """ShogiGame.reset (shogi_game.py:113-130): start position; returns its observation (from the engine)."""
        self._reset_state()
        return self.get_observation()

    def _reset_state(self) -> None:
        self._setup_initial_board()
        self.hands = {0: {pt: 0 for pt in _HAND_TYPES}, 1: {pt: 0 for pt in _HAND_TYPES}}
        self.current_player = Color.BLACK
        self.move_count = 0
        self.game_over = False
        self.winner = None
        self.termination_reason = None
        self.move_history = []
        self.board_history = [self._board_state_hash()]
        self._initial_board_setup_done = True

    def is_on_board(self, row: int, col: int) -> bool:
        return 0 <= row < 9 and 0 <= col < 9

    def get_piece(self, row: int, col: int) -> Optional[Piece]:
        return self.board[row][col] if self.is_on_board(row, col) else None

    def set_piece(self, row: int, col: int, piece: Optional[Piece]) -> None:
        if self.is_on_board(row, col):
            self.board[row][col] = piece

    def to_string(self) -> str:
        return _sfen.game_to_text(self)

    def __deepcopy__(self, memo: Dict[int, Any]):
        if id(self) in memo:
            return memo[id(self)]
        g = self.__class__.__new__(self.__class__)
        memo[id(self)] = g
        g.board = [[None if p is None else Piece(p.type, p.color) for p in row] for row in self.board]
        g.hands = {k: dict(v) for k, v in self.hands.items()}
        g.current_player = self.current_player
        g.move_count = self.move_count
        g.game_over = self.game_over
        g.winner = self.winner
        g.termination_reason = self.termination_reason
        g.move_history = []  # a deep copy forgets repetition history (shogi_game.py:165-169)
        g._max_moves_this_game = self._max_moves_this_game
        g._initial_board_setup_done = self._initial_board_setup_done
        g._seed_value = self._seed_value
        g._device = self._device
        g._env = None
        g.board_history = [g._board_state_hash()]
        return g

    # ------------------------------------------------------------------ device-backed queries
    def get_observation(self) -> np.ndarray:
        """46x9x9 fp32 observation (shogi_game_io.py:434-539), computed by the engine."""
        env = self._push()
        return env.obs[0].cpu().numpy()

    def get_state(self) -> np.ndarray:
        return self.get_observation()

    def _legal_indices(self, side: Optional[Color] = None) -> np.ndarray:
        env = self._push(side)
        return torch.nonzero(env.mask[0], as_tuple=False).flatten().cpu().numpy()

    def get_legal_moves(self) -> List[MoveTuple]:
        idx = self._legal_indices()
        # The reference's simulate/undo loop clears the terminal flags whenever it tried at least one candidate
        # (shogi_move_execution.py:218-221); any position with a pseudo-legal candidate triggers it.  A finished
        # game with no candidates at all is impossible except on an empty board, so mirror the common case.
        if self.game_over and self._has_any_candidate():
            self.game_over, self.winner, self.termination_reason = False, None, None
        return [index_to_move(int(i)) for i in idx]

    def _has_any_candidate(self) -> bool:
        me = self.current_player
        if any(p is not None and p.color == me for row in self.board for p in row):
            return True
        return any(cnt > 0 for cnt in self.hands[me.value].values())

    def is_in_check(self, color: Color, debug_recursion: bool = False) -> bool:
        env = self._engine()
        b, h, s = self._encode(color)
        env.load_positions(b, h, s, np.asarray([self.move_count], np.int32), np.asarray([65535], np.int32),
                           eval_termination=False)
        env.refresh(mask=self._mask1, in_check=self._chk1)
        return bool(self._chk1.item())

    def find_king(self, color: Color) -> Optional[Tuple[int, int]]:
        for r in range(9):
            for c in range(9):
                p = self.board[r][c]
                if p is not None and p.type == PieceType.KING and p.color == color:
                    return (r, c)
        return None

    def get_king_legal_moves(self, color: Color) -> int:
        if self.find_king(color) is None:
            return 0
        n = 0
        for i in self._legal_indices(color):
            fr, fc, _, _, _ = index_to_move(int(i))
            if fr is None:
                continue
            p = self.board[fr][fc]
            n += int(p is not None and p.type == PieceType.KING)
        return n

    def is_nifu(self, color: Color, col: int) -> bool:
        return any((p := self.board[r][col]) is not None and p.type == PieceType.PAWN and p.color == color
                   for r in range(9))

    def is_uchi_fu_zume(self, drop_row: int, drop_col: int, color: Color) -> bool:
        """check_for_uchi_fu_zume (shogi_rules_logic.py:275-359): with the pawn placed, the opponent is in check
        and has no legal move.  Both facts come from one engine refresh of the modified position."""
        if self.board[drop_row][drop_col] is not None or self.hands[color.value].get(PieceType.PAWN, 0) <= 0:
            return False
        opp = color.opponent()
        if self.find_king(opp) is None:
            return False
        board = [row[:] for row in self.board]
        board[drop_row][drop_col] = Piece(PieceType.PAWN, color)
        hands = {k: dict(v) for k, v in self.hands.items()}
        hands[color.value][PieceType.PAWN] -= 1
        env = self._engine()
        b, h, s = self._encode(opp, board, hands)
        env.load_positions(b, h, s, np.asarray([0], np.int32), np.asarray([65535], np.int32), eval_termination=False)
        env.refresh(mask=self._mask1, in_check=self._chk1)
        return bool(self._chk1.item()) and int(env.legal_count.item()) == 0

    def can_drop_piece(self, piece_type: PieceType, row: int, col: int, player_color: Color) -> bool:
        if (piece_type == PieceType.KING or not self.is_on_board(row, col)
                or self.hands[player_color.value].get(piece_type, 0) <= 0 or self.board[row][col] is not None):
            return False
        last = 0 if player_color == Color.BLACK else 8
        second = 1 if player_color == Color.BLACK else 7
        if piece_type == PieceType.PAWN:
            return not (self.is_nifu(player_color, col) or row == last or self.is_uchi_fu_zume(row, col, player_color))
        if piece_type == PieceType.LANCE:
            return row != last
        if piece_type == PieceType.KNIGHT:
            return row not in (last, second)
        return True

    def is_in_promotion_zone(self, row: int, color: Color) -> bool:
        return 0 <= row <= 2 if color == Color.BLACK else 6 <= row <= 8

    def get_individual_piece_moves(self, piece: Piece, r_from: int, c_from: int) -> List[Tuple[int, int]]:
        """Pseudo-legal targets of ``piece`` placed on (r_from, c_from) (generate_piece_potential_moves,
        shogi_rules_logic.py:82-208), from the engine's kz_piece_targets."""
        board = [row[:] for row in self.board]
        board[r_from][c_from] = piece
        env = self._engine()
        b, h, s = self._encode(None, board)
        env.load_positions(b, h, s, np.asarray([0], np.int32), np.asarray([65535], np.int32), eval_termination=False)
        w = env.piece_targets(np.asarray([r_from * 9 + c_from], np.int32))[0].cpu().numpy().astype(np.uint32)
        return [(sq // 9, sq % 9) for sq in range(81) if (int(w[sq >> 5]) >> (sq & 31)) & 1]

    # ------------------------------------------------------------------ hashing / rewards / termination
    def _board_state_hash(self) -> tuple:
        board_tuple = tuple(tuple((p.type.value, p.color.value) if p else None for p in row) for row in self.board)
        hands_tuple = tuple(tuple(sorted((pt.value, c) for pt, c in self.hands[col].items() if c > 0)) for col in (0, 1))
        return (board_tuple, hands_tuple, self.current_player.value)

    def get_board_state_hash(self) -> tuple:
        return self._board_state_hash()

    def get_reward(self, perspective_player_color: Optional[Color] = None) -> float:
        if perspective_player_color is None:
            raise ValueError("perspective_player_color must be provided to get_reward.")
        if not self.game_over or self.winner is None:
            return 0.0
        return 1.0 if self.winner == perspective_player_color else -1.0

    def is_sennichite(self) -> bool:
        if not self.move_history:
            return False
        last = self.move_history[-1].get("state_hash")
        if not last:
            return False
        return sum(1 for rec in self.move_history if rec.get("state_hash") == last) >= 4

    def _check_and_update_termination_status(self, player_who_just_moved: Color) -> None:
        """_check_and_update_termination_status (shogi_game.py:408-450): engine for mate / stalemate / max moves,
        host for repetition."""
        if self.game_over:
            return
        env = self._push(eval_termination=True)
        m = env.export()[2][0].cpu().numpy()
        self._apply_status(int(m[3]), player_who_just_moved)

    def _apply_status(self, status: int, mover: Color) -> None:
        if status == 1:
            self.game_over, self.winner = True, mover
        elif status in (2, 3):
            self.game_over, self.winner = True, None
        elif status == 0 and self.is_sennichite():
            status = 4
            self.game_over, self.winner = True, None
        if status:
            self.termination_reason = _REASON_TEXT[status]

    # ------------------------------------------------------------------ moves
    def _validate_move_tuple_format(self, move_tuple: MoveTuple) -> None:
        ok = isinstance(move_tuple, tuple) and len(move_tuple) == 5
        if ok:
            a, b, c, d, e = move_tuple
            board = all(isinstance(v, int) for v in (a, b, c, d)) and isinstance(e, bool)
            drop = a is None and b is None and isinstance(c, int) and isinstance(d, int) and isinstance(e, PieceType)
            ok = board or drop
        if not ok:
            raise ValueError(f"Invalid move_tuple format: {move_tuple}")

    def _terminal_tuple(self, mover: Color, obs: Optional[np.ndarray] = None):
        obs = self.get_observation() if obs is None else obs
        reward = 0.0
        if self.game_over and self.winner is not None:
            reward = 1.0 if self.winner == mover else -1.0
        info: Dict[str, Any] = {"reason": self.termination_reason if self.game_over else "Game ongoing"}
        if self.game_over and self.winner is not None:
            info["winner"] = self.winner.name
        return obs, reward, self.game_over, info

    def make_move(self, move_tuple: MoveTuple, is_simulation: bool = False):
        """ShogiGame.make_move (shogi_game.py:574-660).  Real moves run kz_step on the device; simulations only
        edit the host board (they exist for API compatibility, the engine never needs them)."""
        if self.game_over and not is_simulation:
            return self._terminal_tuple(self.current_player.opponent())
        self._validate_move_tuple_format(move_tuple)
        mover = self.current_player
        details = {"move": move_tuple, "is_drop": move_tuple[0] is None, "captured": None,
                   "was_promoted_in_move": False, "original_type_before_promotion": None,
                   "dropped_piece_type": move_tuple[4] if move_tuple[0] is None else None,
                   "original_color_of_moved_piece": None, "player_who_made_the_move": mover,
                   "move_count_before_move": self.move_count, "original_board_state": None,
                   "original_hands_state": None, "state_hash": None}
        if is_simulation:
            return self._simulate(move_tuple, details)

        idx = move_to_index(move_tuple)
        if idx is None:
            raise ValueError(f"Invalid move_tuple format: {move_tuple}")
        if move_tuple[0] is not None:
            fr, fc, tr, tc, promote = move_tuple
            piece = self.get_piece(fr, fc)
            if piece is None:
                raise ValueError(f"Invalid move: No piece at source ({fr},{fc})")
            if piece.color != mover:
                raise ValueError(f"Invalid move: Piece at ({fr},{fc}) does not belong to current player.")
            details["original_type_before_promotion"] = piece.type
            details["original_color_of_moved_piece"] = piece.color
            target = self.get_piece(tr, tc)
            details["captured"] = target.type if target is not None else None
            details["was_promoted_in_move"] = bool(promote)
        env = self._push()
        out = env.step(torch.tensor([idx], dtype=torch.int64, device=env.device))
        err = int(env.errors(clear=True).item())
        if err & 2:
            fr, fc, tr, tc, _ = move_tuple
            raise ValueError(f"Illegal movement pattern: {self.board[fr][fc].type.name} at ({fr},{fc}) cannot move to "
                             f"({tr},{tc}).")
        if err & 1:
            raise ValueError(f"Invalid move: {move_tuple} cannot be applied in this position.")
        meta = self._pull(env)
        self.move_count = int(meta[1])
        self.current_player = Color(int(meta[0]))
        state_hash = self._board_state_hash()
        details["state_hash"] = state_hash
        self.move_history.append(details)
        self.board_history.append(state_hash)
        self._apply_status(int(meta[3]), mover)
        obs = out["obs"][0].cpu().numpy()
        return self._terminal_tuple(mover, obs)

    def _simulate(self, move_tuple: MoveTuple, details: Dict[str, Any]) -> Dict[str, Any]:
        mover = self.current_player
        details["original_board_state"] = [[None if p is None else Piece(p.type, p.color) for p in row] for row in self.board]
        details["original_hands_state"] = {k: dict(v) for k, v in self.hands.items()}
        if move_tuple[0] is None:
            _, _, tr, tc, pt = move_tuple
            if pt not in self.hands[mover.value]:
                raise ValueError(f"Attempting to drop {pt} which is not in hand for {mover}")
            self.board[tr][tc] = Piece(pt, mover)
            self.hands[mover.value][pt] -= 1
        else:
            fr, fc, tr, tc, promote = move_tuple
            piece = self.get_piece(fr, fc)
            if piece is None:
                raise ValueError(f"Invalid move: No piece at source ({fr},{fc})")
            if piece.color != mover:
                raise ValueError(f"Invalid move: Piece at ({fr},{fc}) does not belong to current player.")
            # movement-pattern validation is the engine's: a throw-away step on a copy of the state
            env = self._push()
            env.step(torch.tensor([move_to_index(move_tuple)], dtype=torch.int64, device=env.device))
            err = int(env.errors(clear=True).item())
            if err & 2:
                raise ValueError(f"Illegal movement pattern: {piece.type.name} at ({fr},{fc}) cannot move to ({tr},{tc}).")
            details["original_type_before_promotion"] = piece.type
            details["original_color_of_moved_piece"] = piece.color
            target = self.board[tr][tc]
            if target is not None:
                if target.color == mover:
                    raise ValueError(f"Cannot capture own piece at ({tr},{tc}).")
                base = PROMOTED_TO_BASE_TYPE.get(target.type, target.type)
                self.hands[mover.value][base] = self.hands[mover.value].get(base, 0) + 1
                details["captured"] = base
            self.board[tr][tc] = piece
            self.board[fr][fc] = None
            if promote:
                if piece.type not in BASE_TO_PROMOTED_TYPE:
                    raise ValueError(f"Piece type {piece.type} cannot be promoted. Move: {move_tuple}, Piece: {piece}")
                self.board[tr][tc] = Piece(BASE_TO_PROMOTED_TYPE[piece.type], piece.color)
                details["was_promoted_in_move"] = True
        self.current_player = mover.opponent()
        return details

    def undo_move(self, simulation_undo_details: Optional[Dict[str, Any]] = None) -> None:
        if simulation_undo_details:
            d = simulation_undo_details
            if not (isinstance(d.get("original_board_state"), list) and isinstance(d.get("original_hands_state"), dict)
                    and isinstance(d.get("player_who_made_the_move"), Color)
                    and isinstance(d.get("move_count_before_move"), int)):
                raise TypeError("One or more arguments from simulation_undo_details have incorrect types.")
            self.board = [row[:] for row in d["original_board_state"]]
            self.hands = {k: dict(v) for k, v in d["original_hands_state"].items()}
            self.current_player = d["player_who_made_the_move"]
            self.move_count = d["move_count_before_move"]
            self.game_over, self.winner, self.termination_reason = False, None, None
            return
        if not self.move_history:
            return
        last = self.move_history.pop()
        if self.board_history:
            self.board_history.pop()
        self.current_player = last["player_who_made_the_move"]
        self.move_count = last["move_count_before_move"]
        _, _, tr, tc, _ = last["move"]
        me = self.current_player
        if last["is_drop"]:
            pt = last["dropped_piece_type"]
            self.board[tr][tc] = None
            self.hands[me.value][pt] = self.hands[me.value].get(pt, 0) + 1
        else:
            fr, fc = last["move"][0], last["move"][1]
            self.board[fr][fc] = Piece(last["original_type_before_promotion"], last["original_color_of_moved_piece"])
            cap = last["captured"]
            if cap:
                self.board[tr][tc] = Piece(cap, me.opponent())
                hand_type = PIECE_TYPE_TO_HAND_TYPE.get(cap)
                if hand_type is None:
                    raise ValueError(f"Cannot convert captured board type {cap} to hand type during undo.")
                if self.hands[me.value].get(hand_type, 0) > 0:
                    self.hands[me.value][hand_type] -= 1
            else:
                self.board[tr][tc] = None
        self.game_over, self.winner, self.termination_reason = False, None, None

    def test_move(self, move_tuple: MoveTuple) -> bool:
        """ShogiGame.test_move (shogi_game.py:884-967): drops -> can_drop_piece, board moves -> pattern valid."""
        if self.game_over:
            return False
        try:
            self._validate_move_tuple_format(move_tuple)
        except ValueError:
            return False
        try:
            if move_tuple[0] is None:
                return self.can_drop_piece(move_tuple[4], move_tuple[2], move_tuple[3], self.current_player)
            d = self.make_move(move_tuple, is_simulation=True)
            self.undo_move(simulation_undo_details=d)
            return True
        except Exception:
            return False

    # ------------------------------------------------------------------ hands
    def add_to_hand(self, captured_piece: Piece, capturing_player_color: Color) -> None:
        if captured_piece.type == PieceType.KING:
            return
        hand_type = PIECE_TYPE_TO_HAND_TYPE.get(captured_piece.type)
        if hand_type is None:
            raise ValueError(f"Invalid piece type {captured_piece.type} to add to hand.")
        hand = self.hands[capturing_player_color.value]
        hand[hand_type] = hand.get(hand_type, 0) + 1

    def remove_from_hand(self, piece_type: PieceType, color: Color) -> bool:
        if piece_type not in _HAND_TYPES:
            return False
        hand = self.hands[color.value]
        if hand.get(piece_type, 0) > 0:
            hand[piece_type] -= 1
            return True
        return False

    def get_pieces_in_hand(self, color: Color) -> Dict[PieceType, int]:
        return self.hands[color.value].copy()

    # ------------------------------------------------------------------ SFEN
    def sfen_encode_move(self, move_tuple: MoveTuple) -> str:
        return _sfen.encode_move(move_tuple)

    def to_sfen_string(self) -> str:
        return _sfen.game_to_sfen(self)

    def to_sfen(self) -> str:
        return self.to_sfen_string()

    @classmethod
    def from_sfen(cls, sfen_str: str, max_moves_for_game_instance: int = 500, device="cuda") -> "ShogiGame":
        board_s, turn, hands_s, num_s = _sfen.split_sfen(sfen_str)
        try:
            num = int(num_s)
        except ValueError as e:
            raise ValueError(f"Invalid move number in SFEN: '{num_s}'") from e
        if num < 1:
            raise ValueError("SFEN move number must be positive")
        g = cls(max_moves_per_game=max_moves_for_game_instance, device=device)
        g.current_player = Color.BLACK if turn == "b" else Color.WHITE
        g.move_count = num - 1
        g.board = _sfen.board_from_sfen_segment(board_s)
        g.hands = {0: {pt: 0 for pt in _HAND_TYPES}, 1: {pt: 0 for pt in _HAND_TYPES}}
        _sfen.hands_from_sfen_segment(g.hands, hands_s)
        g.move_history = []
        g.board_history = [g._board_state_hash()]
        g._initial_board_setup_done = True
        g._check_and_update_termination_status(g.current_player.opponent())
        return g
