"""Host-side SFEN / USI text I/O for the scalar facade and for loading start positions into device batches.

Behaviour-compatible with keisei/shogi/shogi_game_io.py: SFEN parsing :169-306 (same ValueError cases), SFEN
emission :312-379 (hands in R,B,G,S,N,L,P order), move strings :382-431 and :744-826, text board :542-585.
Text never reaches the GPU: positions are packed into piece-code arrays (``pack_sfen``) before upload."""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Tuple

import numpy as np

from .definitions import (BASE_TO_PROMOTED_TYPE, PROMOTED_TYPES_SET, SYMBOL_TO_PIECE_TYPE, Color, MoveTuple, Piece,
                          PieceType)

_SFEN_RE = re.compile(r"^\s*([^ ]+)\s+([bw])\s+([^ ]+)\s+(\d+)\s*$")
_HAND_RE = re.compile(r"(\d*)([PLNSGBRplnsgbr])")
_HAND_ORDER = [PieceType.ROOK, PieceType.BISHOP, PieceType.GOLD, PieceType.SILVER, PieceType.KNIGHT, PieceType.LANCE,
               PieceType.PAWN]
_BASE_CHAR = {**{PieceType(i): "PLNSGBRK"[i] for i in range(8)},
              **{p: "PLNSGBR"[b.value] for b, p in BASE_TO_PROMOTED_TYPE.items()}}


def split_sfen(sfen_str: str) -> Tuple[str, str, str, str]:
    m = _SFEN_RE.match(sfen_str.strip())
    if not m:
        raise ValueError(f"Invalid SFEN string structure: '{sfen_str}'")
    return m.group(1), m.group(2), m.group(3), m.group(4)


def board_from_sfen_segment(segment: str) -> List[List[Optional[Piece]]]:
    board: List[List[Optional[Piece]]] = [[None] * 9 for _ in range(9)]
    rows = segment.split("/")
    if len(rows) != 9:
        raise ValueError("Expected 9 ranks")
    for r, row in enumerate(rows):
        c = 0
        promo = False
        for ch in row:
            if c >= 9:
                break
            if ch == "+":
                if promo:
                    raise ValueError("Invalid piece character sequence starting with '+'")
                promo = True
            elif ch.isdigit():
                if promo:
                    raise ValueError(f"Invalid SFEN: Digit ('{ch}') cannot immediately follow a promotion token ('+').")
                if ch == "0":
                    raise ValueError("Invalid SFEN piece character for board: 0")
                k = int(ch)
                if not 1 <= k <= 9 - c:
                    raise ValueError(f"Row {r + 1} ('{row}') describes {c + k} columns, expected 9")
                c += k
            else:
                base = SYMBOL_TO_PIECE_TYPE.get(ch.upper())
                if base is None or base in PROMOTED_TYPES_SET:
                    raise ValueError(f"Invalid SFEN piece character for board: {ch}")
                ptype = base
                if promo:
                    if base not in BASE_TO_PROMOTED_TYPE:
                        raise ValueError(f"Invalid promotion: SFEN token '+' applied to non-promotable piece type {base.name}")
                    ptype = BASE_TO_PROMOTED_TYPE[base]
                board[r][c] = Piece(ptype, Color.BLACK if ch.isupper() else Color.WHITE)
                c += 1
                promo = False
        if c != 9:
            raise ValueError(f"Row {r + 1} ('{row}') describes {c} columns, expected 9")
    return board


def hands_from_sfen_segment(hands: Dict[int, Dict[PieceType, int]], segment: str) -> None:
    if segment == "-":
        return
    pos = 0
    seen_white = False
    while pos < len(segment):
        m = _HAND_RE.match(segment, pos)
        if not m:
            if segment[pos:pos + 1] in ("K", "k"):
                raise ValueError("Invalid piece character 'K' or non-droppable piece type in SFEN hands")
            raise ValueError("Invalid character sequence in SFEN hands")
        count = int(m.group(1)) if m.group(1) else 1
        ch = m.group(2)
        if ch.islower():
            seen_white = True
        elif seen_white:
            raise ValueError("Invalid SFEN hands: Black's pieces must precede White's pieces.")
        ptype = SYMBOL_TO_PIECE_TYPE[ch.upper()]
        hand = hands[0 if ch.isupper() else 1]
        hand[ptype] = hand.get(ptype, 0) + count
        pos = m.end()


def piece_char(piece: Piece) -> str:
    if not isinstance(piece, Piece):
        raise TypeError(f"Expected a Piece object, got {type(piece)}")
    s = ("+" if piece.type in PROMOTED_TYPES_SET else "") + _BASE_CHAR[piece.type]
    return s.lower() if piece.color == Color.WHITE else s


def game_to_sfen(game) -> str:
    ranks = []
    for r in range(9):
        out, empty = "", 0
        for c in range(9):
            p = game.board[r][c]
            if p is None:
                empty += 1
            else:
                if empty:
                    out += str(empty)
                    empty = 0
                out += piece_char(p)
        if empty:
            out += str(empty)
        ranks.append(out)
    parts = []
    for color, lower in ((0, False), (1, True)):
        for pt in _HAND_ORDER:
            cnt = game.hands[color].get(pt, 0)
            if cnt > 0:
                ch = _BASE_CHAR[pt].lower() if lower else _BASE_CHAR[pt]
                parts.append((str(cnt) if cnt > 1 else "") + ch)
    turn = "b" if game.current_player == Color.BLACK else "w"
    return f"{'/'.join(ranks)} {turn} {''.join(parts) if parts else '-'} {game.move_count + 1}"


def sq_to_sfen(r: int, c: int) -> str:
    if not (0 <= r <= 8 and 0 <= c <= 8):
        raise ValueError(f"Invalid Shogi coordinate for SFEN: row {r}, col {c}")
    return f"{9 - c}{chr(ord('a') + r)}"


def encode_move(move_tuple: MoveTuple) -> str:
    if len(move_tuple) == 5 and all(isinstance(v, int) for v in move_tuple[:4]) and isinstance(move_tuple[4], bool):
        fr, fc, tr, tc, promote = move_tuple
        return sq_to_sfen(fr, fc) + sq_to_sfen(tr, tc) + ("+" if promote else "")
    if (len(move_tuple) == 5 and move_tuple[0] is None and move_tuple[1] is None and isinstance(move_tuple[2], int)
            and isinstance(move_tuple[3], int) and isinstance(move_tuple[4], PieceType)):
        pt = move_tuple[4]
        if pt.value > 6:
            raise ValueError(f"PieceType {pt.name} is not a standard droppable piece for SFEN notation or is invalid.")
        return f"{'PLNSGBR'[pt.value]}*{sq_to_sfen(move_tuple[2], move_tuple[3])}"
    raise ValueError(f"Invalid MoveTuple format for SFEN conversion: {move_tuple}. "
                     f"Types: {[type(e).__name__ for e in move_tuple]}. Values: {[str(e) for e in move_tuple]}.")


_DROP_RE = re.compile(r"^([PLNSGBR])\*([1-9][a-i])$")
_MOVE_RE = re.compile(r"^([1-9][a-i])([1-9][a-i])(\+)?$")


def _parse_sq(s: str) -> Tuple[int, int]:
    return ord(s[1]) - ord("a"), 9 - int(s[0])


def sfen_to_move_tuple(sfen_move_str: str) -> MoveTuple:
    s = sfen_move_str.strip()
    m = _DROP_RE.match(s)
    if m:
        r, c = _parse_sq(m.group(2))
        return (None, None, r, c, PieceType("PLNSGBR".index(m.group(1))))
    m = _MOVE_RE.match(s)
    if m:
        fr, fc = _parse_sq(m.group(1))
        tr, tc = _parse_sq(m.group(2))
        return (fr, fc, tr, tc, m.group(3) is not None)
    raise ValueError(f"Invalid SFEN move format: {sfen_move_str}")


def game_to_text(game) -> str:
    lines = []
    for r, row in enumerate(game.board):
        cells = []
        for p in row:
            if p is None:
                cells.append(" . ")
            else:
                s = p.symbol()
                cells.append(f" {s} " if len(s) == 1 else f"{s} ")
        lines.append(f"{9 - r} " + "".join(cells))
    lines.append("a b c d e f g h i")
    lines.append(f"Turn: {game.current_player.name}, Move: {game.move_count + 1}")
    for color, name in ((0, "Black"), (1, "White")):
        lines.append(f"{name}'s hand: { {pt.name: c for pt, c in game.hands[color].items() if c > 0} }")
    return "\n".join(lines)


def pack_sfen(sfen_str: str):
    """SFEN -> (board codes int8[81], hands uint8[14], side, move_count) for VecShogiEnv.load_positions."""
    board_s, turn, hands_s, num = split_sfen(sfen_str)
    board = board_from_sfen_segment(board_s)
    hands = {0: {}, 1: {}}
    hands_from_sfen_segment(hands, hands_s)
    b = np.zeros(81, np.int8)
    for r in range(9):
        for c in range(9):
            p = board[r][c]
            if p is not None:
                b[r * 9 + c] = p.code
    h = np.zeros(14, np.uint8)
    for color in (0, 1):
        for pt, cnt in hands[color].items():
            h[color * 7 + pt.value] = cnt
    if int(num) < 1:
        raise ValueError("SFEN move number must be positive")
    return b, h, 0 if turn == "b" else 1, int(num) - 1


class HostPosition:
    """Host snapshot of one device game with the facade's attribute names (``board``, ``hands``, ``current_player``,
    ``move_count``, ``game_over``, ``winner``, ``termination_reason``, ``move_history``), enough for the text
    writers here and in kif.py.  Built by ``VecShogiEnv.to_games`` from kz_export_positions output."""

    _REASONS = {0: None, 1: "Tsumi", 2: "stalemate", 3: "Max moves reached", 4: "Sennichite"}

    def __init__(self, board_codes, hands14, side: int, move_count: int, status: int = 0, winner: int = -1):
        self.board = [[Piece.from_code(int(board_codes[r * 9 + c])) for c in range(9)] for r in range(9)]
        self.hands = {color: {PieceType(t): int(hands14[color * 7 + t]) for t in range(7)} for color in (0, 1)}
        self.current_player = Color(int(side))
        self.move_count = int(move_count)
        self.game_over = int(status) != 0
        self.termination_reason = self._REASONS.get(int(status))
        self.winner = Color(int(winner)) if int(winner) in (0, 1) else None
        self.move_history = []

    def to_sfen_string(self) -> str:
        return game_to_sfen(self)

    def to_string(self) -> str:
        return game_to_text(self)
