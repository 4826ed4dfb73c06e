"""Types and constants of the Shogi engine API, value-compatible with the reference's
keisei/shogi/shogi_core_definitions.py (Color :50-61, PieceType :64-83, TerminationReason :135-150,
MoveTuple :153-168, observation plane constants :272-283, Piece :393-509).

The device engine does not use these objects; they exist so that code written against
``keisei.shogi`` (tests, the TUI, evaluation loops) sees the same names, values and behaviour."""
from __future__ import annotations

from enum import Enum
from typing import Dict, List, Optional, Tuple, Union


class Color(Enum):
    BLACK = 0  # sente, moves toward row 0
    WHITE = 1  # gote, moves toward row 8

    def opponent(self) -> "Color":
        return Color(1 - self.value)


class PieceType(Enum):
    PAWN = 0
    LANCE = 1
    KNIGHT = 2
    SILVER = 3
    GOLD = 4
    BISHOP = 5
    ROOK = 6
    KING = 7
    PROMOTED_PAWN = 8
    PROMOTED_LANCE = 9
    PROMOTED_KNIGHT = 10
    PROMOTED_SILVER = 11
    PROMOTED_BISHOP = 12
    PROMOTED_ROOK = 13

    def to_usi_char(self) -> str:
        if self.value > 6:
            raise ValueError(f"Piece type {self.name} cannot be dropped or has no standard single USI drop character.")
        return "PLNSGBR"[self.value]


class TerminationReason(Enum):
    CHECKMATE = "Tsumi"
    STALEMATE = "stalemate"
    REPETITION = "Sennichite"
    MAX_MOVES_EXCEEDED = "Max moves reached"
    RESIGNATION = "resignation"
    TIME_FORFEIT = "time_forfeit"
    ILLEGAL_MOVE = "illegal_move"
    AGREEMENT = "agreement"
    IMPASSE = "impasse"
    NO_CONTEST = "no_contest"

    def __str__(self) -> str:
        return self.value


BoardMoveTuple = Tuple[int, int, int, int, bool]
DropMoveTuple = Tuple[Optional[int], Optional[int], int, int, PieceType]
MoveTuple = Union[BoardMoveTuple, DropMoveTuple]

_HAND_TYPES = [PieceType(i) for i in range(7)]
_PROMOTES_TO = {0: 8, 1: 9, 2: 10, 3: 11, 5: 12, 6: 13}

BASE_TO_PROMOTED_TYPE: Dict[PieceType, PieceType] = {PieceType(b): PieceType(p) for b, p in _PROMOTES_TO.items()}
PROMOTED_TO_BASE_TYPE: Dict[PieceType, PieceType] = {p: b for b, p in BASE_TO_PROMOTED_TYPE.items()}
PROMOTED_TYPES_SET = set(PROMOTED_TO_BASE_TYPE)
PIECE_TYPE_TO_HAND_TYPE: Dict[PieceType, PieceType] = {**{t: t for t in _HAND_TYPES}, **PROMOTED_TO_BASE_TYPE}

OBS_CURR_PLAYER_UNPROMOTED_START = 0
OBS_CURR_PLAYER_PROMOTED_START = 8
OBS_OPP_PLAYER_UNPROMOTED_START = 14
OBS_OPP_PLAYER_PROMOTED_START = 22
OBS_CURR_PLAYER_HAND_START = 28
OBS_OPP_PLAYER_HAND_START = 35
OBS_CURR_PLAYER_INDICATOR = 42
OBS_MOVE_COUNT = 43
OBS_RESERVED_1 = 44
OBS_RESERVED_2 = 45
OBS_UNPROMOTED_ORDER: List[PieceType] = [PieceType(i) for i in range(8)]
OBS_PROMOTED_ORDER: List[PieceType] = [PieceType(i) for i in range(8, 14)]

_SYMBOLS = {**{PieceType(i): "PLNSGBRK"[i] for i in range(8)},
            **{PieceType(p): "+" + "PLNSGBR"[b] for b, p in _PROMOTES_TO.items()}}
SYMBOL_TO_PIECE_TYPE: Dict[str, PieceType] = {s: t for t, s in _SYMBOLS.items()}

KIF_PIECE_SYMBOLS: Dict[PieceType, str] = dict(zip(
    [PieceType(i) for i in range(14)],
    ["FU", "KY", "KE", "GI", "KI", "KA", "HI", "OU", "TO", "NY", "NK", "NG", "UM", "RY"]))


def get_unpromoted_types() -> List[PieceType]:
    """Piece types that can be held in hand, in hand-plane order (P, L, N, S, G, B, R)."""
    return list(_HAND_TYPES)


def get_piece_type_from_symbol(symbol: str) -> PieceType:
    key = symbol if symbol in SYMBOL_TO_PIECE_TYPE else symbol.upper()
    if key in SYMBOL_TO_PIECE_TYPE and (symbol == key or symbol.lstrip("+").islower()):
        return SYMBOL_TO_PIECE_TYPE[key]
    raise ValueError(f"Unknown piece symbol: {symbol}")


class Piece:
    """A piece = (type, colour); ``is_promoted`` is derived from the type."""

    __slots__ = ("type", "color", "is_promoted")

    def __init__(self, piece_type: PieceType, color: Color):
        if not isinstance(piece_type, PieceType):
            raise TypeError("piece_type must be an instance of PieceType")
        if not isinstance(color, Color):
            raise TypeError("color must be an instance of Color")
        self.type = piece_type
        self.color = color
        self.is_promoted = piece_type in PROMOTED_TYPES_SET

    # engine piece code (0 = empty): 1 + type + 14 * colour
    @property
    def code(self) -> int:
        return 1 + self.type.value + 14 * self.color.value

    @staticmethod
    def from_code(code: int) -> Optional["Piece"]:
        if code == 0:
            return None
        return Piece(PieceType((code - 1) % 14), Color((code - 1) // 14))

    def symbol(self) -> str:
        s = _SYMBOLS[self.type]
        return s.lower() if self.color is Color.WHITE else s

    def promote(self) -> None:
        if self.type in BASE_TO_PROMOTED_TYPE:
            self.type = BASE_TO_PROMOTED_TYPE[self.type]
            self.is_promoted = True

    def unpromote(self) -> None:
        if self.type in PROMOTED_TO_BASE_TYPE:
            self.type = PROMOTED_TO_BASE_TYPE[self.type]
            self.is_promoted = False

    def __repr__(self) -> str:
        return f"Piece({self.type.name}, {self.color.name})"

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, Piece):
            return NotImplemented
        return self.type == other.type and self.color == other.color

    def __hash__(self) -> int:
        return hash((self.type, self.color))

    def __deepcopy__(self, memo) -> "Piece":
        return Piece(self.type, self.color)
