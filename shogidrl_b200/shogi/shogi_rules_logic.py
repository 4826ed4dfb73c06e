"""Function surface of keisei/shogi/shogi_rules_logic.py over the device-backed ``ShogiGame``.

The reference keeps its rules as module-level functions taking the game (generate_all_legal_moves :486-635,
check_for_uchi_fu_zume :275-359, can_drop_specific_piece :424-483, ...) and its tests and evaluation code import them
directly.  Here the rules run in the CUDA engine behind the game object, so every function is a thin delegation to the
facade; only the two stateless promotion predicates are computed on the host."""
from __future__ import annotations

from typing import List, Optional, Tuple

from .definitions import Color, MoveTuple, Piece, PieceType

_PROMOTABLE = (PieceType.PAWN, PieceType.LANCE, PieceType.KNIGHT, PieceType.SILVER, PieceType.BISHOP, PieceType.ROOK)
_SLIDERS = (PieceType.LANCE, PieceType.BISHOP, PieceType.ROOK, PieceType.PROMOTED_BISHOP, PieceType.PROMOTED_ROOK)


def find_king(game, color: Color) -> Optional[Tuple[int, int]]:
    return game.find_king(color)


def is_in_check(game, player_color: Color, debug_recursion: bool = False) -> bool:
    """A missing king counts as in check (shogi_rules_logic.py:42-55)."""
    return game.is_in_check(player_color)


def is_piece_type_sliding(piece_type: PieceType) -> bool:
    return piece_type in _SLIDERS


def generate_piece_potential_moves(game, piece: Piece, r_from: int, c_from: int) -> List[Tuple[int, int]]:
    """Pseudo-legal targets (:82-208) from the engine's kz_piece_targets; the order of the list is not contractual."""
    return game.get_individual_piece_moves(piece, r_from, c_from)


def check_for_nifu(game, color: Color, col: int) -> bool:
    return game.is_nifu(color, col)


def check_if_square_is_attacked(game, r_target: int, c_target: int, attacker_color: Color, debug: bool = False) -> bool:
    """Is (r, c) a pseudo-legal target of any piece of ``attacker_color`` (:234-272)."""
    for r in range(9):
        for c in range(9):
            p = game.get_piece(r, c)
            if p is not None and p.color == attacker_color and (r_target, c_target) in game.get_individual_piece_moves(p, r, c):
                return True
    return False


def check_for_uchi_fu_zume(game, drop_row: int, drop_col: int, color: Color) -> bool:
    return game.is_uchi_fu_zume(drop_row, drop_col, color)


def can_promote_specific_piece(game, piece: Piece, r_from: int, r_to: int) -> bool:
    """A promotable, unpromoted piece whose move starts or ends in its promotion zone (:382-401)."""
    if piece.type not in _PROMOTABLE:
        return False
    return game.is_in_promotion_zone(r_from, piece.color) or game.is_in_promotion_zone(r_to, piece.color)


def must_promote_specific_piece(piece: Piece, r_to: int) -> bool:
    """Pawn / lance on the last rank, knight on the last two (:404-421)."""
    last, second = (0, 1) if piece.color == Color.BLACK else (8, 7)
    if piece.type in (PieceType.PAWN, PieceType.LANCE):
        return r_to == last
    if piece.type == PieceType.KNIGHT:
        return r_to in (last, second)
    return False


def can_drop_specific_piece(game, piece_type: PieceType, r_to: int, c_to: int, color: Color,
                            is_escape_check: bool = False) -> bool:
    """Square empty; pawn: no nifu, not the last rank, no uchifuzume (skipped for ``is_escape_check``); lance: not the
    last rank; knight: not the last two (:424-483).  The hand count is the caller's business, as in the reference."""
    if game.get_piece(r_to, c_to) is not None:
        return False
    last, second = (0, 1) if color == Color.BLACK else (8, 7)
    if piece_type == PieceType.PAWN:
        if game.is_nifu(color, c_to) or r_to == last:
            return False
        return is_escape_check or not game.is_uchi_fu_zume(r_to, c_to, color)
    if piece_type == PieceType.LANCE:
        return r_to != last
    if piece_type == PieceType.KNIGHT:
        return r_to not in (last, second)
    return True


def generate_all_legal_moves(game, is_uchi_fu_zume_check: bool = False) -> List[MoveTuple]:
    """Legal moves of the side to move (:486-635).  ``is_uchi_fu_zume_check`` is the reference's recursion guard (pawn
    drops are then not tested for uchifuzume): the engine's refresh has no such mode on the facade, and none of the
    reference's callers outside its own recursion pass it."""
    return game.get_legal_moves()


def check_for_sennichite(game) -> bool:
    return game.is_sennichite()
