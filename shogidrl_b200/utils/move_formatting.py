"""Move descriptions for demo-mode logging, with the strings of keisei/utils/move_formatting.py (the step manager's
move log and the TUI show them): ``"7g7f - Fuhyō (Pawn) moving from 7g to 7f."``, ``"P*5e - Fuhyō (Pawn) drop to 5e."``."""
from __future__ import annotations

from ..shogi.definitions import PieceType

_NAMES = ["Fuhyō (Pawn)", "Kyōsha (Lance)", "Keima (Knight)", "Ginsho (Silver General)", "Kinshō (Gold General)",
          "Kakugyō (Bishop)", "Hisha (Rook)", "Ōshō (King)", "Tokin (Promoted Pawn)", "Narikyo (Promoted Lance)",
          "Narikei (Promoted Knight)", "Narigin (Promoted Silver)", "Ryūma (Dragon Horse)", "Ryūō (Dragon King)"]
_PROMOTES_TO = {0: 8, 1: 9, 2: 10, 3: 11, 5: 12, 6: 13}


def _get_piece_name(piece_type, is_promoting: bool = False) -> str:
    v = getattr(piece_type, "value", None)
    if not isinstance(piece_type, PieceType) or v is None or not 0 <= v < len(_NAMES):
        return str(piece_type)
    if is_promoting and v in _PROMOTES_TO:
        return f"{_NAMES[v]} → {_NAMES[_PROMOTES_TO[v]]}"
    return _NAMES[v]


def _coords_to_square_name(row: int, col: int) -> str:
    return f"{9 - col}{chr(ord('a') + row)}"


def _describe(move, mapper, piece_type, known: bool) -> str:
    if move is None:
        return "None"
    try:
        usi = mapper.shogi_move_to_usi(move)
        if len(move) == 5 and move[0] is None:
            what = f"{_get_piece_name(move[4], False)} drop to {_coords_to_square_name(move[2], move[3])}"
        else:
            fr, fc, tr, tc, promote = move
            if known:
                try:
                    name = _get_piece_name(piece_type, promote)
                except (AttributeError, KeyError, TypeError):
                    name = "piece"
            else:
                name = "piece promoting" if (promote and piece_type is _NO_GAME) else "piece"
            what = f"{name} moving from {_coords_to_square_name(fr, fc)} to {_coords_to_square_name(tr, tc)}"
        return f"{usi} - {what}."
    except Exception as e:  # the reference falls back to the raw tuple on any formatting problem
        return f"{str(move)} (format error: {e})"


_NO_GAME = object()


def format_move_with_description(selected_shogi_move, policy_output_mapper, game=None) -> str:
    """USI + description, the piece looked up on ``game`` (before the move is made) when one is given."""
    if selected_shogi_move is None:
        return "None"
    if game is None:
        return _describe(selected_shogi_move, policy_output_mapper, _NO_GAME, False)
    piece = None
    if selected_shogi_move[0] is not None:
        try:
            piece = game.get_piece(selected_shogi_move[0], selected_shogi_move[1])
        except (AttributeError, KeyError, TypeError):
            piece = None
    return _describe(selected_shogi_move, policy_output_mapper, getattr(piece, "type", None), piece is not None)


def format_move_with_description_enhanced(selected_shogi_move, policy_output_mapper, piece_info=None) -> str:
    """The same with the moving piece handed in (``game.get_piece`` taken before the move)."""
    if piece_info is None:
        return _describe(selected_shogi_move, policy_output_mapper, None, False)
    try:
        ptype = piece_info.type
    except AttributeError:
        return _describe(selected_shogi_move, policy_output_mapper, None, False)
    return _describe(selected_shogi_move, policy_output_mapper, ptype, True)
