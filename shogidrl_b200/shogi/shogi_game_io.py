"""Names of keisei/shogi/shogi_game_io.py for callers that import the I/O helpers directly."""
from .kif import game_to_kif  # noqa: F401
from .sfen import board_from_sfen_segment as populate_board_from_sfen_segment  # noqa: F401
from .sfen import encode_move as encode_move_to_sfen_string  # noqa: F401
from .sfen import game_to_sfen as convert_game_to_sfen_string  # noqa: F401
from .sfen import game_to_text as convert_game_to_text_representation  # noqa: F401
from .sfen import hands_from_sfen_segment as populate_hands_from_sfen_segment  # noqa: F401
from .sfen import sfen_to_move_tuple, split_sfen as parse_sfen_string_components  # noqa: F401


def generate_neural_network_observation(game):
    """46 x 9 x 9 observation of ``game`` (shogi_game_io.py:434-539), computed by the device engine."""
    return game.get_observation()


def _parse_sfen_square(sfen_sq: str):
    """"7g" -> (row, col), 0-based from the top-left corner; anything else is a ValueError (shogi_game_io.py:744-761)."""
    if len(sfen_sq) != 2 or sfen_sq[0] not in "123456789" or sfen_sq[1] not in "abcdefghi":
        raise ValueError(f"Invalid SFEN square format: {sfen_sq}")
    return "abcdefghi".index(sfen_sq[1]), 9 - int(sfen_sq[0])


def _get_piece_type_from_sfen_char(char: str):
    """Piece letter of a drop ("P*5e") -> PieceType; kings and promoted pieces cannot be dropped (:764-776)."""
    from .definitions import PieceType
    if len(char) == 1 and char in "PLNSGBR":
        return PieceType("PLNSGBR".index(char))
    raise ValueError(f"Invalid SFEN piece character for drop: {char}")
