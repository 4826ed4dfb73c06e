from shogidrl_b200.shogi import *  # noqa: F401,F403
from shogidrl_b200.shogi import Color, MoveTuple, Piece, PieceType, ShogiGame  # noqa: F401
