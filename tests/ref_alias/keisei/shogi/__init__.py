import os

from shogidrl_b200.shogi import *  # noqa: F401,F403
from shogidrl_b200.shogi import Color, MoveTuple, Piece, PieceType, ShogiGame  # noqa: F401

_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))), "baseline", "_ref", "keisei", "shogi")
if os.path.isdir(_REF):
    __path__.append(_REF)  # e.g. keisei.shogi.features: not on the hot path, the reference's own
