"""Re-exports, like keisei/shogi/shogi_engine.py:1-11."""
from .definitions import Color, MoveTuple, Piece, PieceType  # noqa: F401
from .shogi_game import ShogiGame  # noqa: F401

__all__ = ["Color", "PieceType", "Piece", "MoveTuple", "ShogiGame"]
