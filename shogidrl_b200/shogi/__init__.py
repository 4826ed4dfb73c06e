"""Shogi engine API (names of keisei.shogi); rules run on the CUDA engine."""
from .definitions import Color, MoveTuple, Piece, PieceType, TerminationReason, get_unpromoted_types  # noqa: F401
from .shogi_game import ShogiGame  # noqa: F401

__all__ = ["Color", "MoveTuple", "Piece", "PieceType", "TerminationReason", "ShogiGame", "get_unpromoted_types"]
