"""KIF-style game export (keisei/shogi/shogi_game_io.py:588-738, ``game_to_kif``).

Host-side text only: works on any object with the facade's attributes (``hands``, ``current_player``,
``move_history``, ``game_over``, ``termination_reason``, ``winner``), so device batches can be dumped through
``VecShogiEnv.to_games``.  The reference's quirks are kept because its files are what downstream tools read:
the board block is always the hirate position, the hand lines show the *current* hands, the side marker is the
side to move *now*, drops are skipped in the move list, and move numbers are history indices.
"""
from __future__ import annotations

import datetime
from typing import Optional

from .definitions import KIF_PIECE_SYMBOLS, Color, PieceType, TerminationReason

_HEADER = (
    "#KIF version=2.0 encoding=UTF-8",
    "*Event: Casual Game",
    "*Site: Local Machine",
)
_BACK_RANK = ("KY", "KE", "GI", "KI", "OU", "KI", "GI", "KE", "KY")
_HAND_ORDER = (PieceType.ROOK, PieceType.BISHOP, PieceType.GOLD, PieceType.SILVER, PieceType.KNIGHT, PieceType.LANCE,
               PieceType.PAWN)
_REASON_TEXT = {"Tsumi": "詰み", "Toryo": "投了", "Sennichite": "千日手", "Stalemate": "持将棋", "Max moves reached": "持将棋"}
_DRAW_REASONS = {TerminationReason.REPETITION.value, TerminationReason.IMPASSE.value,
                 TerminationReason.MAX_MOVES_EXCEEDED.value}


def _hirate_block():
    """The fixed board block the reference prints (single-blank cell pattern on the sparse ranks)."""
    empty = " *" * 9 + " "
    rows = ["".join("-" + s for s in _BACK_RANK), " * -HI * * * * * -KA * ", "-FU" * 9, empty, empty, empty, "+FU" * 9,
            " * +KA * * * * * +HI * ", "".join("+" + s for s in _BACK_RANK)]
    return [f"P{i + 1}{r}" for i, r in enumerate(rows)]


def _hand_line(prefix: str, hand) -> str:
    return prefix + "".join(f"{hand.get(pt, 0):02d}{KIF_PIECE_SYMBOLS.get(pt, '??')}" for pt in _HAND_ORDER)


def game_to_kif(game, filename: Optional[str] = None, sente_player_name: str = "Sente",
                gote_player_name: str = "Gote") -> Optional[str]:
    """KIF text of ``game``; written to ``filename`` (returns None) or returned as a string."""
    out = list(_HEADER)
    out.append(f"*Date: {datetime.date.today().strftime('%Y/%m/%d')}")
    out.append(f"*Player Sente: {sente_player_name}")
    out.append(f"*Player Gote: {gote_player_name}")
    out.append("*Handicap: HIRATE")
    out.extend(_hirate_block())
    out.append(_hand_line("P+", game.hands[Color.BLACK.value]))
    out.append(_hand_line("P-", game.hands[Color.WHITE.value]))
    out.append("+" if game.current_player == Color.BLACK else "-")
    out.append("moves")
    for i, rec in enumerate(game.move_history):
        mv = rec.get("move")
        if not mv or any(mv[k] is None for k in range(4)):
            continue  # drops have no from-square and are left out, like the reference
        out.append(f"{i + 1} {mv[0] + 1}{chr(mv[1] + ord('a'))}{mv[2] + 1}{chr(mv[3] + ord('a'))}" + ("+" if mv[4] else ""))
    if game.game_over:
        reason = game.termination_reason
        shown = "" if reason is None else _REASON_TEXT.get(reason, reason)
        if shown:
            out.append(shown)
        if game.winner == Color.BLACK:
            out.append("RESULT:SENTE_WIN")
        elif game.winner == Color.WHITE:
            out.append("RESULT:GOTE_WIN")
        elif game.winner is None and reason in _DRAW_REASONS:
            out.append("RESULT:DRAW")
    out.append("*EOF")
    text = "\n".join(out)
    if filename:
        with open(filename, "w", encoding="utf-8") as f:
            f.write(text)
        return None
    return text
