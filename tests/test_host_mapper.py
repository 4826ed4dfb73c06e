"""Host-side API mirror: definitions + PolicyOutputMapper (reference tests/shogi/test_shogi_utils.py)."""
import numpy as np
import pytest
import torch

from shogidrl_b200.shogi.definitions import Color, Piece, PieceType, TerminationReason, get_unpromoted_types
from shogidrl_b200.utils.policy_mapper import PolicyOutputMapper, index_to_move, move_to_index


@pytest.fixture(scope="module")
def mapper():
    return PolicyOutputMapper()


def test_total_actions(mapper):
    assert mapper.get_total_actions() == 13527 == len(mapper.idx_to_move) == len(mapper.move_to_idx)


def test_known_indices(mapper):
    # tests/shogi/test_shogi_utils.py:21-80 known answers
    assert mapper.shogi_move_to_policy_index((None, None, 4, 4, PieceType.PAWN)) == 13240
    assert mapper.shogi_move_to_policy_index((None, None, 0, 0, PieceType.LANCE)) == 12961
    a = mapper.shogi_move_to_policy_index((6, 6, 5, 6, False))
    assert mapper.shogi_move_to_policy_index((6, 6, 5, 6, True)) == a + 1


def test_enumeration_order_matches_reference_construction(mapper):
    # the reference appends in nested-loop order (utils.py:208-266); rebuild that order independently
    i = 0
    for fr in range(9):
        for fc in range(9):
            for tr in range(9):
                for tc in range(9):
                    if (fr, fc) == (tr, tc):
                        continue
                    assert mapper.idx_to_move[i] == (fr, fc, tr, tc, False)
                    assert mapper.idx_to_move[i + 1] == (fr, fc, tr, tc, True)
                    i += 2
    for tr in range(9):
        for tc in range(9):
            for pt in get_unpromoted_types():
                assert mapper.idx_to_move[i] == (None, None, tr, tc, pt)
                i += 1
    assert i == 13527


def test_round_trip_all(mapper):
    for i in range(0, 13527, 7):
        assert move_to_index(index_to_move(i)) == i
        assert mapper.shogi_move_to_policy_index(mapper.policy_index_to_shogi_move(i)) == i


def test_errors(mapper):
    with pytest.raises(IndexError):
        mapper.policy_index_to_shogi_move(13527)
    with pytest.raises(IndexError):
        mapper.policy_index_to_shogi_move(-1)
    with pytest.raises(ValueError):
        mapper.shogi_move_to_policy_index((0, 0, 0, 0, False))
    with pytest.raises(ValueError):
        mapper.get_legal_mask([(None, None, 0, 0, PieceType.KING)], torch.device("cpu"))


def test_legal_mask(mapper):
    moves = [(6, 0, 5, 0, False), (None, None, 4, 4, PieceType.GOLD)]
    m = mapper.get_legal_mask(moves, torch.device("cpu"))
    assert m.dtype == torch.bool and m.shape == (13527,) and int(m.sum()) == 2
    assert bool(m[mapper.shogi_move_to_policy_index(moves[0])]) and bool(m[mapper.shogi_move_to_policy_index(moves[1])])
    assert int(mapper.get_legal_mask([], torch.device("cpu")).sum()) == 0


def test_usi(mapper):
    assert mapper.shogi_move_to_usi((6, 2, 5, 2, False)) == "7g7f"
    assert mapper.shogi_move_to_usi((1, 7, 2, 6, True)) == "2b3c+"
    assert mapper.shogi_move_to_usi((None, None, 4, 4, PieceType.PAWN)) == "P*5e"
    for s in ["7g7f", "2b3c+", "P*5e", "R*1a"]:
        assert mapper.shogi_move_to_usi(mapper.usi_to_shogi_move(s)) == s
    with pytest.raises(ValueError):
        mapper.usi_to_shogi_move("K*5e")


def test_definitions():
    assert Color.BLACK.opponent() is Color.WHITE and PieceType.PROMOTED_ROOK.value == 13
    assert str(TerminationReason.CHECKMATE) == "Tsumi" and TerminationReason.REPETITION.value == "Sennichite"
    p = Piece(PieceType.PAWN, Color.WHITE)
    assert p.symbol() == "p" and p.code == 15 and Piece.from_code(15) == p and Piece.from_code(0) is None
    p.promote()
    assert p.type is PieceType.PROMOTED_PAWN and p.is_promoted and p.symbol() == "+p"
    with pytest.raises(TypeError):
        Piece(0, Color.BLACK)


def test_sfen_square_and_drop_letter_helpers():
    """The two private helpers of keisei/shogi/shogi_game_io.py (:744-776) that the reference's I/O tests import."""
    from shogidrl_b200.shogi.definitions import PieceType
    from shogidrl_b200.shogi.shogi_game_io import _get_piece_type_from_sfen_char, _parse_sfen_square, sfen_to_move_tuple
    assert _parse_sfen_square("9a") == (0, 0) and _parse_sfen_square("1i") == (8, 8) and _parse_sfen_square("7g") == (6, 2)
    for bad in ("", "7", "0a", "7j", "77", "7g7"):
        with pytest.raises(ValueError, match="Invalid SFEN square format"):
            _parse_sfen_square(bad)
    assert [_get_piece_type_from_sfen_char(c) for c in "PLNSGBR"] == [PieceType(i) for i in range(7)]
    for bad in ("K", "p", "+P", "X", ""):
        with pytest.raises(ValueError, match="Invalid SFEN piece character for drop"):
            _get_piece_type_from_sfen_char(bad)
    assert sfen_to_move_tuple("P*5e") == (None, None, 4, 4, PieceType.PAWN) and sfen_to_move_tuple("2b3a+") == (1, 7, 0, 6, True)
