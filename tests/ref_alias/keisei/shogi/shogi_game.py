from shogidrl_b200.shogi.shogi_game import ShogiGame  # noqa: F401
