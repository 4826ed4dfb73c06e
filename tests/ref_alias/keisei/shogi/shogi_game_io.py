from shogidrl_b200.shogi.shogi_game_io import *  # noqa: F401,F403
from shogidrl_b200.shogi import shogi_game_io as _io

globals().update({k: v for k, v in vars(_io).items() if not k.startswith("__")})
