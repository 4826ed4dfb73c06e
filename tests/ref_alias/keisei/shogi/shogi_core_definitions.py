from shogidrl_b200.shogi.definitions import *  # noqa: F401,F403
from shogidrl_b200.shogi import definitions as _d

globals().update({k: v for k, v in vars(_d).items() if not k.startswith("__")})
