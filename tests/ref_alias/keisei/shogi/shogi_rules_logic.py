from shogidrl_b200.shogi.shogi_rules_logic import *  # noqa: F401,F403
