"""shogidrl_b200 -- B200-native (sm_100a) implementation of Keisei's self-play rollout hot path.

Device-resident Shogi games (one warp per game), legal-move generation, make_move, 46x9x9 observation,
13,527-action legal mask, masked sampling and GAE as hand-written CUDA kernels behind a C ABI
(include/keisei_b200.h), with a Python host side that keeps the reference's API names."""
from ._native import NUM_ACTIONS, OBS_FLOATS, MASK_PAD_STRIDE, NativeError  # noqa: F401
from .vec_env import VecShogiEnv  # noqa: F401

__all__ = ["VecShogiEnv", "NUM_ACTIONS", "OBS_FLOATS", "MASK_PAD_STRIDE", "NativeError"]
