"""Alias of the reference's top-level package for running its own test files against shogidrl_b200 (tests only).

The hot-path modules (keisei.shogi.*, keisei.utils.PolicyOutputMapper, keisei.core.experience_buffer / ppo_agent,
keisei.training.step_manager / env_manager / parallel) resolve to this repository's classes; every other module
(config_schema, constants, loggers, ...) falls through to the UNMODIFIED reference installed under baseline/_ref -- the
situation of a maintainer who swaps the hot path into the existing package."""
import os

_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))), "baseline", "_ref", "keisei")
if os.path.isdir(_REF):
    __path__.append(_REF)
