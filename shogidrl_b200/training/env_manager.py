"""EnvManager -- API of keisei/training/env_manager.py (setup :46-83, action-space validation :85-105, reset /
initialise / validate / count / seed helpers :121-261), building the device-backed ShogiGame and, for the
batched fast path, a VecShogiEnv."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

from ..shogi import ShogiGame
from ..utils import PolicyOutputMapper


class EnvManager:
    def __init__(self, config, logger_func: Optional[Callable] = None):
        self.config = config
        self.logger_func = logger_func or (lambda msg: None)
        self.game: Optional[ShogiGame] = None
        self.policy_output_mapper: Optional[PolicyOutputMapper] = None
        self.action_space_size: int = 0
        self.obs_space_shape: Optional[Tuple[int, int, int]] = None
        self.vec_env = None

    def setup_environment(self) -> Tuple[ShogiGame, PolicyOutputMapper]:
        try:
            self.game = ShogiGame(max_moves_per_game=self.config.env.max_moves_per_game,
                                  device=getattr(self.config.env, "device", "cuda"))
            seed = getattr(self.config.env, "seed", None)
            if seed is not None:
                try:
                    self.game.seed(seed)
                    self.logger_func(f"Environment seeded with: {seed}")
                except Exception as e:
                    self.logger_func(f"Warning: Failed to seed environment: {e}")
            self.obs_space_shape = (self.config.env.input_channels, 9, 9)
        except (RuntimeError, ValueError, OSError) as e:
            self.logger_func(f"Error initializing ShogiGame: {e}. Aborting.")
            raise RuntimeError(f"Failed to initialize ShogiGame: {e}") from e
        try:
            self.policy_output_mapper = PolicyOutputMapper()
            self.action_space_size = self.policy_output_mapper.get_total_actions()
            self._validate_action_space()
        except (RuntimeError, ValueError) as e:
            self.logger_func(f"Error initializing PolicyOutputMapper: {e}")
            raise RuntimeError(f"Failed to initialize PolicyOutputMapper: {e}") from e
        return self.game, self.policy_output_mapper

    def setup_vector_environment(self, num_envs: int, env_offset: int = 0, auto_reset: bool = True):
        """Batched environment for the vectorised rollout path: ``num_envs`` device-resident games."""
        from ..vec_env import VecShogiEnv
        seed = getattr(self.config.env, "seed", None)
        self.vec_env = VecShogiEnv(num_envs, max_moves_per_game=self.config.env.max_moves_per_game,
                                   device=getattr(self.config.env, "device", "cuda"),
                                   seed=0 if seed is None else int(seed), env_offset=env_offset, auto_reset=auto_reset)
        return self.vec_env

    def _validate_action_space(self):
        if self.policy_output_mapper is None:
            self.logger_func("CRITICAL: PolicyOutputMapper not initialized before _validate_action_space.")
            raise ValueError("PolicyOutputMapper not initialized.")
        cfg_n = self.config.env.num_actions_total
        got = self.policy_output_mapper.get_total_actions()
        if cfg_n != got:
            msg = (f"Action space mismatch: config specifies {cfg_n} actions but PolicyOutputMapper provides "
                   f"{got} actions")
            self.logger_func(f"CRITICAL: {msg}")
            raise ValueError(msg)
        self.logger_func(f"Action space validated: {got} total actions")

    def get_environment_info(self) -> dict:
        return {"game": self.game, "policy_mapper": self.policy_output_mapper,
                "action_space_size": self.action_space_size, "obs_space_shape": self.obs_space_shape,
                "input_channels": self.config.env.input_channels,
                "num_actions_total": self.config.env.num_actions_total, "seed": getattr(self.config.env, "seed", None),
                "game_type": type(self.game).__name__, "policy_mapper_type": type(self.policy_output_mapper).__name__}

    def _refuse(self, message: str, result=False):
        """Log ``message`` and hand back the failure value: every helper below reports problems this way, never by raising."""
        self.logger_func(message)
        return result

    def _guarded(self, what: str, fn, on_error):
        try:
            return fn()
        except Exception as e:
            return self._refuse(f"{what}: {e}", on_error)

    def reset_game(self) -> bool:
        if not self.game:
            return self._refuse("Error: Game not initialized. Cannot reset.")
        return self._guarded("Error resetting game", lambda: (self.game.reset(), True)[1], False)

    def initialize_game_state(self) -> Optional[np.ndarray]:
        if not self.game:
            return self._refuse("Error: Game not initialized. Cannot get initial observation.", None)
        return self._guarded("Error initializing game state", self.game.reset, None)

    def validate_environment(self) -> bool:
        """Game and mapper present, action space non-empty, reset works, observation shape is (C, 9, 9)."""
        def checks() -> bool:
            for missing, what in ((self.game is None, "game not initialized"),
                                  (self.policy_output_mapper is None, "policy mapper not initialized"),
                                  (self.action_space_size <= 0, "invalid action space size")):
                if missing:
                    return self._refuse(f"Environment validation failed: {what}")
            before = self.game.get_observation()
            if not self.reset_game():
                return self._refuse("Environment validation failed: game reset failed")
            if not np.array_equal(before, self.game.get_observation()):
                self.logger_func("Environment validation warning: Observation after reset differs from initial "
                                 "observation. This might be expected if seeding is not deterministic or initial "
                                 "state has randomness.")
            if self.obs_space_shape is None or len(self.obs_space_shape) != 3:
                return self._refuse("Environment validation failed: invalid observation space shape")
            self.logger_func("Environment validation passed")
            return True
        return self._guarded("Environment validation failed with exception", checks, False)

    def get_legal_moves_count(self) -> int:
        return self._guarded("Error getting legal moves count",
                             lambda: len(self.game.get_legal_moves()) if self.game else 0, 0)

    def setup_seeding(self, seed: Optional[int] = None):
        seed_value = getattr(self.config.env, "seed", None) if seed is None else seed
        if not self.game:
            return self._refuse("Error: Game not initialized. Cannot set seed.")
        if seed_value is None:
            return self._refuse("No seed value provided for re-seeding.")
        if not hasattr(self.game, "seed"):
            return self._refuse(f"Warning: Game object does not have a 'seed' method. Cannot re-seed with {seed_value}.")

        def reseed() -> bool:
            self.game.seed(seed_value)
            self.logger_func(f"Environment re-seeded with: {seed_value}")
            return True
        return self._guarded("Error setting environment seed", reseed, False)
