from .env_manager import EnvManager  # noqa: F401
from .step_manager import EpisodeState, StepManager, StepResult, VecStepManager  # noqa: F401
