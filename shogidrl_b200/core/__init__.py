from .base_actor_critic import ActorCritic, ActorCriticResTower, BaseActorCriticModel, model_factory  # noqa: F401
from .experience_buffer import Experience, ExperienceBuffer, RolloutBuffer  # noqa: F401
from .ppo_agent import PPOAgent  # noqa: F401
