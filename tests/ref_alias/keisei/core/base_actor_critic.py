from shogidrl_b200.core.base_actor_critic import BaseActorCriticModel  # noqa: F401
