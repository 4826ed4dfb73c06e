from shogidrl_b200.core.base_actor_critic import ActorCritic  # noqa: F401
