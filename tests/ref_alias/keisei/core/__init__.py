import os

_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))), "baseline", "_ref", "keisei", "core")
if os.path.isdir(_REF):
    __path__.append(_REF)

from shogidrl_b200.core import ActorCritic, BaseActorCriticModel  # noqa: F401,E402
