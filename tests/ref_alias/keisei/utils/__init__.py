from shogidrl_b200.utils import *  # noqa: F401,F403
from shogidrl_b200.utils import PolicyOutputMapper  # noqa: F401
