import sys

import shogidrl_b200.core.ppo_agent as _m

sys.modules[__name__] = _m  # the very module: unittest.mock.patch on this name patches the class under test
