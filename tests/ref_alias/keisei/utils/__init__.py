import os

_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))), "baseline", "_ref", "keisei", "utils")
if os.path.isdir(_REF):
    __path__.append(_REF)
    from .move_formatting import (_coords_to_square_name, _get_piece_name, format_move_with_description,  # noqa: F401
                                  format_move_with_description_enhanced)
    from .utils import BaseOpponent, EvaluationLogger, TrainingLogger, load_config  # noqa: F401  (the reference's own)

from shogidrl_b200.utils import *  # noqa: F401,F403,E402
from shogidrl_b200.utils import PolicyOutputMapper  # noqa: F401,E402  (the hot-path class: this repository's)
from shogidrl_b200.utils.move_formatting import (_coords_to_square_name, _get_piece_name,  # noqa: F401,E402  (this repository's)
                                                 format_move_with_description, format_move_with_description_enhanced)
