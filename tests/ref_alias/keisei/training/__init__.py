import os

_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))), "baseline", "_ref", "keisei", "training")
if os.path.isdir(_REF):
    __path__.append(_REF)
