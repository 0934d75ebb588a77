"""drop-in for the reference's f2py module ``f_init`` (BROADCAST_npz.py:773): re-exports broadcast_b200.f_init"""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from broadcast_b200 import f_init as _m  # noqa: E402

globals().update(vars(_m))
__all__ = sorted(vars(_m))
