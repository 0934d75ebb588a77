"""drop-in for the reference's f2py module ``srcfv.f_lindz`` (BROADCAST_npz.py:15-34): re-exports broadcast_b200.f_lindz"""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))))
from broadcast_b200 import f_lindz as _m  # noqa: E402

globals().update(vars(_m))
__all__ = sorted(vars(_m))
