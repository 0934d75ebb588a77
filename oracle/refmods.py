"""f2py-shaped modules backed by ``oracle/_ref`` (the machine-translated reference Fortran).

TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.
Gives the translated reference the same call-site signatures as the product's drop-in modules
(broadcast_b200.f2py_api) so that one driver sequence can run on either.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import ref as _ref  # noqa: E402

sys.path.insert(0, os.path.dirname(_HERE))
from broadcast_b200.f2py_api import build as _build  # noqa: E402  (signature layer only, no compute)


def _fortran_interf(a):
    # product ABI passes {imin, jmin, imax, jmax}; Fortran integer(2,2) memory is imin, imax, jmin, jmax
    return np.array([a[0], a[2], a[1], a[3]], dtype=np.int32)


_WINDOW_ARGS = {
    "bc_wall_viscous_adia_2d": (3,), "bc_no_reflexion_2d": (3,), "bc_supandsubinlet_2d": (2,), "bc_extrapolate_o2_2d": (2,),
    "bc_wall_viscous_adia_2d_d": (4,), "bc_no_reflexion_2d_d": (4,), "bc_supandsubinlet_2d_d": (3,),
    "bc_extrapolate_o2_2d_d": (3,), "jn_match_2d": (1, 9), "jn_match_geom_2d": (1, 9),
    "bc_wall_viscous_iso_2d": (5,), "bc_wall_viscous_iso_2d_d": (6,), "bc_symmetry_2d": (2,), "bc_symmetry_2d_d": (3,),
    "bc_antisymmetry_2d": (2,), "bc_antisymmetry_2d_d": (3,), "bc_pressure_2d": (2,), "bc_pressure_2d_d": (3,),
    "bc_wall_blow_profile_2d": (4,), "bc_wall_blow_profile_2d_d": (7,), "bc_wall_viscous_iso_profile_2d": (5,),
    "bc_wall_viscous_iso_profile_2d_d": (9,), "bc_general_2d": (2,), "bc_general_2d_d": (3,),
}


def make(fast: bool = False):
    lib = _ref.load(fast)

    def backend(name, *args):
        args = list(args)
        for pos in _WINDOW_ARGS.get(name, ()):
            args[pos] = _fortran_interf(args[pos])
        getattr(lib, name)(*args)

    mods = _build(backend)
    mods["_lib"] = lib
    return mods


def available(fast: bool = False) -> bool:
    return _ref.available(fast)
