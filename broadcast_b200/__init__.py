"""broadcast_b200 -- B200 (sm_100a) implementation of BROADCAST's finite-volume hot path.

Drop-in modules (same names and call-site signatures as the reference's f2py modules,
SURVEY.md section 8(b)):

    from broadcast_b200 import f_sch, f_lin, f_bnd, f_geom, f_norm, f_misc, f_dz, f_lindz, f_init

or, to run unmodified reference drivers, put ``broadcast_b200/dropin`` on ``sys.path`` so that
``import srcfv.f_sch`` / ``import misc.f_misc`` resolve to this package.
"""
from . import _lib
from ._lib import BroadcastB200Error
from .f2py_api import build as _build

_mods = _build(_lib.call_host)
f_sch = _mods["f_sch"]
f_lin = _mods["f_lin"]
f_bnd = _mods["f_bnd"]
f_geom = _mods["f_geom"]
f_norm = _mods["f_norm"]
f_misc = _mods["f_misc"]
f_dz = _mods["f_dz"]
f_lindz = _mods["f_lindz"]
f_init = _mods["f_init"]

__all__ = ["f_sch", "f_lin", "f_bnd", "f_geom", "f_norm", "f_misc", "f_dz", "f_lindz", "f_init", "BroadcastB200Error"]
