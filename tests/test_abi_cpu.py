"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/broadcast_b200.h declares; calls fail loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "broadcast_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bcd?_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from broadcast_b200 import _lib
    lib = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import broadcast_b200 as bb
    wd = np.zeros((16, 16, 5), order="F")
    with pytest.raises(bb.BroadcastB200Error):
        bb.f_misc.testvector(wd, 0, 0, 0, 3, 10, 10)


def test_signature_layer_validates_like_f2py():
    import broadcast_b200 as bb
    w = np.zeros((16, 16, 5))  # C-ordered: f2py rejects it for intent(inout)
    with pytest.raises(ValueError):
        bb.f_bnd.bc_wall_viscous_adia_2d(w, "Jlo", 1.4, np.array([[1, 1], [10, 1]]), 3, 10, 10)
    w = np.zeros((16, 16, 5), order="F")
    with pytest.raises(ValueError):
        bb.f_bnd.bc_wall_viscous_adia_2d(w, "Klo", 1.4, np.array([[1, 1], [10, 1]]), 3, 10, 10)


def test_dropin_modules_resolve_the_reference_imports():
    """`import srcfv.f_sch as f_sch` etc. (BROADCAST_npz.py:15-34, BROADCAST_npz_sens.py:37) with broadcast_b200/dropin on sys.path:
    every routine name the drivers call is there (names only: no compute without a GPU)"""
    import importlib
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys
sys.path.insert(0, sys.argv[1])
import srcfv.f_sch as f_sch, srcfv.f_lin as f_lin, srcfv.f_bnd as f_bnd, srcfv.f_geom as f_geom, srcfv.f_norm as f_norm
import srcfv.f_dz as f_dz, srcfv.f_lindz as f_lindz, misc.f_misc as f_misc, f_init
need = {f_sch: ["flux_num_dnc5_2d", "flux_num_dnc5_nowall_2d"], f_lin: ["flux_num_dnc5_2d_d", "bc_wall_viscous_adia_2d_d",
        "bc_no_reflexion_2d_d", "bc_supandsubinlet_2d_d", "bc_extrapolate_o2_2d_d"],
        f_bnd: ["bc_wall_viscous_adia_2d", "bc_no_reflexion_2d", "bc_supandsubinlet_2d", "bc_extrapolate_o2_2d", "jn_match_2d",
                "jn_match_geom_2d"], f_geom: ["computegeom_2d"], f_norm: ["compute_norml2inf"],
        f_dz: ["coeffs_5p_dz", "coeffs_5p_dz2"], f_lindz: ["coeffs_5p_dz_d", "coeffs_5p_dz2_d"],
        f_misc: ["testvector", "testvector_partial", "computejacobianfromjv_relaxed", "computejacobianfromjv_relaxed_withjn",
                 "computejacobianfromjv_relaxed_withjnandcheck", "computejacobianfromjv_withjn", "computejacobianfromdz"],
        f_init: ["set_bndbl_2d"]}
for mod, names in need.items():
    for n in names:
        assert callable(eval("mod." + n)), (mod.__name__, n)     # the drivers resolve routines by eval(), BROADCAST_npz.py:1018
print("ok")
'''
    out = subprocess.run([sys.executable, "-c", code, os.path.join(root, "broadcast_b200", "dropin")], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_context_api_fails_loudly_without_a_device():
    """no CPU fallback: bcast_ctx_create reports BC_ERR_NODEV (-2) when there is no GPU (skipped on a GPU box)"""
    import ctypes
    from broadcast_b200 import _lib
    if _lib.device_count() > 0:
        import pytest
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    D = ctypes.c_double
    rc = _lib.lib().bcast_ctx_create(ctypes.byref(h), 10, 10, 3, *[D(1.0)] * 11, 1)
    assert rc == -2 and not h.value
