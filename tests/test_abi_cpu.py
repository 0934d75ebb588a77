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


def test_scalars_are_coerced_to_the_declared_c_types():
    """The reference drivers pass 0-d arrays (cp = dic['Cp'], BROADCAST_npz.py:433-443), Python ints for real dummies
    (k4 = 1) and floats for integer ones (gh = 3.0, card_cyl2d.py:67).  Every wrapper of f2py_api is called with such
    arguments against a recording backend; the recorded lists must marshal to exactly the C types the header declares."""
    import ctypes
    from broadcast_b200 import _lib, f2py_api
    calls = []
    mods = f2py_api.build(lambda name, *a: calls.append((name, a)))
    im, jm, gh = 10, 8, 3
    st = lambda p=5: np.zeros((im + 2 * gh, jm + 2 * gh, p), order="F")
    nd = lambda p=2: np.zeros((im + 2 * gh + 1, jm + 2 * gh + 1, p), order="F")
    c2 = np.zeros((im + 2 * gh, jm + 2 * gh), order="F")
    n2 = np.zeros((im + 2 * gh + 1, jm + 2 * gh + 1), order="F")
    z = np.array  # 0-d arrays
    phys = (z(1004.5), z(717.5), z(0.72), z(1.4), z(287.0), z(0.38), z(1.7e-5), z(0.95), z(0.38), 1, 1)   # k2, k4 as ints
    geo = (n2, n2, nd(), nd(), c2, c2, c2, st(2))
    mods["f_sch"].flux_num_dnc5_2d(st(), st(), *geo, 3.0, *phys, im, jm)
    mods["f_lin"].flux_num_dnc5_2d_d(st(), st(), st(), st(), *geo, np.int64(gh), *phys)
    itf = np.array([[1.0, 1.0], [float(im), 1.0]])
    mods["f_bnd"].bc_wall_viscous_adia_2d(st(), "Jlo", z(1.4), itf, 3.0, np.int32(im), jm)
    mods["f_bnd"].bc_no_reflexion_2d(st(), np.zeros((im + gh, 5), order="F"), "Jhi", itf, nd(), nd(), 1, gh, im, jm)
    mods["f_bnd"].bc_supandsubinlet_2d(st(), "Ilo", itf, np.zeros((jm, gh, 5), order="F"), nd(), nd(), z(1.4), im, jm)
    mods["f_bnd"].bc_extrapolate_o2_2d(st(), "Ihi", itf, im, jm, z(3))
    mods["f_misc"].testvector(st(), z(0), 1.0, 2, gh, im, jm)
    nb = 25 * 49 * im * jm
    mods["f_misc"].computejacobianfromjv_relaxed(np.zeros(nb), np.zeros(nb, np.int32), np.zeros(nb, np.int32), st(), 0, 0, 0, gh,
                                                 np.zeros((im, jm), order="F"))
    mods["f_dz"].coeffs_5p_dz(st(), st(), st(), *geo, gh, *phys[:9], im, jm)
    mods["f_norm"].compute_norml2inf(st(), im, jm, 3.0)
    assert len(calls) >= 10
    want = {"f64": ctypes.c_double, "i32": ctypes.c_int, "i64": ctypes.c_int64, "ptr": ctypes.c_void_p, "str": ctypes.c_char_p}
    for name, args in calls:
        kinds = _lib.signatures()["bc_" + name]
        cargs = _lib.marshal("bc_" + name, args)
        assert [type(c) for c in cargs] == [want[k] for k in kinds], name
    name, args = calls[0]
    cargs = _lib.marshal("bc_" + name, args)
    assert cargs[10].value == 3 and cargs[11].value == 1004.5 and cargs[20].value == 1.0 and cargs[21].value == 1.0
    # a non-integral value for an integer dummy is an error, not a truncation
    with pytest.raises(ValueError):
        _lib.marshal("bc_testvector", (st(), 0.5, 0, 0, gh, im, jm))
    # every bc_* entry point of the header has a parsed signature
    assert set(s for s in declared_symbols() if s.startswith("bc_") and s not in ("bc_last_error", "bc_launch_count", "bc_desc_t")) <= set(
        _lib.signatures())
