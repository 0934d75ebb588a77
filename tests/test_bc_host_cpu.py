"""The product's boundary-fill line functions (broadcast_b200/csrc/bc.cuh: the four fills of the named cards and the seven of SURVEY.md
8(f3)) built for the HOST unchanged (tests/host/bc_host.cpp), primal and tangent arithmetic, against the reference routines
(srcfv/prepro/bc_*.f90, srcfv/tangent/bc_*_d.f90) run on oracle/_ref: every side of a boundary-layer and an O-mesh block, 1e-12 of the
plane maximum.  On the GPU the same functions run one thread per boundary-line cell (k_bc_*); tests/test_parity_gpu.py checks that."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import helpers as H
import test_extra_bcs_cpu as T

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "bc_host.cpp")
SO = os.path.join(HERE, "host", "libbc_host.so")
TOL = 1e-12
WHICH = {"wall": 0, "noref": 1, "inflow": 2, "outflow": 3, "iso": 4, "sym": 5, "anti": 6, "pres": 7, "presnr": 7, "blow": 8, "isoprof": 9}


@pytest.fixture(scope="module")
def hostlib():
    deps = [SRC] + [os.path.join(HERE, "..", "broadcast_b200", "csrc", f) for f in ("bc.cuh", "grid.cuh", "dual.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    return ctypes.CDLL(SO)


def host_fill(lib, c, name, w, wd, loc, interf, p=(0.0, 0.0, 0.0), tab=None, tabd=None):
    D, P = ctypes.c_double, lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else ctypes.c_void_p(None)
    it = np.ascontiguousarray(np.asarray(interf, dtype=np.int32).ravel())
    tab = np.asfortranarray(tab, dtype=np.float64) if tab is not None else None
    tabd = np.ascontiguousarray(tabd, dtype=np.float64) if tabd is not None else None
    rc = lib.bc_host_fill(WHICH[name], P(w), P(wd), loc.encode(), P(it), D(c.phys["gam"]), D(p[0]), D(p[1]), D(p[2]), P(c.nx), P(c.ny), P(tab),
                          P(tabd), int(tab.shape[0]) if tab is not None else 0, c.gh, c.im, c.jm)
    assert rc == 0


def extra_params(c, name, loc, interf, tangent):
    p = c.phys
    if name == "iso":
        return (T.TWALL, p["rgaz"], 0.0), None, None
    if name in ("pres", "presnr"):
        g = c.gh
        q = c.w[g, g]
        return (0.97 * (p["gam"] - 1.0) * (q[4] - 0.5 * (q[1] ** 2 + q[2] ** 2 + q[3] ** 2) / q[0]), float(name == "presnr"), 0.0), None, None
    if name == "blow":
        pr, prd = T.profiles(c, name, loc, interf)
        return (T.GAMD if tangent else 0.0, 0.0, 0.0), pr, prd
    if name == "isoprof":
        pr, prd = T.profiles(c, name, loc, interf)
        return (T.GAMD if tangent else 0.0, p["rgaz"], T.RGAZD if tangent else 0.0), pr, prd
    return (0.0, 0.0, 0.0), None, None


@pytest.mark.parametrize("name", T.NAMES)
@pytest.mark.parametrize("kind,im,jm", [("bl", 24, 14), ("cyl", 26, 12), ("bl", 7, 7)])
def test_extra_fills_host_build_vs_reference(ref, hostlib, name, kind, im, jm):
    c = H.make_case(kind, im, jm, ref, with_w=True)
    w0, _ = H.residual_sequence(ref, c)
    d = np.asfortranarray(np.random.default_rng(4).standard_normal(w0.shape))
    for loc, interf in T.sides(c):
        for tangent in (False, True):
            wr, wdr = w0.copy(order="F"), d.copy(order="F")
            T.fill(ref, name, c, wr, loc, interf, wdr if tangent else None)
            wh, wdh = w0.copy(order="F"), d.copy(order="F")
            p, tab, tabd = extra_params(c, name, loc, interf, tangent)
            host_fill(hostlib, c, name, wh, wdh if tangent else None, loc, interf, p, tab, tabd if tangent else None)
            assert np.all(H.rel_err(wh, wr) < TOL), (name, loc, tangent, H.rel_err(wh, wr))
            if tangent:
                assert np.all(H.rel_err(wdh, wdr) < TOL), (name, loc, H.rel_err(wdh, wdr))


@pytest.mark.parametrize("kind,im,jm", [("bl", 24, 14), ("cyl", 26, 12)])
def test_card_fills_host_build_vs_reference(ref, hostlib, kind, im, jm):
    """the boundary list of the case itself (inlet, non-reflecting, extrapolation, adiabatic wall), one fill at a time"""
    c = H.make_case(kind, im, jm, ref, with_w=True)
    w0 = c.w.copy(order="F")
    d = np.asfortranarray(np.random.default_rng(5).standard_normal(w0.shape))
    gam, gh = c.phys["gam"], c.gh
    seen = set()
    for bc in c.bcs:
        k = bc[0]
        if k == "jn":
            continue
        seen.add(k)
        loc, interf = bc[1], bc[2]
        for tangent in (False, True):
            wr, wdr = w0.copy(order="F"), d.copy(order="F")
            fb, fl = ref["f_bnd"], ref["f_lin"]
            if k == "wall":
                fl.bc_wall_viscous_adia_2d_d(wr, wdr, loc, gam, interf, gh, im, jm) if tangent else fb.bc_wall_viscous_adia_2d(wr, loc, gam, interf, gh, im, jm)
            elif k == "outflow":
                fl.bc_extrapolate_o2_2d_d(wr, wdr, loc, interf, im, jm, gh) if tangent else fb.bc_extrapolate_o2_2d(wr, loc, interf, im, jm, gh)
            elif k == "noref":
                (fl.bc_no_reflexion_2d_d(wr, wdr, bc[3], loc, interf, c.nx, c.ny, gam, gh, im, jm) if tangent
                 else fb.bc_no_reflexion_2d(wr, bc[3], loc, interf, c.nx, c.ny, gam, gh, im, jm))
            else:
                (fl.bc_supandsubinlet_2d_d(wr, wdr, loc, interf, bc[3], c.nx, c.ny, gam, im, jm) if tangent
                 else fb.bc_supandsubinlet_2d(wr, loc, interf, bc[3], c.nx, c.ny, gam, im, jm))
            wh, wdh = w0.copy(order="F"), d.copy(order="F")
            host_fill(hostlib, c, k, wh, wdh if tangent else None, loc, interf, tab=bc[3] if len(bc) > 3 else None)
            assert np.all(H.rel_err(wh, wr) < TOL), (k, tangent, H.rel_err(wh, wr))
            if tangent:
                assert np.all(H.rel_err(wdh, wdr) < TOL), (k, H.rel_err(wdh, wdr))
    assert seen >= ({"noref", "wall"} if kind == "cyl" else {"inflow", "noref", "outflow", "wall"})
