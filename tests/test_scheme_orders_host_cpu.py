"""Row f2 of SURVEY.md section 8: the other members of the scheme family, orders 3 / 7 / 9 (srcfv/rhs/flux_num_dnc{3,7,9}.F90,
their _nowall variants and the Tapenade tangents srcfv/tangent/flux_num_dnc{3,7,9}_d.f90).  The product evaluates them with the
order-templated face formulas of broadcast_b200/csrc/scheme.cuh through the kernels of generic_impl.cuh; here the SAME templates
are built for the host (tests/host/scheme_orders_host.cpp, the kernels' per-cell bodies as loops) and compared with the reference
routines run on oracle/_ref: boundary-layer block with its wall rows (off-centred Euler fluxes on rows 2 .. gh, wall flux on row 1),
periodic O-mesh, the nowall variants, a dense random tangent direction and a colouring seed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from broadcast_b200 import cases

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "scheme_orders_host.cpp")
SO = os.path.join(HERE, "host", "libscheme_orders_host.so")


@pytest.fixture(scope="module")
def hostlib():
    deps = [SRC] + [os.path.join(HERE, "..", "broadcast_b200", "csrc", f) for f in ("scheme.cuh", "grid.cuh", "dual.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    return ctypes.CDLL(SO)


def host_eval(lib, order, c, w, wd=None, wall=True):
    p, gh = c.phys, c.gh
    D = ctypes.c_double
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    out = np.zeros_like(w, order="F")
    rc = lib.so_host_residual(order, P(out), P(w), P(wd), P(c.nx), P(c.ny), P(c.vol), P(c.volf), gh, D(p["cp"]), D(p["cv"]), D(p["prandtl"]),
                              D(p["gam"]), D(p["rgaz"]), D(p["cs"]), D(p["muref"]), D(p["tref"]), D(p["cs"]), D(c.k2), D(c.k4), c.im, c.jm,
                              int(wall))
    assert rc == 0
    return out


@pytest.mark.parametrize("order", [3, 5, 7, 9])
@pytest.mark.parametrize("kind,im,jm,wall", [("bl", 37, 23, True), ("bl", 30, 26, False), ("cyl", 44, 25, True)])
def test_residual_of_every_order(ref, hostlib, order, kind, im, jm, wall):
    c = H.make_case(kind, im, jm, ref, with_w=True, order=order)
    gh = c.gh
    assert gh == (order + 1) // 2
    name = f"flux_num_dnc{order}{'' if wall else '_nowall'}_2d"
    w, res = H.residual_sequence(ref, c, name)
    out = host_eval(hostlib, order, c, w, wall=wall)
    err = H.rel_err(out[gh:-gh, gh:-gh], res[gh:-gh, gh:-gh])
    # same operation order as the Fortran, no contraction on either side: agreement to the last few ulps of the face fluxes
    assert np.all(err < 1e-14), (order, err)


@pytest.mark.parametrize("order", [3, 7, 9])
@pytest.mark.parametrize("kind,im,jm,wall", [("bl", 33, 24, True), ("cyl", 40, 23, False)])
def test_tangent_of_every_order(ref, hostlib, order, kind, im, jm, wall):
    c = H.make_case(kind, im, jm, ref, with_w=True, order=order)
    gh = c.gh
    name = f"flux_num_dnc{order}{'' if wall else '_nowall'}_2d_d"
    w, _ = H.residual_sequence(ref, c)   # ghosts of w filled
    rng = np.random.default_rng(order)
    dirs = [np.asfortranarray(rng.standard_normal(w.shape))]
    seed = c.zeros_state()
    ref["f_misc"].testvector(seed, 1, 2, gh, gh, im, jm)
    dirs.append(seed)
    for wd in dirs:
        wdd, resd = H.tangent_sequence(ref, c, w, wd, name)
        ww = w.copy(order="F")
        wd2 = wd.copy(order="F")
        cases.apply_bcs_lin(c, ww, wd2, ref["f_bnd"], ref["f_lin"])
        out = host_eval(hostlib, order, c, ww, wd2, wall=wall)
        scale = np.abs(resd[gh:-gh, gh:-gh]).max()
        err = np.abs(out[gh:-gh, gh:-gh] - resd[gh:-gh, gh:-gh]).max() / scale
        assert err < 1e-12, (order, err)
