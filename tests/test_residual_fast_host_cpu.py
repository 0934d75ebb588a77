"""The product's fused residual tile algorithm (broadcast_b200/csrc/residual_fast.cuh is host+device code) compiled for
the HOST, the CTA emulated phase by phase, checked against the reference residual run on oracle/_ref: boundary-layer
(wall scheme) and O-mesh (periodic in i) cases, grids that are not multiples of the 32 x 8 tile, the nowall scheme, a
non-zero spanwise velocity, k2 = 0, and an i-slab with internal edges.  Tolerance 1e-12 of the plane maximum (the fast
formulas are re-associations of the reference's).  On the GPU the same phase functions run inside k_residual_fast
(tests/test_parity_gpu.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "residual_fast_host.cpp")
SO = os.path.join(HERE, "host", "libresidual_fast_host.so")


@pytest.fixture(scope="module")
def hostlib():
    deps = [SRC] + [os.path.join(HERE, "..", "broadcast_b200", "csrc", f) for f in ("residual_fast.cuh", "scheme.cuh", "grid.cuh", "dual.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    return ctypes.CDLL(SO)


def host_residual(lib, c, w, wall=True, slab=(0, 0, 0), k2=None, staged=0):
    p, gh = c.phys, c.gh
    D = ctypes.c_double
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    res = np.zeros_like(w, order="F")
    rc = lib.rf_host_residual(P(res), P(w), P(c.nx), P(c.ny), P(c.vol), P(c.volf), gh, D(p["cp"]), D(p["cv"]), D(p["prandtl"]),
                              D(p["gam"]), D(p["rgaz"]), D(p["cs"]), D(p["muref"]), D(p["tref"]), D(p["cs"]),
                              D(c.k2 if k2 is None else k2), D(c.k4), c.im, c.jm, int(wall), *slab, staged)
    assert rc == 0
    return res


@pytest.mark.parametrize("kind,im,jm", [("bl", 70, 21), ("bl", 32, 8), ("bl", 33, 9), ("cyl", 45, 19)])
def test_fast_tile_residual_matches_reference(ref, hostlib, kind, im, jm):
    c = H.make_case(kind, im, jm, ref, with_w=True)
    w, res_ref = H.residual_sequence(ref, c)
    res = host_residual(hostlib, c, w)
    err = H.rel_err(res[c.gh:-c.gh, c.gh:-c.gh], res_ref[c.gh:-c.gh, c.gh:-c.gh])
    assert np.all(err < 1e-12), err
    assert not np.any(res[:c.gh]) and not np.any(res[:, :c.gh])   # ghost frame of residu untouched
    # w planes delivered as the TMA box (zero fill outside the padded array) instead of loaded cell by cell: same bits
    assert np.array_equal(host_residual(hostlib, c, w, staged=1), res)
    # k_residual_fast_bulk: metrics from the shared-memory images of the vol / volf boxes and the (shifted) node rows: same bits
    assert np.array_equal(host_residual(hostlib, c, w, staged=2), res)


def test_fast_tile_residual_nowall_spanwise_and_k2_zero(ref, hostlib):
    c = H.make_case("bl", 41, 18, ref, with_w=True)
    rng = np.random.default_rng(3)
    w = c.w.copy(order="F")
    w[:, :, 3] = 0.05 * w[:, :, 0] * (1.0 + 0.1 * rng.standard_normal(w.shape[:2]))   # rho*w != 0
    from broadcast_b200 import cases
    cases.apply_bcs(c, w, ref["f_bnd"])
    gh = c.gh
    for scheme, wall in (("flux_num_dnc5_2d", True), ("flux_num_dnc5_nowall_2d", False)):
        res_ref = c.zeros_state()
        getattr(ref["f_sch"], scheme)(res_ref, w, *c.scheme_args())
        res = host_residual(hostlib, c, w, wall=wall)
        err = H.rel_err(res[gh:-gh, gh:-gh], res_ref[gh:-gh, gh:-gh])
        assert np.all(err < 1e-12), (scheme, err)
    # k2 = 0 (the cylinder cards): the sensor is skipped
    args = list(c.scheme_args())
    args[-4] = 0.0
    res_ref = c.zeros_state()
    ref["f_sch"].flux_num_dnc5_2d(res_ref, w, *args)
    res = host_residual(hostlib, c, w, k2=0.0)
    err = H.rel_err(res[gh:-gh, gh:-gh], res_ref[gh:-gh, gh:-gh])
    assert np.all(err < 1e-12), err


def test_fast_tile_residual_on_an_i_slab_equals_the_single_block(ref, hostlib):
    from broadcast_b200 import sharding
    c = H.make_case("bl", 70, 21, ref, with_w=True)
    w, _ = H.residual_sequence(ref, c)
    full = host_residual(hostlib, c, w)
    gh = c.gh
    cw = c
    cw.w = w
    for rank in range(3):
        cs, slab = sharding.slab_of(cw, rank, 3)
        part = host_residual(hostlib, cs, np.asfortranarray(cs.w), slab=slab)
        lo = slab[0]
        assert np.array_equal(part[gh:-gh, gh:-gh], full[gh + lo:gh + lo + cs.im, gh:-gh])


def test_isothermal_wall_scheme_variant(ref, hostlib):
    """flux_num_dnc5_iso_2d (srcfv/rhs/flux_num_dnc5_iso.F90: rhs/fluxwall_iso.F instead of rhs/fluxwall.F) through the product's wall-row
    templates (scheme.cuh, host build) against the reference routine: only the energy flux through the wall face differs"""
    c = H.make_case("bl", 70, 21, ref, with_w=True)
    w, res_adia = H.residual_sequence(ref, c)
    twall = 1.2 * float(c.phys.get("tinf", 1.0)) if "tinf" in c.phys else 1.1
    res_ref = c.zeros_state()
    ref["f_sch"].flux_num_dnc5_iso_2d(res_ref, w, twall, *c.scheme_args())
    hostlib.rf_host_wall_iso(1, ctypes.c_double(twall))
    try:
        res = host_residual(hostlib, c, w)
        res_b = host_residual(hostlib, c, w, staged=2)
    finally:
        hostlib.rf_host_wall_iso(0, ctypes.c_double(0.0))
    gh = c.gh
    err = H.rel_err(res[gh:-gh, gh:-gh], res_ref[gh:-gh, gh:-gh])
    assert np.all(err < 1e-12), err
    assert np.array_equal(res, res_b)
    # the variant really differs from the adiabatic scheme, and only in the energy equation of the first row
    d = np.abs(res_ref - res_adia)
    assert d[gh:-gh, gh, 4].max() > 0 and d[gh:-gh, gh + 1:-gh].max() == 0 and d[..., :4].max() == 0
