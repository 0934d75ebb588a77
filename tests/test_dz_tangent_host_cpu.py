"""The product's fused tangent-of-Dz tile algorithm (broadcast_b200/csrc/dz_tangent.cuh: hyper-dual arithmetic, 32 x 8 tiles) built
for the HOST, CTA emulated phase by phase, checked against the reference's Tapenade code (srcfv/tangentdz/coeffs_5p_dz_d.f90,
coeffs_5p_dz2_d.f90 = f_lindz of BROADCAST_npz_sens.py:1768-1797) run on oracle/_ref: dense random directions and a colour seed as the
base-flow variation, boundary-layer and O-mesh grids, tile-ragged sizes.  Tolerance 1e-12 of the plane maximum.  Also pins the
reference routine itself: a central finite difference of its un-differentiated parent f_dz.coeffs_5p_dz along wd0."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "dz_tangent_host.cpp")
SO = os.path.join(HERE, "host", "libdz_tangent_host.so")
TOL = 1e-12


@pytest.fixture(scope="module")
def hostlib():
    deps = [SRC] + [os.path.join(HERE, "..", "broadcast_b200", "csrc", f) for f in ("dz_tangent.cuh", "grid.cuh", "dual.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    return ctypes.CDLL(SO)


def dz_args(c):
    a = c.scheme_args()
    return a[:18] + a[20:]   # no k2, k4 (BROADCAST_npz_sens.py:1768)


def host_dz_d(lib, c, w, wd0, wd):
    p = c.phys
    D = ctypes.c_double
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    o1, o2 = c.zeros_state(), c.zeros_state()
    rc = lib.dzt_host(P(o1), P(o2), P(w), P(wd0), P(wd), P(c.nx), P(c.ny), P(c.vol), c.gh, D(p["cp"]), D(p["cv"]), D(p["prandtl"]),
                      D(p["gam"]), D(p["cs"]), D(p["muref"]), D(p["tref"]), D(p["cs"]), c.im, c.jm)
    assert rc == 0
    return o1, o2


def ref_dz_d(ref, c, w, wd0, wd):
    out = []
    for name in ("coeffs_5p_dz_d", "coeffs_5p_dz2_d"):
        dz = np.asfortranarray(np.full(w.shape, 7.0))
        dzd = np.asfortranarray(np.full(w.shape, 5.0))
        getattr(ref["f_lindz"], name)(dz, dzd, w, wd0, wd, *dz_args(c))
        assert np.all(dz == 7.0)                                   # dz_out is never assigned (sliced by Tapenade)
        g = c.gh
        assert np.all(dzd[:g] == 0.0) and np.all(dzd[:, :g] == 0.0) and np.all(dzd[-g:] == 0.0) and np.all(dzd[:, -g:] == 0.0)
        out.append(dzd)
    return out


def directions(ref, c, w, seed):
    rng = np.random.default_rng(seed)
    wd = np.asfortranarray(rng.standard_normal(w.shape))
    wd0 = np.asfortranarray(rng.standard_normal(w.shape) * np.abs(w).max(axis=(0, 1)))
    wseed = c.zeros_state()
    ref["f_misc"].testvector(wseed, 2, 3, 1, c.gh, c.im, c.jm)
    return wd, wd0, wseed


@pytest.mark.parametrize("kind,im,jm", [("bl", 40, 14), ("bl", 33, 9), ("bl", 64, 16), ("cyl", 45, 17), ("bl", 5, 3)])
def test_dz_tangent_tile_matches_the_reference(ref, hostlib, kind, im, jm):
    c = H.make_case(kind, im, jm, ref, with_w=True)
    w, _ = H.residual_sequence(ref, c)
    wd, wd0, wseed = directions(ref, c, w, 11)
    for a, b in ((wd0, wd), (wseed, wd), (wd0, wseed)):
        h1, h2 = host_dz_d(hostlib, c, w, a, b)
        r1, r2 = ref_dz_d(ref, c, w, a, b)
        assert np.abs(r1).max() > 0 and np.abs(r2).max() > 0
        assert np.all(H.rel_err(h1, r1) < TOL), H.rel_err(h1, r1)
        assert np.all(H.rel_err(h2, r2) < TOL), H.rel_err(h2, r2)


def test_reference_dz_tangent_is_the_derivative_of_the_operator_rows(ref):
    """pins oracle/_ref's tangentdz translation: d/d eps f_dz.coeffs_5p_dz(w + eps wd0; wd) by central differences"""
    c = H.make_case("bl", 30, 12, ref, with_w=True)
    w, _ = H.residual_sequence(ref, c)
    wd, wd0, _ = directions(ref, c, w, 3)
    wd0 *= 1e-2
    r = ref_dz_d(ref, c, w, wd0, wd)
    for name, rd in zip(("coeffs_5p_dz", "coeffs_5p_dz2"), r):
        eps = 1e-4
        zp, zm = c.zeros_state(), c.zeros_state()
        getattr(ref["f_dz"], name)(zp, np.asfortranarray(w + eps * wd0), wd, *dz_args(c))
        getattr(ref["f_dz"], name)(zm, np.asfortranarray(w - eps * wd0), wd, *dz_args(c))
        fd = (zp - zm) / (2 * eps)
        assert np.all(H.rel_err(fd, rd) < 1e-6), (name, H.rel_err(fd, rd))


GOLD = sorted(__import__("glob").glob(os.path.join(HERE, "golden", "lindz", "*.npz")))


def test_lindz_golden_fixtures_present():
    assert len(GOLD) >= 2


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_ref_and_host_build_reproduce_the_lindz_golden(ref, hostlib, path):
    """tests/golden/lindz/*.npz = outputs of the reference's tangentdz code (oracle/make_golden.py --dz-tangent): oracle/_ref rebuilt
    here reproduces them (libm differences only) and the product's tile algorithm matches them to 1e-12"""
    g = np.load(path)
    c = H.make_case(str(g["kind"]), int(g["im"]), int(g["jm"]), ref, with_w=True)
    w, _ = H.residual_sequence(ref, c)
    wd0, wd = np.asfortranarray(g["wd0"]), np.asfortranarray(g["wd"])
    r1, r2 = ref_dz_d(ref, c, w, wd0, wd)
    assert np.all(H.rel_err(r1, g["dzd"]) < 1e-13) and np.all(H.rel_err(r2, g["dz2d"]) < 1e-13)
    h1, h2 = host_dz_d(hostlib, c, w, wd0, wd)
    assert np.all(H.rel_err(h1, g["dzd"]) < TOL) and np.all(H.rel_err(h2, g["dz2d"]) < TOL)
