"""The product's fused TANGENT tile algorithm (residual_fast.cuh compiled in dual-number arithmetic, 32 x 3 tiles) built for
the HOST, CTA emulated phase by phase, checked against the reference's Tapenade tangent (srcfv/tangent/flux_num_dnc5_d.f90)
run on oracle/_ref: random dense directions and colour seeds, boundary-layer (wall rows) and O-mesh cases, nowall scheme,
a rectangle of rows only.  Tolerance 1e-12 of the plane maximum.  On the GPU the same phase functions run in k_tangent_tile."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from broadcast_b200 import cases

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "residual_tangent_host.cpp")
SO = os.path.join(HERE, "host", "libresidual_tangent_host.so")


@pytest.fixture(scope="module")
def hostlib():
    deps = [SRC] + [os.path.join(HERE, "..", "broadcast_b200", "csrc", f) for f in ("residual_fast.cuh", "scheme.cuh", "grid.cuh", "dual.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    return ctypes.CDLL(SO)


def host_tangent(lib, c, w, wd, wall=True, rect=None):
    p, gh = c.phys, c.gh
    D = ctypes.c_double
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    out = np.zeros_like(w, order="F")
    r = rect or (1, c.im, 1, c.jm)
    rc = lib.rfd_host_tangent(P(out), P(w), P(wd), P(c.nx), P(c.ny), P(c.vol), P(c.volf), gh, D(p["cp"]), D(p["cv"]), D(p["prandtl"]),
                              D(p["gam"]), D(p["rgaz"]), D(p["cs"]), D(p["muref"]), D(p["tref"]), D(p["cs"]), D(c.k2), D(c.k4),
                              c.im, c.jm, int(wall), *r)
    assert rc == 0
    return out


def ref_tangent(ref, c, w, wd, scheme="flux_num_dnc5_2d_d"):
    ww, wdd = w.copy(order="F"), wd.copy(order="F")
    cases.apply_bcs_lin(c, ww, wdd, ref["f_bnd"], ref["f_lin"])
    res, resd = c.zeros_state(), c.zeros_state()
    getattr(ref["f_lin"], scheme)(res, resd, ww, wdd, *c.scheme_args())
    return ww, wdd, resd


@pytest.mark.parametrize("kind,im,jm", [("bl", 40, 14), ("bl", 33, 9), ("cyl", 45, 17)])
def test_tangent_tile_matches_the_reference_tangent(ref, hostlib, kind, im, jm):
    c = H.make_case(kind, im, jm, ref, with_w=True)
    w, _ = H.residual_sequence(ref, c)
    gh = c.gh
    rng = np.random.default_rng(7)
    # a dense random direction, then two colour seeds (sparse directions, as in the Jacobian loop)
    dirs = [np.asfortranarray(rng.standard_normal(w.shape))]
    for (m, l, k) in [(0, 2, 1), (4, 5, 6)]:
        wd = c.zeros_state()
        ref["f_misc"].testvector(wd, m, l, k, gh, im, jm)
        dirs.append(wd)
    for wd in dirs:
        ww, wdd, resd = ref_tangent(ref, c, w, wd)
        out = host_tangent(hostlib, c, ww, wdd)
        err = H.rel_err(out[gh:-gh, gh:-gh], resd[gh:-gh, gh:-gh])
        assert np.all(err < 1e-12), err


def test_tangent_tile_nowall_and_row_rectangle(ref, hostlib):
    c = H.make_case("bl", 41, 16, ref, with_w=True)
    w, _ = H.residual_sequence(ref, c)
    gh = c.gh
    wd = np.asfortranarray(np.random.default_rng(9).standard_normal(w.shape))
    ww, wdd, resd = ref_tangent(ref, c, w, wd, "flux_num_dnc5_nowall_2d_d")
    out = host_tangent(hostlib, c, ww, wdd, wall=False)
    assert np.all(H.rel_err(out[gh:-gh, gh:-gh], resd[gh:-gh, gh:-gh]) < 1e-12)
    # only the rows of a rectangle (a boundary strip): the other rows stay untouched
    ww, wdd, resd = ref_tangent(ref, c, w, wd)
    rect = (1, c.im, c.jm - 2, c.jm)
    out = host_tangent(hostlib, c, ww, wdd, rect=rect)
    sl = (slice(gh, -gh), slice(gh + rect[2] - 1, gh + rect[3]))
    scale = np.abs(resd[gh:-gh, gh:-gh]).max(axis=(0, 1))
    assert np.all(np.abs(out[sl] - resd[sl]).max(axis=(0, 1)) < 1e-12 * scale)
    assert not np.any(out[gh:-gh, gh:gh + rect[2] - 1])
