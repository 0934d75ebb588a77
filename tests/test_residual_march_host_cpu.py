"""The product's j-marching residual algorithm (broadcast_b200/csrc/residual_march.cuh is host+device code) compiled for the
HOST, the persistent CTA emulated phase by phase with its rings, checked against the reference residual run on oracle/_ref:
boundary-layer (wall scheme) and O-mesh cases, grids that are not multiples of the 32-column strip or of the 4-row step,
several segment lengths (the rings wrap many times), the nowall scheme, a non-zero spanwise velocity, k2 = 0 and an i-slab with
internal edges; and bit-for-bit against the tile algorithm of residual_fast.cuh (same formulas, different staging).  On the GPU
the same phase functions run inside k_residual_march (tests/test_parity_gpu.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "broadcast_b200", "csrc")


def _build(name, headers):
    src = os.path.join(HERE, "host", name + ".cpp")
    so = os.path.join(HERE, "host", "lib" + name + ".so")
    deps = [src] + [os.path.join(CSRC, f) for f in headers]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(scope="module")
def hostlib():
    return _build("residual_march_host", ("residual_march.cuh", "scheme.cuh", "grid.cuh", "dual.cuh"))


@pytest.fixture(scope="module")
def tilelib():
    return _build("residual_fast_host", ("residual_fast.cuh", "scheme.cuh", "grid.cuh", "dual.cuh"))


def _args(c, w, res, wall, k2):
    p, gh = c.phys, c.gh
    D = ctypes.c_double
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    return (P(res), P(w), P(c.nx), P(c.ny), P(c.vol), P(c.volf), gh, D(p["cp"]), D(p["cv"]), D(p["prandtl"]), D(p["gam"]), D(p["rgaz"]),
            D(p["cs"]), D(p["muref"]), D(p["tref"]), D(p["cs"]), D(c.k2 if k2 is None else k2), D(c.k4), c.im, c.jm, int(wall))


def host_residual(lib, c, w, wall=True, slab=(0, 0, 0), k2=None, seglen=16):
    res = np.zeros_like(w, order="F")
    assert lib.rm_host_residual(*_args(c, w, res, wall, k2), *slab, seglen) == 0
    return res


def tile_residual(lib, c, w, wall=True):
    res = np.zeros_like(w, order="F")
    assert lib.rf_host_residual(*_args(c, w, res, wall, None), 0, 0, 0, 0) == 0
    return res


@pytest.mark.parametrize("kind,im,jm,seglen", [("bl", 70, 21, 8), ("bl", 32, 8, 8), ("bl", 33, 9, 4), ("cyl", 45, 19, 12),
                                                ("bl", 40, 61, 64), ("bl", 64, 50, 20), ("cyl", 63, 37, 16), ("bl", 500, 150, 128),
                                                ("cyl", 630, 300, 64)])
def test_march_residual_matches_reference(ref, hostlib, kind, im, jm, seglen):
    c = H.make_case(kind, im, jm, ref, with_w=True)
    w, res_ref = H.residual_sequence(ref, c)
    res = host_residual(hostlib, c, w, seglen=seglen)
    gh = c.gh
    H.assert_residual_parity(res, res_ref, c, w, floor=H.fma_floor(c), what=(kind, im, jm))
    assert not np.any(res[:gh]) and not np.any(res[:, :gh]) and not np.any(res[-gh:]) and not np.any(res[:, -gh:])   # ghost frame untouched
    # the segment length only changes where the march restarts: same bits
    assert np.array_equal(host_residual(hostlib, c, w, seglen=4 * ((jm + 3) // 4)), res)


@pytest.mark.parametrize("kind,im,jm", [("bl", 70, 45), ("cyl", 63, 37)])
def test_march_equals_the_tile_algorithm_bit_for_bit(ref, hostlib, tilelib, kind, im, jm):
    c = H.make_case(kind, im, jm, ref, with_w=True)
    w, _ = H.residual_sequence(ref, c)
    a = host_residual(hostlib, c, w, seglen=16)
    b = tile_residual(tilelib, c, w)
    assert np.array_equal(a, b), np.abs(a - b).max()


def test_march_residual_nowall_spanwise_and_k2_zero(ref, hostlib):
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
        res = host_residual(hostlib, c, w, wall=wall, seglen=8)
        H.assert_residual_parity(res, res_ref, c, w, what=scheme)
    args = list(c.scheme_args())
    args[-4] = 0.0
    res_ref = c.zeros_state()
    ref["f_sch"].flux_num_dnc5_2d(res_ref, w, *args)
    res = host_residual(hostlib, c, w, k2=0.0, seglen=8)
    H.assert_residual_parity(res, res_ref, c, w, what="k2 = 0")


def test_march_residual_on_an_i_slab_equals_the_single_block(ref, hostlib):
    from broadcast_b200 import sharding
    c = H.make_case("bl", 70, 21, ref, with_w=True)
    w, _ = H.residual_sequence(ref, c)
    full = host_residual(hostlib, c, w, seglen=8)
    gh = c.gh
    cw = c
    cw.w = w
    for rank in range(3):
        cs, slab = sharding.slab_of(cw, rank, 3)
        part = host_residual(hostlib, cs, np.asfortranarray(cs.w), slab=slab, seglen=8)
        lo = slab[0]
        assert np.array_equal(part[gh:-gh, gh:-gh], full[gh + lo:gh + lo + cs.im, gh:-gh])
