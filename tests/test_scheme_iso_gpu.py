"""SURVEY.md 8(f2), `_iso`: the isothermal-wall variant of the order-5 scheme, f_sch.flux_num_dnc5_iso_2d
(srcfv/rhs/flux_num_dnc5_iso.F90:7-227) and its tangent f_lin.flux_num_dnc5_iso_2d_d (srcfv/tangent/flux_num_dnc5_iso_d.f90), through
the drop-in entry points against oracle/_ref, and in resident mode (bcd_wall_iso) on every fused kernel variant."""
import numpy as np
import pytest

import helpers as H
from broadcast_b200 import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("im,jm", [(70, 21), (300, 70)])
def test_iso_scheme_residual_and_tangent(gpu, ref, im, jm):
    a = H.make_case("bl", im, jm, gpu, with_w=True)
    b = H.make_case("bl", im, jm, ref, with_w=True)
    twall = 1.15
    wa, _ = H.residual_sequence(gpu, a)
    wb, res_adia = H.residual_sequence(ref, b)
    ra, rb = a.zeros_state(), b.zeros_state()
    gpu["f_sch"].flux_num_dnc5_iso_2d(ra, wa, twall, *a.scheme_args())
    ref["f_sch"].flux_num_dnc5_iso_2d(rb, wb, twall, *b.scheme_args())
    H.assert_residual_parity(ra, rb, b, wb, floor=None, what="iso residual")
    gh = a.gh
    assert np.all(H.rel_err(ra[gh:-gh, gh:-gh], rb[gh:-gh, gh:-gh]) < 1e-12)
    assert np.abs(rb - res_adia)[gh:-gh, gh, 4].max() > 0          # the variant is not the adiabatic scheme
    # the context is scoped to the call: the adiabatic entry point right after gives the adiabatic residual again
    r2 = a.zeros_state()
    gpu["f_sch"].flux_num_dnc5_2d(r2, wa, *a.scheme_args())
    assert np.all(H.rel_err(r2[gh:-gh, gh:-gh], res_adia[gh:-gh, gh:-gh]) < 1e-12)
    # tangent along a random direction (twall passive)
    rng = np.random.default_rng(5)
    wd = np.asfortranarray(rng.standard_normal(wa.shape))
    out = []
    for mods, c, w in ((gpu, a, wa), (ref, b, wb)):
        w2, wd2 = w.copy(order="F"), wd.copy(order="F")
        cases.apply_bcs_lin(c, w2, wd2, mods["f_bnd"], mods["f_lin"])
        res, resd = c.zeros_state(), c.zeros_state()
        mods["f_lin"].flux_num_dnc5_iso_2d_d(res, resd, w2, wd2, twall, *c.scheme_args())
        out.append(resd)
    assert np.all(H.rel_err(out[0][gh:-gh, gh:-gh], out[1][gh:-gh, gh:-gh]) < 1e-12)


def test_iso_scheme_resident_variants(gpu, ref):
    """bcd_wall_iso in resident mode: tile, bulk-staged, marching and reference-shaped kernels all evaluate the isothermal wall flux"""
    import ctypes
    import torch
    from broadcast_b200 import _lib
    from broadcast_b200.resident import Block
    c = H.make_case("bl", 96, 48, gpu, with_w=True)
    b = H.make_case("bl", 96, 48, ref, with_w=True)
    wb, _ = H.residual_sequence(ref, b)
    rb = b.zeros_state()
    twall = 0.9
    ref["f_sch"].flux_num_dnc5_iso_2d(rb, wb, twall, *b.scheme_args())
    blk = Block(c)
    blk.apply_bcs()
    L = _lib.lib()
    gh = c.gh
    try:
        _lib.check(L.bcd_wall_iso(1, ctypes.c_double(twall)), "bcd_wall_iso")
        for v in (4, 6, 5, 1):
            r = blk.residual(variant=v).clone()
            got = np.asfortranarray(r.cpu().numpy().transpose(2, 1, 0))
            assert np.all(H.rel_err(got[gh:-gh, gh:-gh], rb[gh:-gh, gh:-gh]) < 1e-12), v
    finally:
        L.bcd_wall_iso(0, ctypes.c_double(0.0))
