"""Row f2 of SURVEY.md section 8 on the GPU: the scheme family orders 3 / 7 / 9 (srcfv/rhs/flux_num_dnc{3,7,9}.F90, _nowall variants,
srcfv/tangent/flux_num_dnc{3,7,9}_d.f90) through the drop-in entry points (C ABI -> order-templated kernels of generic_impl.cuh)
against the reference routines on oracle/_ref: boundary fills with gh = 2 / 4 / 5 ghost layers, residual (backward-error bar of
DESIGN.md section 2 and the forward error against the reference's own FMA / no-FMA spread), tangent, and the whole Jacobian by the
device colour loop ((2 gh + 1)^2 passes of five directions) against the reference's 5 (2 gh + 1)^2-colour host loop."""
import numpy as np
import pytest

import helpers as H
from broadcast_b200 import cases

pytestmark = pytest.mark.gpu

ORDERS = [3, 7, 9]


@pytest.mark.parametrize("order", ORDERS)
@pytest.mark.parametrize("kind,im,jm,wall", [("bl", 97, 41, True), ("bl", 64, 30, False), ("cyl", 90, 37, True)])
def test_residual_orders(gpu, ref, order, kind, im, jm, wall):
    a = H.make_case(kind, im, jm, gpu, with_w=True, order=order)
    b = H.make_case(kind, im, jm, ref, with_w=True, order=order)
    gh = a.gh
    assert gh == (order + 1) // 2
    name = f"flux_num_dnc{order}{'' if wall else '_nowall'}_2d"
    wa, ra = H.residual_sequence(gpu, a, name)
    wb, rb = H.residual_sequence(ref, b, name)
    assert np.all(H.rel_err(wa, wb) < 1e-12), H.rel_err(wa, wb)   # boundary fills with gh ghost layers (FMA contraction on the device)
    # backward error below 1e-13; forward error within 1e-12 or four times the spread between the reference's own FMA and no-FMA
    # builds on this input, whichever is larger
    H.assert_residual_parity(ra, rb, b, wb, floor=H.fma_floor(b, name), what=name)


@pytest.mark.parametrize("order", ORDERS)
@pytest.mark.parametrize("kind,im,jm,wall", [("bl", 60, 33, True), ("cyl", 70, 29, False)])
def test_tangent_orders(gpu, ref, order, kind, im, jm, wall):
    a = H.make_case(kind, im, jm, gpu, with_w=True, order=order)
    b = H.make_case(kind, im, jm, ref, with_w=True, order=order)
    gh = a.gh
    name = f"flux_num_dnc{order}{'' if wall else '_nowall'}_2d_d"
    wa, _ = H.residual_sequence(gpu, a, name[:-2])
    wb, _ = H.residual_sequence(ref, b, name[:-2])
    wd = np.asfortranarray(np.random.default_rng(order).standard_normal(wa.shape))
    wda, rda = H.tangent_sequence(gpu, a, wa, wd, name)
    wdb, rdb = H.tangent_sequence(ref, b, wb, wd, name)
    assert np.all(H.rel_err(wda, wdb) < 1e-12)          # linearised boundary fills
    assert np.all(H.rel_err(rda[gh:-gh, gh:-gh], rdb[gh:-gh, gh:-gh]) < 1e-12)
    assert not rda[:gh].any() and not rda[:, :gh].any()  # residud zeroed outside the interior, as the Tapenade routine leaves it


@pytest.mark.parametrize("order,kind,im,jm", [(3, "bl", 26, 21), (7, "bl", 30, 22), (9, "bl", 25, 24), (7, "cyl", 36, 20)])
def test_device_colour_loop_orders(gpu, ref, order, kind, im, jm):
    """whole Jacobian of an order-3 / 7 / 9 block: device colour loop vs the reference's host loop on the oracle; IA / JA bit-exact
    in the reference's slot order, values to 1e-12 of the largest entry, filtered pattern identical up to rounding-level entries"""
    from broadcast_b200.resident import Block, jacobian_coo
    a = H.make_case(kind, im, jm, gpu, with_w=True, order=order)
    b = H.make_case(kind, im, jm, ref, with_w=True, order=order)
    coef = np.asfortranarray(np.random.default_rng(11).uniform(0.5, 1.5, size=(im, jm)))
    blk = Block(a)
    blk.apply_bcs()
    jac, ia, ja = (t.cpu().numpy() for t in jacobian_coo(blk, coefdiag=coef))
    wb, _ = H.residual_sequence(ref, b, f"flux_num_dnc{order}_2d")
    jb, ib, jbb = H.jacobian_sequence(ref, b, wb, None, coef, scheme_d=f"flux_num_dnc{order}_2d_d")
    assert np.array_equal(ia, ib) and np.array_equal(ja, jbb)
    scale = np.abs(jb).max()
    assert np.abs(jac - jb).max() < 1e-12 * scale, np.abs(jac - jb).max() / scale
    flips = np.flatnonzero((np.abs(jac) > 2e-16) != (np.abs(jb) > 2e-16))
    assert flips.size == 0 or np.abs(jb[flips]).max() < 1e-15


def test_order_and_ghost_depth_must_agree(gpu):
    """a routine of one order handed arrays of another order's ghost depth is refused (the reference would read out of bounds)"""
    from broadcast_b200._lib import BroadcastB200Error
    c = H.make_case("bl", 30, 20, gpu, with_w=True, order=7)
    res = c.zeros_state()
    with pytest.raises(BroadcastB200Error):
        gpu["f_sch"].flux_num_dnc5_2d(res, c.w, *c.scheme_args())


@pytest.mark.parametrize("order", ORDERS)
def test_residual_orders_at_the_c1_grid(gpu, ref, order):
    """the other orders at the size of BASELINE.json's C1 (500 x 150 boundary layer): same parity bar as the order-5 configs"""
    a = H.make_case("bl", 500, 150, gpu, with_w=True, order=order)
    b = H.make_case("bl", 500, 150, ref, with_w=True, order=order)
    name = f"flux_num_dnc{order}_2d"
    wa, ra = H.residual_sequence(gpu, a, name)
    wb, rb = H.residual_sequence(ref, b, name)
    H.assert_residual_parity(ra, rb, b, wb, floor=H.fma_floor(b, name), what=name + " 500x150")


@pytest.mark.parametrize("order", [3, 7])
def test_c_abi_context_runs_the_other_orders(gpu, ref, order):
    """the resident C-ABI context (bcast_ctx_*: what a C / Fortran host binds) with gh = 2 / 4: state, boundary fills, residual
    and norms of the order-3 / 7 scheme; its block-Jacobian entry is order 5 and says so"""
    from broadcast_b200.cabi_ctx import Context
    from broadcast_b200._lib import BroadcastB200Error
    a = H.make_case("bl", 60, 31, gpu, with_w=True, order=order)
    b = H.make_case("bl", 60, 31, ref, with_w=True, order=order)
    name = f"flux_num_dnc{order}_2d"
    wb, rb = H.residual_sequence(ref, b, name)
    ctx = Context(a)
    ctx.upload_state(a.w)
    ra = ctx.residual()
    H.assert_residual_parity(ra, rb, b, wb, floor=H.fma_floor(b, name), what="context " + name)
    n2, ninf = ctx.norms()
    n2b, ninfb = ref["f_norm"].compute_norml2inf(rb, b.im, b.jm, b.gh)
    assert np.allclose(n2, n2b, rtol=1e-10) and np.allclose(ninf, ninfb, rtol=1e-10)
    with pytest.raises(BroadcastB200Error):
        ctx.jacobian_csr()
    ctx.close()
