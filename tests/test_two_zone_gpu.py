"""The two-zone colour loop of the O-mesh driver when im is not a multiple of the colour period (cylinder.py:896-939): per zone and
colour, f_misc.testvector_partial (misc/ComputeJacobian.f90:1075-1092) -> linearised fills (handleBC.applyBC mode 1) -> tangent ->
f_misc.computejacobianfromjv_relaxed_withjnandcheck(..., mini, zone) (misc/ComputeJacobian.f90:1095-1204), then remove_zero_jac and
the duplicate-summing CSR constructor -- the product's drop-in modules against oracle/_ref, entry point by entry point as the driver
calls them.  Also: testvector_partial on its own (bit-exact seeds for windows with offsets) and compute_norml2."""
import numpy as np
import pytest

import helpers as H
from broadcast_b200 import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("window", [(0, 20, 0, 18), (21, 44, 0, 18), (7, 30, 3, 12), (0, 44, 0, 18)])
def test_testvector_partial_bit_exact(gpu, ref, window):
    im, jm, gh = 45, 19, 3
    for m, l, k in [(0, 0, 0), (3, 5, 2), (4, 6, 6), (1, 2, 4)]:
        out = []
        for mods in (gpu, ref):
            wd = np.asfortranarray(np.random.default_rng(1).standard_normal((im + 2 * gh, jm + 2 * gh, 5)))   # the routine zeroes it
            mods["f_misc"].testvector_partial(wd, m, l, k, gh, im, jm, *window)
            out.append(wd)
        assert np.array_equal(out[0], out[1]), (window, m, l, k)
        assert out[1].sum() > 0


def _two_zone_loop(mods, c, w, coefdiag, colours):
    im, jm, gh = c.im, c.jm, c.gh
    s = 2 * gh + 1
    half = im // 2 // s * s                               # cylinder.py:897 (Python 2 integer division)
    zones = [(0, half - 1, 0, jm - 1), (half, im - 1, 0, jm - 1)]
    nb = im * jm * s * s * 25 * 2
    jac = np.zeros(nb)
    ia = np.zeros(nb, dtype=np.int32)
    ja = np.zeros(nb, dtype=np.int32)
    wd = c.zeros_state()
    res, resd = c.zeros_state(), c.zeros_state()
    f_misc, f_lin = mods["f_misc"], mods["f_lin"]
    for n, (i0, i1, j0, j1) in enumerate(zones):
        for (m, l, k) in colours:
            wd *= 0.0
            f_misc.testvector_partial(wd, m, l, k, gh, im, jm, i0, i1, j0, j1)
            ww = w.copy(order="F")
            cases.apply_bcs_lin(c, ww, wd, mods["f_bnd"], f_lin)
            f_lin.flux_num_dnc5_2d_d(res, resd, ww, wd, *c.scheme_args())
            f_misc.computejacobianfromjv_relaxed_withjnandcheck(jac, ia, ja, resd, m, l, k, gh, coefdiag, 2e-16, n)
    return jac, ia, ja


def test_two_zone_driver_sequence_matches_reference(gpu, ref):
    """im = 45 (45 % 7 = 3): zone 0 = columns 1 .. 21, zone 1 = columns 22 .. 45, every colour that can touch the cut and the zone
    boundary plus a spread of interior ones (the full 245-colour loop on both sides would take minutes through the host API)"""
    im, jm = 45, 19
    a = H.make_case("cyl", im, jm, gpu, with_w=True)
    b = H.make_case("cyl", im, jm, ref, with_w=True)
    wa, _ = H.residual_sequence(gpu, a)
    wb, _ = H.residual_sequence(ref, b)
    coef = np.asfortranarray(np.random.default_rng(4).uniform(0.5, 1.5, size=(im, jm)))
    s = 7
    colours = [(m, l, k) for m in (0, 2, 4) for l in range(s) for k in (0, 3, 6)]
    ja_, ia_, jja_ = _two_zone_loop(gpu, a, wa, coef, colours)
    jb_, ib_, jjb_ = _two_zone_loop(ref, b, wb, coef, colours)
    assert np.array_equal(ia_, ib_) and np.array_equal(jja_, jjb_)            # slot-exact integer lists (the check reads values > mini)
    scale = np.abs(jb_).max()
    assert np.abs(ja_ - jb_).max() < 1e-12 * scale
    A = H.coo_to_dict(ja_, ia_, jja_)
    B = H.coo_to_dict(jb_, ib_, jjb_)
    n = 5 * im * jm
    A.resize((n, n)); B.resize((n, n))
    D = (A - B).tocoo()
    assert A.nnz == B.nnz and (D.nnz == 0 or np.abs(D.data).max() < 1e-12 * np.abs(B.data).max())
    # every visited column of zone 0 and zone 1 is there: the columns seeded in a zone appear exactly once per (row, column) after the
    # filter (no double counting across the zones: the check skips slots that already hold a value)
    C = B.tocoo()
    assert len(set(zip(C.row.tolist(), C.col.tolist()))) == C.nnz


def test_compute_norml2(gpu, ref):
    """srcfv/norm.F90:2-32: L2 norm and mean per equation (tree reduction on the device, sequential sum in the Fortran)"""
    im, jm, gh = 70, 21, 3
    rhs = np.asfortranarray(np.random.default_rng(2).standard_normal((im + 2 * gh, jm + 2 * gh, 5)) * np.array([1.0, 1e3, 1e-3, 0.0, 7.0]))
    n1, m1 = gpu["f_norm"].compute_norml2(rhs, im, jm, gh)
    n2, m2 = ref["f_norm"].compute_norml2(rhs, im, jm, gh)
    assert np.allclose(n1, n2, rtol=1e-13, atol=0.0) and np.allclose(m1, m2, rtol=1e-11, atol=1e-300)


def test_csr_transpose_matches_scipy(gpu):
    """adjoint operator (SURVEY.md 8(f4) first step): device transpose of the assembled Jacobian's CSR == scipy's csr(A.T), bit for bit
    (values are moved, never added), on a BL case (CSR / vol as the drivers store it) and on two row blocks of it"""
    import scipy.sparse as sp
    import torch
    from broadcast_b200.resident import Block, jacobian_hybrid, csr_transpose
    c = H.make_case("bl", 66, 28, gpu, with_w=True)
    blk = Block(c)
    blk.apply_bcs()
    ip, idx, dat = jacobian_hybrid(blk).to_csr(divide_by_vol=True)
    n = 5 * c.im * c.jm
    A = sp.csr_matrix((dat.cpu().numpy(), idx.cpu().numpy(), ip.cpu().numpy()), shape=(n, n))
    AT = sp.csr_matrix(A.T)
    AT.sort_indices()
    tp, ti, td = csr_transpose(ip, idx, dat, n)
    assert np.array_equal(tp.cpu().numpy(), AT.indptr) and np.array_equal(ti.cpu().numpy(), AT.indices)
    assert np.array_equal(td.cpu().numpy(), AT.data)
    # a row block (the rows of one i-slab): the transposed block has n rows and the block's rows as columns
    r0, r1 = 5 * c.jm * 20, 5 * c.jm * 45
    B = A[r0:r1]
    bp = torch.from_numpy(B.indptr.astype(np.int64)).cuda()
    bi = torch.from_numpy(B.indices.astype(np.int32)).cuda()
    bd = torch.from_numpy(B.data).cuda()
    tp, ti, td = csr_transpose(bp, bi, bd, n, row0=r0)
    BT = sp.csr_matrix(B.T)          # (n, r1 - r0): local column numbers
    BT.sort_indices()
    assert np.array_equal(tp.cpu().numpy(), BT.indptr) and np.array_equal(ti.cpu().numpy() - r0, BT.indices)
    assert np.array_equal(td.cpu().numpy(), BT.data)
