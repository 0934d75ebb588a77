"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle (oracle/_ref) on the same
seeded inputs.  Tolerance: 1e-12 relative to the max magnitude of each equation plane (the bound
BASELINE.json's north_star states); integer outputs (colouring, IA, JA) bit-exact."""
import numpy as np
import pytest

import helpers as H
from broadcast_b200 import cases

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = [("bl", 60, 40), ("bl", 97, 33), ("cyl", 70, 40)]


@pytest.mark.parametrize("kind,im,jm", CASES)
def test_geometry_bit_exact(gpu, ref, kind, im, jm):
    a = H.make_case(kind, im, jm, gpu)
    b = H.make_case(kind, im, jm, ref)
    for name in ("x0", "y0", "nx", "ny", "xc", "yc", "vol", "volf"):
        x, y = getattr(a, name), getattr(b, name)
        assert np.array_equal(x, y), (name, np.abs(x - y).max(), np.argwhere(x != y)[:5])


@pytest.mark.parametrize("kind,im,jm", CASES)
def test_boundary_fill_and_residual(gpu, ref, kind, im, jm):
    a = H.make_case(kind, im, jm, gpu)
    b = H.make_case(kind, im, jm, ref)
    wa, ra = H.residual_sequence(gpu, a)
    wb, rb = H.residual_sequence(ref, b)
    assert np.all(H.rel_err(wa, wb) < TOL), H.rel_err(wa, wb)
    assert np.all(H.rel_err(ra, rb) < TOL), H.rel_err(ra, rb)
    # the generic (unfused) kernels give the same answer as the fused tile kernel
    import os
    os.environ['BROADCAST_B200_GENERIC'] = '1'
    try:
        _, rg = H.residual_sequence(gpu, a)
    finally:
        del os.environ['BROADCAST_B200_GENERIC']
    assert np.all(H.rel_err(rg, rb) < TOL), H.rel_err(rg, rb)
    # (the fused kernel evaluates re-associated face formulas, residual_fast.cuh: same bound, not tighter)
    assert np.all(H.rel_err(rg, ra) < TOL), H.rel_err(rg, ra)
    # ghosts of the residual are never written
    gh = a.gh
    assert np.all(ra[:gh] == 0) and np.all(ra[:, :gh] == 0)


@pytest.mark.parametrize("kind,im,jm", [("bl", 60, 40), ("bl", 97, 33), ("cyl", 70, 40), ("bl", 300, 70)])
def test_residual_kernel_variants_agree(gpu, ref, kind, im, jm):
    """default 32 x 9 tile kernel (0 = 4), reference-shaped pipeline (1), tile kernel + TMA/persistent (2; odd leading dimension: its
    LDG fallback), first-generation tile kernel (3) and the j-marching kernel (5): each within the parity bound of the oracle.
    Variants 0 and 2 run the same phase functions on the same operands and agree bit for bit; the marching kernel evaluates the
    same face formulas from another instruction stream: it agrees with 0 to a few ulp of the face fluxes (measured: bit for bit)."""
    import torch
    from broadcast_b200.resident import Block
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    wb, rb = H.residual_sequence(ref, b)
    floor = H.fma_floor(b)
    blk = Block(a)
    blk.apply_bcs()
    gh = a.gh
    outs, np_outs = {}, {}
    for v in (0, 1, 2, 3, 4, 5):
        r = blk.residual(variant=v).clone()
        outs[v] = r
        np_outs[v] = np.asfortranarray(r.cpu().numpy().transpose(2, 1, 0))     # (planes, j, i) image -> (i, j, planes)
        H.assert_residual_parity(np_outs[v], rb, b, wb, floor=floor, what=("variant", v))
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[0], outs[4])
    assert np.all(H.backward_err(np_outs[5], np_outs[0], b, wb) < 1e-14)


@pytest.mark.parametrize("kind,im,jm", [("bl", 7, 7), ("bl", 8, 8), ("bl", 33, 7), ("bl", 9, 10), ("cyl", 14, 9), ("bl", 3, 12), ("bl", 12, 5)])
def test_smallest_grids(gpu, ref, kind, im, jm):
    """grids of one tile or less (a stencil wide, fewer rows than the wall scheme's special rows + 1 fall back to the generic
    pipeline inside the same entry point): boundary fills, residual, one tangent direction vs the oracle"""
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    wa, ra = H.residual_sequence(gpu, a)
    wb, rb = H.residual_sequence(ref, b)
    assert np.all(H.rel_err(wa, wb) < TOL), H.rel_err(wa, wb)
    assert np.all(H.rel_err(ra, rb) < TOL), H.rel_err(ra, rb)
    wd = np.asfortranarray(np.random.default_rng(1).standard_normal(wa.shape))
    _, da = H.tangent_sequence(gpu, a, wa, wd)
    _, db = H.tangent_sequence(ref, b, wb, wd)
    assert np.all(H.rel_err(da, db) < TOL), H.rel_err(da, db)


@pytest.mark.parametrize("im,jm", [(200, 64), (97, 33), (64, 18), (40, 12)])
def test_residual_in_two_parts_equals_the_whole(gpu, im, jm):
    """inner tiles + ring of tiles (bcd_residual_part 1, 2) == one launch, bit for bit; the inner part must not depend on any
    ghost cell (poisoned while it runs); Block.step_overlapped == apply_bcs + residual"""
    import torch
    from broadcast_b200.resident import Block
    c = H.make_case("bl", im, jm, gpu, with_w=True)
    blk = Block(c)
    blk.apply_bcs()
    whole = blk.residual().clone()
    gh = c.gh
    w_ok = blk.w.clone()
    blk.res.zero_()
    blk.w[:, :gh] = float("nan"); blk.w[:, -gh:] = float("nan"); blk.w[:, :, :gh] = float("nan"); blk.w[:, :, -gh:] = float("nan")
    blk.residual_part(1)
    assert not torch.isnan(blk.res).any()
    blk.w.copy_(w_ok)
    blk.residual_part(2)
    assert torch.equal(blk.res, whole)
    blk.upload_state(c.w)
    blk.res.zero_()
    blk.step_overlapped()
    torch.cuda.synchronize()
    assert torch.equal(blk.res, whole)


def test_residual_full_size_properties(gpu):
    """C5 (8192 x 2048, BASELINE.json's bench configuration): the oracle cannot run there in seconds, so size-independent
    properties: (1) the fused kernels agree with the reference-shaped pipeline to TOL-level noise (the reference's own
    FMA / no-FMA builds differ by 1e-12 of the plane maximum at 126 x 60 already, profiles/r1_d_summary.md; bound 5e-12),
    (2) no NaN, ghost frame untouched, (3) the TMA tile variant equals the tile kernel bit for bit, (4) translation invariance of the
    tiling: the residual of an i-window of the grid computed as its own block equals the same cells of the full grid away from
    the window's edges bit for bit (each cell's result must not depend on which tile it falls in)."""
    import torch
    import broadcast_b200 as bb
    from broadcast_b200.resident import Block
    im, jm = 8192, 2048
    c = cases.make_bl_case(im, jm, f_geom=bb.f_geom)
    blk = Block(c)
    blk.apply_bcs()
    gh = c.gh
    r0 = blk.residual(variant=5).clone()     # marching kernel
    rg = blk.residual(variant=1).clone()
    r2 = blk.residual(variant=2).clone()
    r4 = blk.residual(variant=0).clone()     # tile kernel (default)
    assert not torch.isnan(r0).any()
    assert torch.equal(r4, r2)
    inner = (slice(None), slice(gh, -gh), slice(gh, -gh))
    scale = rg[inner].abs().amax(dim=(1, 2))
    for r in (r0, r4):
        err = (r[inner] - rg[inner]).abs().amax(dim=(1, 2))
        ok = (err <= 5e-12 * scale) | (scale == 0)
        assert bool(ok.all()), (err / scale).tolist()
    assert float(r0[:, :gh].abs().max()) == 0.0 and float(r0[:, :, :gh].abs().max()) == 0.0
    assert float(r0[:, -gh:].abs().max()) == 0.0 and float(r0[:, :, -gh:].abs().max()) == 0.0
    # tiling invariance: an i-window of the grid as an i-slab with two internal edges whose first column is not a multiple of 32
    # away from the full grid's tiles / strips.  Window (1 of 3) has an odd width (the marching kernel needs an even leading
    # dimension for its tensor maps and falls back to the tile kernel): tile kernel; window (2 of 5) has an even width: marching.
    from broadcast_b200 import sharding
    for (rank, world, variant, full) in ((1, 3, 0, r4), (2, 5, 5, r0)):
        case_w, desc = sharding.slab_of(c, rank, world)
        lo, hi = sharding.slab_range(im, rank, world)
        assert (lo - 1) % 32 != 0 and (variant != 5 or case_w.im % 2 == 0)
        wb = Block(case_w, slab=desc)
        wb.w.copy_(blk.w[:, :, lo - 1:hi + 2 * gh])      # the full block's state, boundary fills included
        rw = wb.residual(variant=variant)
        assert torch.equal(rw[:, gh:-gh, gh:-gh], full[:, gh:-gh, gh + lo - 1:gh + hi])
        del wb


@pytest.mark.parametrize("kind,im,jm", CASES[:2])
def test_residual_nowall(gpu, ref, kind, im, jm):
    a = H.make_case(kind, im, jm, gpu)
    b = H.make_case(kind, im, jm, ref)
    _, ra = H.residual_sequence(gpu, a, "flux_num_dnc5_nowall_2d")
    _, rb = H.residual_sequence(ref, b, "flux_num_dnc5_nowall_2d")
    assert np.all(H.rel_err(ra, rb) < TOL), H.rel_err(ra, rb)


def test_residual_with_spanwise_velocity(gpu, ref):
    a = H.make_case("bl", 60, 40, gpu, with_w=True)
    b = H.make_case("bl", 60, 40, ref, with_w=True)
    _, ra = H.residual_sequence(gpu, a)
    _, rb = H.residual_sequence(ref, b)
    assert np.abs(rb[..., 3]).max() > 0
    assert np.all(H.rel_err(ra, rb) < TOL), H.rel_err(ra, rb)


@pytest.mark.parametrize("kind,im,jm", CASES)
def test_tangent_random_direction(gpu, ref, kind, im, jm):
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    wa, _ = H.residual_sequence(gpu, a)
    wb, _ = H.residual_sequence(ref, b)
    rng = np.random.default_rng(1)
    wd = np.asfortranarray(rng.standard_normal(wa.shape))
    wda, rda = H.tangent_sequence(gpu, a, wa, wd)
    wdb, rdb = H.tangent_sequence(ref, b, wb, wd)
    assert np.all(H.rel_err(wda, wdb) < 1e-13), H.rel_err(wda, wdb)
    assert np.all(H.rel_err(rda, rdb) < TOL), H.rel_err(rda, rdb)


@pytest.mark.parametrize("kind,im,jm", [("bl", 60, 40), ("cyl", 70, 40)])
def test_colour_loop_coo(gpu, ref, kind, im, jm):
    """a sample of colours through the full drop-in sequence: IA/JA bit-exact, values to TOL"""
    a = H.make_case(kind, im, jm, gpu)
    b = H.make_case(kind, im, jm, ref)
    wa, _ = H.residual_sequence(gpu, a)
    wb, _ = H.residual_sequence(ref, b)
    colours = [(0, 0, 0), (1, 3, 2), (4, 6, 6), (2, 5, 0), (3, 0, 4), (4, 4, 1)]
    rng = np.random.default_rng(2)
    coef = np.asfortranarray(rng.uniform(0.5, 1.5, size=(im, jm)))
    ja_, ia_a, ja_a = H.jacobian_sequence(gpu, a, wa, colours, coef)
    jb_, ia_b, ja_b = H.jacobian_sequence(ref, b, wb, colours, coef)
    assert np.array_equal(ia_a, ia_b)
    assert np.array_equal(ja_a, ja_b)
    scale = np.abs(jb_).max()
    assert np.abs(ja_ - jb_).max() < TOL * scale


def test_scatter_variants_integer_exact(gpu, ref):
    im, jm, gh = 35, 23, 3
    s = 2 * gh + 1
    rng = np.random.default_rng(3)
    resd = np.asfortranarray(rng.standard_normal((im + 2 * gh, jm + 2 * gh, 5)))
    coef = np.asfortranarray(rng.uniform(0.5, 1.5, size=(im, jm)))
    nb = 25 * s * s * im * jm
    for name, extra in (("computejacobianfromjv", (im, jm)), ("computejacobianfromjv_relaxed", (coef,)),
                        ("computejacobianfromjv_relaxed_withjn", (coef,)), ("computejacobianfromjv_withjn", (im, jm)),
                        ("computejacobianfromdz", (im, jm))):
        out = []
        for mods in (gpu, ref):
            jac, ia, ja = np.zeros(nb), np.zeros(nb, np.int32), np.zeros(nb, np.int32)
            for (m, l, k) in [(0, 0, 0), (2, 3, 3), (4, 6, 5), (1, 4, 0), (3, 0, 6), (4, 6, 6)]:
                getattr(mods["f_misc"], name)(jac, ia, ja, resd, m, l, k, gh, *extra)
            out.append((jac, ia, ja))
        assert np.array_equal(out[0][1], out[1][1]), name
        assert np.array_equal(out[0][2], out[1][2]), name
        assert np.array_equal(out[0][0], out[1][0]), name


def test_colour_sharded_device_loop_merges_to_the_full_jacobian(gpu):
    """colour sharding (bcd_colour_range): three 'ranks' on one device each run their range of the 49 passes of the device colour
    loop on the i-periodic O-mesh; the merged COO lists equal the unsharded loop exactly, slot by slot"""
    import torch
    from broadcast_b200 import sharding
    from broadcast_b200.resident import Block, jacobian_coo
    c = H.make_case("cyl", 42, 30, gpu, with_w=True)
    blk = Block(c)
    blk.apply_bcs()
    full = [t.clone() for t in jacobian_coo(blk)]
    s = 2 * c.gh + 1
    world = 3
    acc = None
    seen = torch.zeros_like(full[0], dtype=torch.bool)
    for r in range(world):
        c0, c1 = sharding.colour_range(s * s, r, world)
        jac, ia, ja = jacobian_coo(blk, colours=(c0, c1))
        mine = (ia != 0) | (ja != 0) | (jac != 0)
        assert not bool((mine & seen).any())                  # disjoint slots
        seen |= mine
        acc = [jac.clone(), ia.clone(), ja.clone()] if acc is None else [acc[0] + jac, acc[1] + ia, acc[2] + ja]
    assert torch.equal(acc[0], full[0]) and torch.equal(acc[1], full[1]) and torch.equal(acc[2], full[2])
    n = 5 * c.im * c.jm
    parts = []
    for r in range(world):
        c0, c1 = sharding.colour_range(s * s, r, world)
        parts.append(tuple(t.cpu().numpy() for t in jacobian_coo(blk, colours=(c0, c1))))
    A = sharding.merge_colour_shards(parts, n)
    B = H.coo_to_dict(*(t.cpu().numpy() for t in full))
    B.resize((n, n))
    assert A.nnz == B.nnz and (A - B).nnz == 0


def test_two_zone_scatter_with_check_integer_exact(gpu, ref):
    """computejacobianfromjv_relaxed_withjnandcheck (misc/ComputeJacobian.f90:1095-1204, cylinder.py:1159): read-modify-write
    of the slot arrays, zone 0 then zone 1 over the same colours, partially pre-filled slots"""
    im, jm, gh = 37, 24, 3
    s = 2 * gh + 1
    rng = np.random.default_rng(5)
    coef = np.asfortranarray(rng.uniform(0.5, 1.5, size=(im, jm)))
    nb = 2 * 25 * s * s * im * jm
    colours = [(0, 0, 0), (2, 3, 3), (4, 6, 5), (1, 4, 0), (3, 0, 6), (4, 6, 6), (2, 5, 2)]
    resds = [np.asfortranarray(rng.standard_normal((im + 2 * gh, jm + 2 * gh, 5))) for _ in range(2 * len(colours))]
    pre = rng.standard_normal(nb) * (rng.uniform(size=nb) < 0.3)          # 30 % of the slots already taken
    out = []
    for mods in (gpu, ref):
        jac, ia, ja = pre.copy(), np.full(nb, 7, np.int32), np.full(nb, 11, np.int32)
        q = 0
        for n in (0, 1):
            for (m, l, k) in colours:
                mods["f_misc"].computejacobianfromjv_relaxed_withjnandcheck(jac, ia, ja, resds[q], m, l, k, gh, coef, 1e-14, n)
                q += 1
        out.append((jac, ia, ja))
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2])
    assert np.array_equal(out[0][0], out[1][0])
    assert np.count_nonzero(out[0][0] != pre) > 0


def test_testvector(gpu, ref):
    im, jm, gh = 30, 22, 3
    for (m, l, k) in [(0, 0, 0), (4, 6, 6), (2, 3, 1)]:
        a = np.asfortranarray(np.ones((im + 2 * gh, jm + 2 * gh, 5)))
        b = a.copy(order="F")
        gpu["f_misc"].testvector(a, m, l, k, gh, im, jm)
        ref["f_misc"].testvector(b, m, l, k, gh, im, jm)
        assert np.array_equal(a, b)
        assert a.sum() == len(range(l + 1, im + 1, 7)) * len(range(k + 1, jm + 1, 7))


def test_norms(gpu, ref):
    a = H.make_case("bl", 60, 40, gpu)
    _, ra = H.residual_sequence(gpu, a)
    na, ia = gpu["f_norm"].compute_norml2inf(ra, a.im, a.jm, a.gh)
    nb, ib = ref["f_norm"].compute_norml2inf(ra, a.im, a.jm, a.gh)
    assert np.allclose(na, nb, rtol=1e-13, atol=0)
    assert np.allclose(ia, ib, rtol=1e-13, atol=0)


@pytest.mark.parametrize("kind,im,jm", [("bl", 40, 30), ("cyl", 42, 30)])
def test_full_jacobian_device_colour_loop(gpu, ref, kind, im, jm):
    """whole Jacobian: device-resident colour loop (5 directions per pass) vs the reference's 245-colour
    host loop on the oracle: IA/JA bit-exact in the reference's own slot order, values to TOL, and the
    filtered sparsity pattern (|v| > 2e-16) identical."""
    import torch
    from broadcast_b200.resident import Block, jacobian_coo
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    rng = np.random.default_rng(5)
    coef = np.asfortranarray(rng.uniform(0.5, 1.5, size=(im, jm)))
    blk = Block(a)
    blk.apply_bcs()
    jac, ia, ja = jacobian_coo(blk, coefdiag=coef)
    jac, ia, ja = jac.cpu().numpy(), ia.cpu().numpy(), ja.cpu().numpy()
    wb, _ = H.residual_sequence(ref, b)
    jb, ib, jbb = H.jacobian_sequence(ref, b, wb, None, coef)
    assert np.array_equal(ia, ib)
    assert np.array_equal(ja, jbb)
    scale = np.abs(jb).max()
    assert np.abs(jac - jb).max() < TOL * scale, np.abs(jac - jb).max() / scale
    ka, kb = np.abs(jac) > 2e-16, np.abs(jb) > 2e-16
    flips = np.flatnonzero(ka != kb)
    # entries that flip across the 2e-16 filter must themselves be at rounding level
    assert flips.size == 0 or np.abs(jb[flips]).max() < 1e-15, (flips.size, np.abs(jb[flips]).max())
    # blockwise relative accuracy of the significant entries
    big = np.abs(jb) > 1e-6 * scale
    assert np.max(np.abs(jac[big] - jb[big]) / np.abs(jb[big])) < 1e-9


@pytest.mark.parametrize("kind,im,jm", [("bl", 40, 30), ("cyl", 42, 30)])
def test_hybrid_jacobian_matches_reference_loop(gpu, ref, kind, im, jm):
    """direct block kernels (interior) + strip colour loop == the reference's 245-colour loop, compared
    as canonical CSR after the reference's 2e-16 filter: identical pattern, values to TOL."""
    import scipy.sparse as sp
    from broadcast_b200.resident import Block, jacobian_hybrid
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    rng = np.random.default_rng(5)
    coef = np.asfortranarray(rng.uniform(0.5, 1.5, size=(im, jm)))
    blk = Block(a)
    blk.apply_bcs()
    A = jacobian_hybrid(blk, coefdiag=coef).to_scipy_csr()
    wb, _ = H.residual_sequence(ref, b)
    jb, ib, jbb = H.jacobian_sequence(ref, b, wb, None, coef)
    keep = np.abs(jb) > 2e-16
    n = 5 * im * jm
    B = sp.csr_matrix((jb[keep], (ib[keep], jbb[keep])), shape=(n, n))
    A.sort_indices(); B.sort_indices()
    D = (A - B).tocoo()
    scale = np.abs(B.data).max()
    assert np.abs(D.data).max() < TOL * scale, np.abs(D.data).max() / scale
    # pattern: entries present in one and absent in the other must be at rounding level
    PA = sp.csr_matrix((np.ones_like(A.data), A.indices, A.indptr), shape=A.shape)
    PB = sp.csr_matrix((np.ones_like(B.data), B.indices, B.indptr), shape=B.shape)
    X = (PA - PB).tocoo()
    odd = X.data != 0
    if odd.any():
        vals = np.abs(np.asarray((A + B)[X.row[odd], X.col[odd]])).ravel()
        assert vals.max() < 1e-15, (odd.sum(), vals.max())


def _dz_args(c):
    a = c.scheme_args()
    return a[:18] + a[20:]   # coeffs_5p_dz takes no k2, k4 (BROADCAST_npz.py:1242-1243)


@pytest.mark.parametrize("kind,im,jm", CASES)
def test_dz_operator_rows(gpu, ref, kind, im, jm):
    """f_dz.coeffs_5p_dz / coeffs_5p_dz2 on a seeded random direction and on a colour seed, vs the reference"""
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    wa, _ = H.residual_sequence(gpu, a)
    wb, _ = H.residual_sequence(ref, b)
    rng = np.random.default_rng(8)
    wd = np.asfortranarray(rng.standard_normal(wa.shape))
    wds = a.zeros_state()
    gpu["f_misc"].testvector(wds, 2, 3, 1, a.gh, im, jm)
    for d in (wd, wds):
        for name in ("coeffs_5p_dz", "coeffs_5p_dz2"):
            za = np.asfortranarray(np.full(wa.shape, 7.0))
            zb = za.copy(order="F")
            getattr(gpu["f_dz"], name)(za, wa, d, *_dz_args(a))
            getattr(ref["f_dz"], name)(zb, wb, d, *_dz_args(b))
            assert np.all(za[:a.gh] == 7.0) and np.all(za[:, :a.gh] == 7.0)   # ghosts untouched
            assert np.all(H.rel_err(za, zb) < TOL), (name, H.rel_err(za, zb))


def test_dz_colour_loop_device(gpu, ref):
    """device colour loop of the spanwise operators vs the reference loop of BROADCAST_npz.py:1231-1246"""
    from broadcast_b200.resident import Block, dz_coo
    im, jm = 30, 22
    a = H.make_case("bl", im, jm, gpu, with_w=True)
    b = H.make_case("bl", im, jm, ref, with_w=True)
    blk = Block(a)
    blk.apply_bcs()
    out = dz_coo(blk)
    wb, _ = H.residual_sequence(ref, b)
    gh = b.gh
    s = 2 * gh + 1
    nb = 25 * s * s * im * jm
    J1, I1, K1 = np.zeros(nb), np.zeros(nb, np.int32), np.zeros(nb, np.int32)
    J2, I2, K2 = np.zeros(nb), np.zeros(nb, np.int32), np.zeros(nb, np.int32)
    wd = b.zeros_state()
    dz, dz2 = b.zeros_state(), b.zeros_state()
    for m in range(5):
        for l in range(s):
            for k in range(s):
                wd *= 0.0
                ref["f_misc"].testvector(wd, m, l, k, gh, im, jm)
                ww = wb.copy(order="F")
                cases.apply_bcs_lin(b, ww, wd, ref["f_bnd"], ref["f_lin"])
                ref["f_dz"].coeffs_5p_dz(dz, ww, wd, *_dz_args(b))
                ref["f_dz"].coeffs_5p_dz2(dz2, ww, wd, *_dz_args(b))
                ref["f_misc"].computejacobianfromdz(J1, I1, K1, dz, m, l, k, gh, im, jm)
                ref["f_misc"].computejacobianfromdz(J2, I2, K2, dz2, m, l, k, gh, im, jm)
    for wh, (J, I, K) in ((1, (J1, I1, K1)), (2, (J2, I2, K2))):
        jac, ia, ja = (t.cpu().numpy() for t in out[wh])
        assert np.array_equal(ia, I) and np.array_equal(ja, K)
        assert np.abs(jac - J).max() < TOL * np.abs(J).max()


@pytest.mark.parametrize("kind,im,jm", [("bl", 40, 30), ("cyl", 42, 30)])
def test_face_linearisation_matches_direct_ad_blocks(gpu, kind, im, jm):
    """interior blocks: semi-analytic face linearisation (facejac.cuh) == direct forward-AD kernels (jac_blocks.cuh)"""
    import torch
    from broadcast_b200.resident import Block, jacobian_hybrid
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    blk = Block(a)
    blk.apply_bcs()
    coef = np.asfortranarray(np.random.default_rng(5).uniform(0.5, 1.5, size=(im, jm)))
    A = jacobian_hybrid(blk, coefdiag=coef, interior="faces")
    B = jacobian_hybrid(blk, coefdiag=coef, interior="ad")
    gh = a.gh
    x = A.blocks[:, :, :, gh:jm - gh, gh:im - gh]
    y = B.blocks[:, :, :, gh:jm - gh, gh:im - gh]
    scale = y.abs().max().item()
    assert (x - y).abs().max().item() < TOL * scale
    # structurally empty entries agree exactly
    assert torch.equal(x == 0, y == 0) or ((x == 0) != (y == 0)).sum().item() < 1e-3 * x.numel()


def test_set_bndbl(gpu, ref):
    a = H.make_case("bl", 30, 22, gpu)
    f1, w1 = np.zeros((22, 3, 5), order="F"), np.zeros((33, 5), order="F")
    f2, w2 = f1.copy(order="F"), w1.copy(order="F")
    gpu["f_init"].set_bndbl_2d(a.w, f1, w1, 30)
    ref["f_init"].set_bndbl_2d(a.w, f2, w2, 30)
    assert np.array_equal(f1, f2) and np.array_equal(w1, w2)


@pytest.mark.parametrize("world", [2, 3])
def test_slab_sharded_residual_and_jacobian_match_single_block(gpu, world):
    """i-slabs on one device: halo exchange + per-slab boundary fill, residual and Jacobian row blocks
    in GLOBAL numbering == the single-block result (gradients at slab-internal edges computed, not extrapolated)."""
    import torch
    from broadcast_b200 import sharding
    from broadcast_b200.resident import Block, jacobian_hybrid, local_halo_exchange
    im, jm = 66, 28
    g = H.make_case("bl", im, jm, gpu, with_w=True)
    coef = np.asfortranarray(np.random.default_rng(9).uniform(0.5, 1.5, size=(im, jm)))
    G = Block(g)
    G.apply_bcs()
    resG = G.residual().clone()
    JG = jacobian_hybrid(G, coefdiag=coef).to_scipy_csr()
    n2G, ninfG = G.norms()
    blocks, sums = [], torch.zeros(16, dtype=torch.float64, device="cuda")
    for r in range(world):
        sl, desc = sharding.slab_of(g, r, world)
        b = Block(sl, slab=desc)
        if desc[2] & 1:
            b.w[:, :, :g.gh] = float("nan")     # the halo must come from the exchange
        if desc[2] & 2:
            b.w[:, :, -g.gh:] = float("nan")
        blocks.append(b)
    local_halo_exchange(blocks)
    gh = g.gh
    rows, cols, vals = [], [], []
    for r, b in enumerate(blocks):
        b.apply_bcs()
        res = b.residual()
        lo, hi = sharding.slab_range(im, r, world)
        own = res[:, gh:gh + jm, gh:gh + b.im]
        ref = resG[:, gh:gh + jm, gh + lo - 1:gh + hi]
        assert (own - ref).abs().max().item() <= 1e-13 * ref.abs().max().item()
        cd = np.asfortranarray(coef[lo - 1:hi])
        v, ri, ci = jacobian_hybrid(b, coefdiag=cd).to_coo()
        assert ri.min().item() >= 5 * jm * (lo - 1) and ri.max().item() < 5 * jm * hi     # a contiguous row block
        rows.append(ri.cpu().numpy()); cols.append(ci.cpu().numpy()); vals.append(v.cpu().numpy())
        b.norms()
        sums += b.out10
    import scipy.sparse as sp
    n = 5 * im * jm
    JS = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    D = (JS - JG).tocoo()
    assert D.nnz == 0 or np.abs(D.data).max() < TOL * np.abs(JG.data).max()
    assert JS.nnz == JG.nnz
    h = sums.cpu().numpy()
    assert np.allclose(np.sqrt(h[:5]), n2G, rtol=1e-12) and np.allclose(h[5:10] ** 0.1, ninfG, rtol=1e-12)


def test_streamed_host_step_matches_resident_block(gpu):
    """slab-pipelined host step (H2D / compute / D2H overlapped over i-slabs) == the single-block host step"""
    import torch
    from broadcast_b200.resident import Block, StreamedBlock
    g = H.make_case("bl", 130, 40, gpu, with_w=True)
    blk = Block(g)
    wp = torch.empty(blk.w.shape, dtype=torch.float64).pin_memory()
    wp.copy_(blk.w.cpu())
    r1 = torch.zeros(blk.w.shape, dtype=torch.float64).pin_memory()
    r2 = torch.zeros(blk.w.shape, dtype=torch.float64).pin_memory()
    blk.step_from_host(wp, r1)
    sb = StreamedBlock(g, nslab=4)
    sb.step_from_host(wp, r2)
    gh = g.gh
    a, b = r1[:, gh:-gh, gh:-gh], r2[:, gh:-gh, gh:-gh]
    assert (a - b).abs().max().item() <= 1e-13 * a.abs().max().item()
    assert a.abs().max().item() > 0
    # tapered slabs (short first upload / last download): same residual as the even split, bit for bit
    bounds = [0, 10, 30, 70, 110, 122, 130]
    r3 = torch.zeros(blk.w.shape, dtype=torch.float64).pin_memory()
    StreamedBlock(g, nslab=6, bounds=bounds).step_from_host(wp, r3)
    assert torch.equal(r3[:, gh:-gh, gh:-gh], b)
    tb = StreamedBlock.tapered_bounds(8192, 8)
    assert tb[0] == 0 and tb[-1] == 8192 and all(w % 32 == 0 for w in np.diff(tb)) and tb[1] < 8192 // 8


@pytest.mark.parametrize("im,jm,nwin,margin,taper", [(130, 120, 3, 2, False), (64, 97, 4, 2, False), (70, 40, 1, 2, False),
                                                     (96, 200, 6, 2, True), (64, 128, 8, 5, True), (40, 300, 9, 3, False)])
def test_row_streamed_host_step_is_bit_identical(gpu, im, jm, nwin, margin, taper):
    """host step pipelined over ROW windows (contiguous host-link copies, windows computed with a margin of gh + 2 rows as blocks of
    their own: wall rows only in the first, top fill only in the last, clipped side fills) == the single-block host step, bit for bit"""
    import torch
    from broadcast_b200.resident import Block, RowStreamedBlock
    g = H.make_case("bl", im, jm, gpu, with_w=True)
    blk = Block(g)
    wp = torch.empty(blk.w.shape, dtype=torch.float64).pin_memory()
    wp.copy_(blk.w.cpu())
    r1 = torch.zeros(blk.w.shape, dtype=torch.float64).pin_memory()
    r2 = torch.full(blk.w.shape, 7.0, dtype=torch.float64).pin_memory()
    blk.step_from_host(wp, r1)
    rb = RowStreamedBlock(g, nwin=nwin, margin=margin, taper=taper)
    rb.step_from_host(wp, r2)
    gh = g.gh
    assert torch.equal(r1[:, gh:-gh, gh:-gh], r2[:, gh:-gh, gh:-gh])
    assert r1[:, gh:-gh, gh:-gh].abs().max().item() > 0
    assert bool((r2[:, :gh] == 7.0).all()) and bool((r2[:, -gh:] == 7.0).all())     # ghost rows of the host array untouched
    rb.step_from_host(wp, r2)                                                         # a second step on the same object
    assert torch.equal(r1[:, gh:-gh, gh:-gh], r2[:, gh:-gh, gh:-gh])


def test_csr_row_blocks_and_petsc_file(gpu, tmp_path):
    """device-side filter + divide-by-volume + CSR == the reference's remove_zero_jac / Jacvol loop / csr_matrix on the
    hybrid Jacobian; slab row blocks gather to the same matrix; PETSc binary AIJ round trip"""
    import scipy.sparse as sp
    from broadcast_b200 import sharding, formats
    from broadcast_b200.resident import Block, jacobian_hybrid, local_halo_exchange
    im, jm = 48, 26
    g = H.make_case("bl", im, jm, gpu, with_w=True)
    G = Block(g)
    G.apply_bcs()
    Hj = jacobian_hybrid(G)
    v, r, c = (t.cpu().numpy() for t in Hj.to_coo())
    gh = g.gh
    jacvol = v / g.vol[r // (5 * jm) + gh, (r % (5 * jm)) // 5 + gh]          # BROADCAST_npz.py:1206-1209
    n = 5 * im * jm
    A = sp.csr_matrix((jacvol, (r, c)), shape=(n, n)); A.sort_indices()
    ip, idx, dat = (t.cpu().numpy() for t in Hj.to_csr(divide_by_vol=True))
    assert np.array_equal(ip, A.indptr) and np.array_equal(idx, A.indices)
    assert np.allclose(dat, A.data, rtol=1e-15, atol=0)
    ipt, idxt, datt = (t.cpu().numpy() for t in Hj.to_csr_torch(divide_by_vol=True))      # torch-op cross-check of the kernels
    assert np.array_equal(ipt, ip) and np.array_equal(idxt, idx) and np.array_equal(datt, dat)
    # two slabs -> two row blocks -> host gather
    blocks = []
    for rk in range(2):
        sl, desc = sharding.slab_of(g, rk, 2)
        blocks.append(Block(sl, slab=desc))
    local_halo_exchange(blocks)
    parts = []
    for b in blocks:
        b.apply_bcs()
        parts.append(tuple(t.cpu().numpy() for t in jacobian_hybrid(b).to_csr(divide_by_vol=True)))
    ip2, idx2, dat2 = sharding.gather_row_blocks(parts)
    assert np.array_equal(ip2, A.indptr) and np.array_equal(idx2, A.indices)
    assert np.allclose(dat2, A.data, rtol=1e-14, atol=0)
    p = str(tmp_path / "Jacsurvol")
    formats.write_petsc_aij(p, ip2, idx2, dat2, n)
    ip3, idx3, dat3, shape = formats.read_petsc_aij(p)
    assert shape == (n, n) and np.array_equal(ip3, A.indptr) and np.array_equal(dat3.real, dat2)


@pytest.mark.parametrize("kind,im,jm", [("bl", 48, 26), ("cyl", 42, 30), ("bl", 97, 33)])
def test_csr_kernels_match_scipy_semantics(gpu, kind, im, jm):
    """hand-written CSR conversion (csrc/csr.cu: counts, warp-shuffle scan, ballot-compacted fill, warp rank sort of the strip
    rows) == remove_zero_jac + csr_matrix((Jac,(IA,JA))) (+ the divide-by-volume loop) of the reference drivers: indptr, column
    order and values bit for bit; also on the two row blocks of an i-slab split"""
    import scipy.sparse as sp
    from broadcast_b200 import sharding
    from broadcast_b200.resident import Block, jacobian_hybrid, local_halo_exchange
    g = H.make_case(kind, im, jm, gpu, with_w=True)
    G = Block(g)
    G.apply_bcs()
    coef = np.asfortranarray(np.random.default_rng(2).uniform(0.5, 1.5, size=(im, jm)))
    Hj = jacobian_hybrid(G, coefdiag=coef)
    n = 5 * im * jm
    gh = g.gh
    v, r, c = (t.cpu().numpy() for t in Hj.to_coo())
    for dbv in (False, True):
        vv = v / g.vol[r // (5 * jm) + gh, (r % (5 * jm)) // 5 + gh] if dbv else v
        A = sp.csr_matrix((vv, (r, c)), shape=(n, n))
        A.sort_indices()
        ip, idx, dat = (t.cpu().numpy() for t in Hj.to_csr_device(divide_by_vol=dbv))
        assert ip[-1] == A.nnz == len(v)                      # no duplicate (row, column) pair in the reference's slot scheme
        assert np.array_equal(ip, A.indptr) and np.array_equal(idx, A.indices)
        assert np.array_equal(dat, A.data)
    if kind == "bl":
        blocks = []
        for rk in range(2):
            sl, desc = sharding.slab_of(g, rk, 2)
            blocks.append(Block(sl, slab=desc))
        local_halo_exchange(blocks)
        parts = []
        for rk, b in enumerate(blocks):
            b.apply_bcs()
            lo, hi = sharding.slab_range(im, rk, 2)
            parts.append(tuple(t.cpu().numpy() for t in jacobian_hybrid(b, coefdiag=np.asfortranarray(coef[lo - 1:hi])).to_csr_device()))
        ip2, idx2, dat2 = sharding.gather_row_blocks(parts)
        A = sp.csr_matrix((v, (r, c)), shape=(n, n))
        A.sort_indices()
        assert np.array_equal(ip2, A.indptr) and np.array_equal(idx2, A.indices)
        assert np.abs(dat2 - A.data).max() <= 1e-12 * np.abs(A.data).max()


# ---------------------------------------------------------------------------------------------------------------------
# tangent of the spanwise operator rows w.r.t. the base flow (f_lindz of BROADCAST_npz_sens.py:1768-1797)
# ---------------------------------------------------------------------------------------------------------------------
def _dz_tangent_dirs(mods, c, w, seed):
    rng = np.random.default_rng(seed)
    wd = np.asfortranarray(rng.standard_normal(w.shape))
    wd0 = np.asfortranarray(rng.standard_normal(w.shape) * np.abs(w).max(axis=(0, 1)))
    ws = c.zeros_state()
    mods["f_misc"].testvector(ws, 2, 3, 1, c.gh, c.im, c.jm)
    return wd, wd0, ws


@pytest.mark.parametrize("kind,im,jm", CASES + [("bl", 5, 3)])
def test_dz_tangent_rows(gpu, ref, kind, im, jm):
    """f_lindz.coeffs_5p_dz_d / coeffs_5p_dz2_d vs the reference's Tapenade code (srcfv/tangentdz) on dense directions and a
    colour seed; dz_out untouched, the whole of dz_outd written (ghost frame = 0)"""
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    wa, _ = H.residual_sequence(gpu, a)
    wb, _ = H.residual_sequence(ref, b)
    wd, wd0, ws = _dz_tangent_dirs(ref, b, wb, 12)
    g = a.gh
    for d0, d in ((wd0, wd), (ws, wd), (wd0, ws)):
        for name in ("coeffs_5p_dz_d", "coeffs_5p_dz2_d"):
            za, zb = (np.asfortranarray(np.full(wa.shape, 7.0)) for _ in range(2))
            zda, zdb = (np.asfortranarray(np.full(wa.shape, 5.0)) for _ in range(2))
            getattr(gpu["f_lindz"], name)(za, zda, wa, d0, d, *_dz_args(a))
            getattr(ref["f_lindz"], name)(zb, zdb, wb, d0, d, *_dz_args(b))
            assert np.all(za == 7.0) and np.all(zb == 7.0)
            assert np.all(zda[:g] == 0.0) and np.all(zda[:, :g] == 0.0) and np.all(zda[-g:] == 0.0) and np.all(zda[:, -g:] == 0.0)
            assert np.abs(zdb).max() > 0
            assert np.all(H.rel_err(zda, zdb) < TOL), (name, H.rel_err(zda, zdb))


def test_dz_tangent_device_fused_pass(gpu):
    """Block.dz_tangent: both operators in ONE pass equal the two single-operator passes bit for bit; a rectangle writes only its
    cells and its values do not depend on where the tiles start"""
    import torch
    from broadcast_b200.resident import Block, _t
    im, jm = 97, 33
    a = H.make_case("bl", im, jm, gpu, with_w=True)
    blk = Block(a)
    blk.apply_bcs()
    w = np.asfortranarray(blk.w.cpu().numpy().T)
    wd, wd0, _ = _dz_tangent_dirs(gpu, a, w, 5)
    dwd, dwd0 = _t(wd, blk.device), _t(wd0, blk.device)
    o1, o2, s1, s2 = (torch.zeros_like(blk.w) for _ in range(4))
    blk.dz_tangent(dwd0, dwd, o1, o2)
    blk.dz_tangent(dwd0, dwd, s1, None)
    blk.dz_tangent(dwd0, dwd, None, s2)
    assert torch.equal(o1, s1) and torch.equal(o2, s2) and float(o1.abs().max()) > 0 and float(o2.abs().max()) > 0
    gh = a.gh
    r1, r2 = (torch.full_like(blk.w, 3.0) for _ in range(2))
    rect = (12, 70, 4, 29)
    blk.dz_tangent(dwd0, dwd, r1, r2, rect=rect)
    sl = (slice(None), slice(gh + rect[2] - 1, gh + rect[3]), slice(gh + rect[0] - 1, gh + rect[1]))
    assert torch.equal(r1[sl], o1[sl]) and torch.equal(r2[sl], o2[sl])
    r1[sl] = 3.0
    r2[sl] = 3.0
    assert bool((r1 == 3.0).all()) and bool((r2 == 3.0).all())


def test_dz_tangent_full_size_properties(gpu):
    """C5 (8192 x 2048): size-independent properties of the fused pass -- no NaN, ghost frame untouched, exact homogeneity in both
    directions (powers of two), additivity in the base-flow variation to rounding, tiling invariance on a shifted rectangle --
    and its device time against the algorithmic traffic (240 B per cell for both operators)."""
    import json
    import os
    import torch
    import broadcast_b200 as bb
    from broadcast_b200.resident import Block
    im, jm = 8192, 2048
    c = cases.make_bl_case(im, jm, f_geom=bb.f_geom, with_w=True)
    blk = Block(c)
    blk.apply_bcs()
    gh = c.gh
    gen = torch.Generator(device=blk.device).manual_seed(4)
    wd = torch.randn(blk.w.shape, dtype=torch.float64, device=blk.device, generator=gen)
    wd0 = torch.randn(blk.w.shape, dtype=torch.float64, device=blk.device, generator=gen) * blk.w.abs().amax(dim=(1, 2), keepdim=True)
    wd1 = torch.randn(blk.w.shape, dtype=torch.float64, device=blk.device, generator=gen) * blk.w.abs().amax(dim=(1, 2), keepdim=True)
    o1, o2, p1, p2 = (torch.zeros_like(blk.w) for _ in range(4))
    blk.dz_tangent(wd0, wd, o1, o2)
    assert not torch.isnan(o1).any() and not torch.isnan(o2).any()
    assert float(o1[:, :gh].abs().max()) == 0.0 and float(o1[:, :, :gh].abs().max()) == 0.0
    blk.dz_tangent(wd0 * 4.0, wd * 0.5, p1, p2)
    assert torch.equal(p1, o1 * 2.0) and torch.equal(p2, o2 * 2.0)
    blk.dz_tangent(wd0 + wd1, wd, p1, p2)
    q1, q2 = (torch.zeros_like(blk.w) for _ in range(2))
    blk.dz_tangent(wd1, wd, q1, q2)
    for s, a, b in ((p1, o1, q1), (p2, o2, q2)):
        scale = s.abs().amax(dim=(1, 2))
        err = (s - a - b).abs().amax(dim=(1, 2))
        assert bool(((err <= 1e-12 * scale) | (scale == 0)).all()), (err / scale).tolist()
    rect = (12, 5000, 3, 1999)
    blk.dz_tangent(wd0, wd, q1, q2, rect=rect)
    sl = (slice(None), slice(gh + rect[2] - 1, gh + rect[3]), slice(gh + rect[0] - 1, gh + rect[1]))
    assert torch.equal(q1[sl], o1[sl]) and torch.equal(q2[sl], o2[sl])
    # device time (CUDA events on the launching stream, after warm-up)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        blk.dz_tangent(wd0, wd, o1, o2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    line = {"kernel": "k_dz_tangent", "grid": [im, jm], "ms": ms, "algorithmic_bytes_per_cell": 240, "GBps": 240.0 * im * jm / ms / 1e6}
    print("DZ_TANGENT", json.dumps(line))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/dz_tangent_c5.json", "w") as fh:
        fh.write(json.dumps(line) + "\n")


@pytest.mark.parametrize("name", ["bl_24x16", "cyl_28x16"])
def test_dz_tangent_golden(gpu, name):
    """product vs the committed outputs of the reference's tangentdz code (tests/golden/lindz, oracle/make_golden.py --dz-tangent)"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lindz", name + ".npz"))
    a = H.make_case(str(g["kind"]), int(g["im"]), int(g["jm"]), gpu, with_w=True)
    wa, _ = H.residual_sequence(gpu, a)
    wd0, wd = np.asfortranarray(g["wd0"]), np.asfortranarray(g["wd"])
    for fn, key in (("coeffs_5p_dz_d", "dzd"), ("coeffs_5p_dz2_d", "dz2d")):
        z, zd = a.zeros_state(), a.zeros_state()
        getattr(gpu["f_lindz"], fn)(z, zd, wa, wd0, wd, *_dz_args(a))
        assert np.all(H.rel_err(zd, g[key]) < TOL), (fn, H.rel_err(zd, g[key]))


# ---------------------------------------------------------------------------------------------------------------------
# boundary fills of SURVEY.md 8(f3): isothermal wall, symmetry plane (primal and tangent), every side of the block
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["iso", "sym", "anti", "pres", "presnr", "blow", "isoprof"])
@pytest.mark.parametrize("kind,im,jm", [("bl", 60, 40), ("cyl", 70, 40), ("bl", 7, 7)])
def test_isothermal_wall_and_symmetry_fills(gpu, ref, name, kind, im, jm):
    import test_extra_bcs_cpu as T
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    wa0, _ = H.residual_sequence(gpu, a)
    wb0, _ = H.residual_sequence(ref, b)
    d = np.asfortranarray(np.random.default_rng(9).standard_normal(wa0.shape))
    for loc, interf in T.sides(a):
        wa, wb = wa0.copy(order="F"), wb0.copy(order="F")
        T.fill(gpu, name, a, wa, loc, interf)
        T.fill(ref, name, b, wb, loc, interf)
        assert np.all(H.rel_err(wa, wb) < TOL), (loc, H.rel_err(wa, wb))
        assert not np.array_equal(wa, wa0)
        wa, wb, da, db = wa0.copy(order="F"), wb0.copy(order="F"), d.copy(order="F"), d.copy(order="F")
        T.fill(gpu, name, a, wa, loc, interf, da)
        T.fill(ref, name, b, wb, loc, interf, db)
        assert np.all(H.rel_err(wa, wb) < TOL) and np.all(H.rel_err(da, db) < TOL), (loc, H.rel_err(da, db))


@pytest.mark.parametrize("fixture", ["bl_24x16", "cyl_28x16"])
def test_isothermal_wall_and_symmetry_golden(gpu, fixture):
    """product vs the committed outputs of the reference's routines (tests/golden/bcs, oracle/make_golden.py --bcs)"""
    import os
    import test_extra_bcs_cpu as T
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bcs", fixture + ".npz"))
    a = H.make_case(str(g["kind"]), int(g["im"]), int(g["jm"]), gpu, with_w=True)
    w0, _ = H.residual_sequence(gpu, a)
    d = np.asfortranarray(np.random.default_rng(int(g["seed"])).standard_normal(w0.shape))
    for name in T.NAMES:
        for loc, interf in T.sides(a):
            w, wd = w0.copy(order="F"), d.copy(order="F")
            T.fill(gpu, name, a, w, loc, interf, wd)
            T.check_against_golden(g, name, loc, a.gh, w, wd, w0, d, TOL)


# ---------------------------------------------------------------------------------------------------------------------
# resident C-ABI context (bcast_ctx_*): what a C / Fortran host binds; host arrays in and out, no torch memory
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,im,jm", [("bl", 60, 40), ("cyl", 70, 40)])
def test_c_abi_context_residual_and_norms(gpu, ref, kind, im, jm):
    from broadcast_b200.cabi_ctx import Context
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    wb, rb = H.residual_sequence(ref, b)
    ctx = Context(a)
    ctx.upload_state(a.w)
    ra = ctx.residual()
    wa = ctx.state()
    assert np.all(H.rel_err(wa, wb) < TOL) and np.all(H.rel_err(ra, rb) < TOL), H.rel_err(ra, rb)
    _, rp = H.residual_sequence(gpu, a)          # f2py-shaped path of the product: same kernels, same bits
    assert np.array_equal(ra, rp)
    n2, ninf = ctx.norms()
    m2, minf = ref["f_norm"].compute_norml2inf(rb, im, jm, b.gh)
    assert np.allclose(n2, m2, rtol=1e-12) and np.allclose(ninf, minf, rtol=1e-12)
    ctx.close()


@pytest.mark.parametrize("kind,im,jm", [("bl", 48, 26), ("cyl", 42, 30)])
def test_c_abi_context_jacobian_csr(gpu, ref, kind, im, jm):
    """bcast_ctx_jacobian_csr = the Python resident path (jacobian_hybrid + to_csr_device) array for array, and the reference's
    colour loop + remove_zero_jac + csr_matrix + division by the volume to TOL"""
    import scipy.sparse as sp
    from broadcast_b200.cabi_ctx import Context
    from broadcast_b200.resident import Block, jacobian_hybrid
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    coef = np.asfortranarray(np.random.default_rng(3).uniform(0.5, 1.5, size=(im, jm)))
    ctx = Context(a)
    ctx.upload_state(a.w)
    indptr, indices, data = ctx.jacobian_csr(coefdiag=coef, divide_by_vol=True)
    blk = Block(a)
    blk.apply_bcs()
    hj = jacobian_hybrid(blk, coefdiag=coef)
    p2, i2, d2 = (t.cpu().numpy() for t in hj.to_csr_device(divide_by_vol=True))
    assert np.array_equal(indptr, p2) and np.array_equal(indices, i2) and np.array_equal(data, d2)
    # against the reference loop on the CPU
    b = H.make_case(kind, im, jm, ref, with_w=True)
    wb, _ = H.residual_sequence(ref, b)
    jac, ia, ja = H.jacobian_sequence(ref, b, wb, None, coef)
    keep = np.abs(jac) > 2e-16
    n = 5 * im * jm
    A = sp.csr_matrix((jac[keep], (ia[keep], ja[keep])), shape=(n, n))
    vol = b.vol[b.gh:-b.gh, b.gh:-b.gh]
    rowvol = np.repeat(vol.reshape(-1), 5)      # row = e + 5 (j-1) + 5 jm (i-1): C-order ravel of (i, j), 5 equations each
    A = sp.diags(1.0 / rowvol) @ A
    B = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    diff = abs(A - B)
    assert diff.max() <= TOL * abs(A).max(), diff.max() / abs(A).max()
    ctx.close()


def _bl_case_with_extra_bcs(mods, im, jm, variant="iso+pressure"):
    """boundary layer with an ISOTHERMAL wall at Jlo and a PRESSURE outlet (characteristic blend) at Ihi instead of the adiabatic
    wall / extrapolation of card_bl2d_fv_npz.py -- the variants card_bl2d_fv.py:110 and card_bl2d_fv_cgns.py:102 select"""
    c = H.make_case("bl", im, jm, mods, with_w=True)
    p, g = c.phys, c.gh
    q = c.w[g, g]
    pr = (p["gam"] - 1.0) * (q[4] - 0.5 * (q[1] ** 2 + q[2] ** 2 + q[3] ** 2) / q[0])
    bcs = []
    for bc in c.bcs:
        x = np.random.default_rng(23).uniform(0.5, 1.5, im + 2 * g)        # the wall line of this case spans 1-gh .. im+gh
        if bc[0] == "wall" and variant == "iso+pressure":
            bcs.append(("wall_iso", bc[1], bc[2], 1.05 * pr / (q[0] * p["rgaz"]), p["rgaz"]))
        elif bc[0] == "wall" and variant == "blow_profile":      # the sensitivity driver's wall (BROADCAST_npz_sens.py:1763)
            bcs.append(("wall_blow_profile", bc[1], bc[2], 1e-3 * (x - 1.0)))
        elif bc[0] == "wall" and variant == "iso_profile":
            bcs.append(("wall_iso_profile", bc[1], bc[2], 1.05 * pr / (q[0] * p["rgaz"]) * x, p["rgaz"]))
        elif bc[0] == "outflow" and variant == "iso+pressure":
            bcs.append(("pressure", bc[1], bc[2], 0.97 * pr, 1.0))
        else:
            bcs.append(bc)
    c.bcs = bcs
    return c


@pytest.mark.parametrize("variant", ["iso+pressure", "blow_profile", "iso_profile"])
def test_jacobian_with_isothermal_wall_and_pressure_outlet(gpu, ref, variant):
    """the new boundary kinds inside the device colour loops: residual step, hybrid Jacobian (interior blocks + strip loop with the
    linearised isothermal-wall / pressure fills) and the C-ABI context, against the reference loop on the CPU"""
    import scipy.sparse as sp
    from broadcast_b200.cabi_ctx import Context
    from broadcast_b200.resident import Block, jacobian_hybrid
    im, jm = 44, 28
    a, b = _bl_case_with_extra_bcs(gpu, im, jm, variant), _bl_case_with_extra_bcs(ref, im, jm, variant)
    wa, ra = H.residual_sequence(gpu, a)
    wb, rb = H.residual_sequence(ref, b)
    assert np.all(H.rel_err(wa, wb) < TOL) and np.all(H.rel_err(ra, rb) < TOL)
    coef = np.asfortranarray(np.random.default_rng(6).uniform(0.5, 1.5, size=(im, jm)))
    blk = Block(a)
    blk.apply_bcs()
    assert np.array_equal(np.ascontiguousarray(blk.w.cpu().numpy().transpose(2, 1, 0)), wa)    # bcd_apply_bcs == the f2py-shaped fills
    hj = jacobian_hybrid(blk, coefdiag=coef)
    A = hj.to_scipy_csr()
    jb, ib, jbb = H.jacobian_sequence(ref, b, wb, None, coef)
    keep = np.abs(jb) > 2e-16
    n = 5 * im * jm
    B = sp.csr_matrix((jb[keep], (ib[keep], jbb[keep])), shape=(n, n))
    assert abs(A - B).max() < TOL * abs(B).max(), abs(A - B).max() / abs(B).max()
    ctx = Context(a)
    ctx.upload_state(a.w)
    assert np.array_equal(ctx.residual(), ra)
    indptr, indices, data = ctx.jacobian_csr(coefdiag=coef, divide_by_vol=False)
    p2, i2, d2 = (t.cpu().numpy() for t in hj.to_csr_device(divide_by_vol=False))
    assert np.array_equal(indptr, p2) and np.array_equal(indices, i2) and np.array_equal(data, d2)
    ctx.close()


def test_dz_tangent_colour_loop_device(gpu, ref):
    """device colour loop of the sensitivity matrices d/dw [Dz(w) mode], d/dw [Dz2(w) mode] (bcd_dz_tangent_coo) vs the loop of
    BROADCAST_npz_sens.py:1741-1800 run on the oracle: seeds, linearised boundary fills, f_lindz.coeffs_5p_dz_d / dz2_d on the real
    and the imaginary part of a mode, computejacobianfromdz; indices slot-exact, values to TOL"""
    import torch
    from broadcast_b200.resident import Block, dz_tangent_coo, _t
    im, jm = 30, 22
    a = H.make_case("bl", im, jm, gpu, with_w=True)
    b = H.make_case("bl", im, jm, ref, with_w=True)
    rng = np.random.default_rng(14)
    mr, mi = (np.asfortranarray(rng.standard_normal(a.w.shape)) for _ in range(2))
    blk = Block(a)
    blk.apply_bcs()
    out = dz_tangent_coo(blk, _t(mr, blk.device), _t(mi, blk.device))
    torch.cuda.synchronize()
    wb, _ = H.residual_sequence(ref, b)
    gh = b.gh
    s = 2 * gh + 1
    nb = 25 * s * s * im * jm
    J = {k: np.zeros(nb) for k in ("1r", "1i", "2r", "2i")}
    I1, K1, I2, K2 = (np.zeros(nb, np.int32) for _ in range(4))
    dz = b.zeros_state()
    for m in range(5):
        for l in range(s):
            for k in range(s):
                wd = b.zeros_state()
                ref["f_misc"].testvector(wd, m, l, k, gh, im, jm)
                ww = wb.copy(order="F")
                cases.apply_bcs_lin(b, ww, wd, ref["f_bnd"], ref["f_lin"])
                for part, mode in (("r", mr), ("i", mi)):
                    zd = b.zeros_state()
                    ref["f_lindz"].coeffs_5p_dz_d(dz, zd, ww, wd, mode, *_dz_args(b))
                    ref["f_misc"].computejacobianfromdz(J["1" + part], I1, K1, zd, m, l, k, gh, im, jm)
                    zd = b.zeros_state()
                    ref["f_lindz"].coeffs_5p_dz2_d(dz, zd, ww, wd, mode, *_dz_args(b))
                    ref["f_misc"].computejacobianfromdz(J["2" + part], I2, K2, zd, m, l, k, gh, im, jm)
    for wh, Ir, Kr in ((1, I1, K1), (2, I2, K2)):
        jr, ji, ia, ja = (t.cpu().numpy() for t in out[wh])
        assert np.array_equal(ia, Ir) and np.array_equal(ja, Kr)
        for got, key in ((jr, f"{wh}r"), (ji, f"{wh}i")):
            want = J[key]
            assert np.abs(want).max() > 0
            assert np.abs(got - want).max() < TOL * np.abs(want).max(), (key, np.abs(got - want).max() / np.abs(want).max())
