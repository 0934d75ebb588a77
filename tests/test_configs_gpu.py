"""GPU parity at the sizes BASELINE.json names: the CUDA path vs oracle/_ref on
C1 500 x 150 (card_bl2d_fv_npz.py:39-40), C2 / C3 630 x 300 O-mesh (card_cyl2d.py:41-42), C4 = C1 + Dz / Dz2, C5 8192 x 2048.

Every comparison reports, per equation,
  (i)   the forward error relative to the plane maximum (the metric of round 1),
  (ii)  the backward error: difference / magnitude of the face fluxes of the cell (helpers.flux_scale) -- what north_star's
        "1e-12 relative" means for a quantity that is a difference of fluxes,
  (iii) the same two numbers between the reference's own FMA and no-FMA builds on the same input: the noise floor.
Assertions: backward error < 1e-13; forward error <= max(1e-12, 4 x floor).  The measured numbers are appended to
gpurun_out/r2_parity_configs.jsonl (copied to profiles/ by the round's summary).  The authors' own check of the tangent
(BROADCAST_npz.py:1091-1125: finite differences of the residual) is the pin of the oracle itself (tests/test_oracle_cpu.py)."""
import json
import os
import time

import numpy as np
import pytest

import helpers as H
from broadcast_b200 import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(ROOT, "gpurun_out", "r2_parity_configs.jsonl")

CONFIGS = {"C1": ("bl", 500, 150), "C2": ("cyl", 630, 300), "C5w": ("bl", 1024, 512)}


def _log(rec):
    os.makedirs(os.path.dirname(LOG), exist_ok=True)
    with open(LOG, "a") as f:
        f.write(json.dumps(rec, default=lambda x: np.asarray(x).tolist()) + "\n")


def _block_residual(blk, variant):
    r = blk.residual(variant=variant).clone()
    return np.asfortranarray(r.cpu().numpy().transpose(2, 1, 0))   # (planes, j, i) image -> (i, j, planes)


@pytest.mark.parametrize("cfg", ["C1", "C2", "C5w", "C5"])
def test_residual_at_named_config(gpu, ref, cfg):
    """boundary fills + residual through the drop-in entry points, then every kernel variant on the resident block:
    0 = k_residual_fast (tile kernel, default), 5 = k_residual_march (j-marching kernel), 1 = reference-shaped pipeline (the reference's operation
    order: must sit at the FMA floor)"""
    import torch
    from broadcast_b200.resident import Block
    kind, im, jm = CONFIGS.get(cfg, ("bl", 8192, 2048))
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    for name in ("nx", "ny", "vol", "volf"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name      # geometry bit-exact at this size too
    t0 = time.time()
    wb, rb = H.residual_sequence(ref, b)
    t_ref = time.time() - t0
    floor = H.fma_floor(b)
    wa, ra = H.residual_sequence(gpu, a)
    # boundary fills: a few ulp of the filled value (pow / sqrt inside; the device's pow is not glibc's), the rest of w is a copy
    assert np.all(H.rel_err(wa, wb) < 1e-12), H.rel_err(wa, wb)
    rec = {"test": "residual", "config": cfg, "grid": [im, jm], "oracle_s": t_ref, "floor": floor, "variants": {}}
    e = H.assert_residual_parity(ra, rb, b, wb, floor=floor, what=(cfg, "drop-in f_sch.flux_num_dnc5_2d"))
    rec["variants"]["dropin"] = e
    blk = Block(a)
    blk.apply_bcs()
    outs = {}
    for v, name in ((5, "march"), (4, "tile"), (6, "bulk"), (0, "default"), (1, "generic")):
        outs[v] = _block_residual(blk, v)
        rec["variants"][name] = H.assert_residual_parity(outs[v], rb, b, wb, floor=floor, what=(cfg, name))
    # the two fused kernels evaluate the same formulas on the same operands: they differ at most by FMA contraction choices
    rec["march_vs_tile"] = H.residual_errors(outs[5], outs[4], b, wb)
    assert np.all(rec["march_vs_tile"]["backward"] < 1e-14), rec["march_vs_tile"]
    # the bulk-staged tile kernel (TMA / bulk copies of w and of every metric) runs the phase functions of the LDG tile kernel on
    # shared-memory images of the same numbers: bit for bit the same residual; the default is one of the two
    assert np.array_equal(outs[6], outs[4]), np.abs(outs[6] - outs[4]).max()
    assert np.array_equal(outs[0], outs[4])
    gh = a.gh
    assert not np.any(outs[0][:gh]) and not np.any(outs[0][:, :gh]) and not np.any(outs[0][-gh:]) and not np.any(outs[0][:, -gh:])
    _log(rec)


@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_tangent_at_named_config(gpu, ref, cfg):
    """one dense tangent (random direction, linearised fills included) at the named size"""
    kind, im, jm = CONFIGS[cfg]
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    wa, _ = H.residual_sequence(gpu, a)
    wb, _ = H.residual_sequence(ref, b)
    wd = np.asfortranarray(np.random.default_rng(1).standard_normal(wa.shape))
    wda, rda = H.tangent_sequence(gpu, a, wa, wd)
    wdb, rdb = H.tangent_sequence(ref, b, wb, wd)
    assert np.all(H.rel_err(wda, wdb) < 1e-13), H.rel_err(wda, wdb)
    # the tangent is a difference of linearised face fluxes of size |dF/dw| |wd| ~ flux_scale / W * |wd|: with wd ~ N(0, 1)
    # the scale of equation e is flux_scale_e / min_e W_e of the cell -- bounded below by using the plane maximum metric with
    # the reference's own FMA spread of the same tangent as floor
    from oracle import refmods
    _, rdf = H.tangent_sequence(refmods.make(fast=True), b, wb, wd)
    gh = a.gh
    fl = H.rel_err(rdf[gh:-gh, gh:-gh], rdb[gh:-gh, gh:-gh])
    err = H.rel_err(rda[gh:-gh, gh:-gh], rdb[gh:-gh, gh:-gh])
    _log({"test": "tangent", "config": cfg, "plane_max": err, "floor_plane_max": fl})
    assert np.all(err <= np.maximum(1e-12, 4.0 * fl)), (err, fl)


def _oracle_csr(ref, b, wb, coef, colours=None):
    """the reference's assembly: colour loop -> remove_zero_jac (2e-16) -> csr (duplicates summed) -> / vol  (BROADCAST_npz.py:
    1068-1137, 1206-1209; misc/PETSc_func.py:71-95)"""
    import scipy.sparse as sp
    jb, ib, jbb = H.jacobian_sequence(ref, b, wb, colours, coef)
    keep = np.abs(jb) > 2e-16
    jb, ib, jbb = jb[keep], ib[keep], jbb[keep]
    gh, jm = b.gh, b.jm
    vol = b.vol[ib // (5 * jm) + gh, (ib % (5 * jm)) // 5 + gh]
    n = 5 * b.im * b.jm
    B = sp.csr_matrix((jb / vol, (ib, jbb)), shape=(n, n))
    B.sort_indices()
    return B


def _compare_csr(A, B, what, restrict_cols=None):
    """pattern identical up to entries at rounding level; values relative to the matrix maximum and relative to the row maximum"""
    import scipy.sparse as sp
    if restrict_cols is not None:
        mask = np.zeros(A.shape[1], dtype=bool)
        mask[restrict_cols] = True
        A = A.tocsc()[:, mask].tocsr()
        B = B.tocsc()[:, mask].tocsr()
    A.sort_indices(); B.sort_indices()
    D = (A - B).tocoo()
    scale = np.abs(B.data).max()
    emax = np.abs(D.data).max() / scale if D.nnz else 0.0
    rowmax = np.maximum(np.asarray(abs(B).max(axis=1).todense()).ravel(), 1e-300)
    erow = (np.abs(D.data) / rowmax[D.row]).max() if D.nnz else 0.0
    PA = sp.csr_matrix((np.ones_like(A.data), A.indices, A.indptr), shape=A.shape)
    PB = sp.csr_matrix((np.ones_like(B.data), B.indices, B.indptr), shape=B.shape)
    X = (PA - PB).tocoo()
    odd = X.data != 0
    nflip, flipmax = int(odd.sum()), 0.0
    if nflip:
        flipmax = float(np.abs(np.asarray((A + B)[X.row[odd], X.col[odd]])).ravel().max() / scale)
    rec = {"test": "jacobian_csr", "what": what, "nnz": int(B.nnz), "err_over_matrix_max": float(emax), "err_over_row_max": float(erow),
           "pattern_flips": nflip, "largest_flipped_entry_over_max": flipmax}
    _log(rec)
    assert emax < 1e-12, rec
    assert erow < 1e-10, rec     # rows whose largest entry is itself small relative to the matrix (far field)
    assert flipmax < 1e-15, rec  # entries that flip across the 2e-16 filter are themselves at rounding level
    return rec


@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_full_jacobian_csr_at_named_config(gpu, ref, cfg):
    """C1 / C3: the hybrid assembly (face linearisation + boundary strips) -> zero filter -> CSR / vol on the device against the
    reference's 245-colour loop on the oracle (9 s at C1, 27 s at C2), as canonical CSR"""
    import scipy.sparse as sp
    from broadcast_b200.resident import Block, jacobian_hybrid
    kind, im, jm = CONFIGS[cfg]
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    coef = np.asfortranarray(np.random.default_rng(5).uniform(0.5, 1.5, size=(im, jm)) * b.vol[b.gh:-b.gh, b.gh:-b.gh])
    blk = Block(a)
    blk.apply_bcs()
    ip, idx, dat = jacobian_hybrid(blk, coefdiag=coef).to_csr(divide_by_vol=True)
    n = 5 * im * jm
    A = sp.csr_matrix((dat.cpu().numpy(), idx.cpu().numpy(), ip.cpu().numpy()), shape=(n, n))
    wb, _ = H.residual_sequence(ref, b)
    t0 = time.time()
    B = _oracle_csr(ref, b, wb, coef)
    rec = _compare_csr(A, B, cfg)
    rec["oracle_s"] = time.time() - t0
    # A v = -tangent(v) / vol + coefdiag / vol v on the device matrix (the defining property, at full size)
    v = np.random.default_rng(9).standard_normal(n)
    y = A @ v
    yb = B @ v
    assert np.abs(y - yb).max() <= 1e-12 * np.abs(yb).max()


def test_jacobian_sampled_colours_on_c5_window(gpu, ref):
    """C5 recipe on a 1024 x 512 block: the device CSR against the reference colour loop for a sample of 8 of the 49 (l, k) offsets
    x 5 variables (40 seed vectors; the full loop would need 10 GB of COO lists on the host)"""
    import scipy.sparse as sp
    from broadcast_b200.resident import Block, jacobian_hybrid
    kind, im, jm = CONFIGS["C5w"]
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    coef = np.zeros((im, jm), order="F")
    blk = Block(a)
    blk.apply_bcs()
    ip, idx, dat = jacobian_hybrid(blk, coefdiag=coef).to_csr(divide_by_vol=True)
    n = 5 * im * jm
    A = sp.csr_matrix((dat.cpu().numpy(), idx.cpu().numpy(), ip.cpu().numpy()), shape=(n, n))
    wb, _ = H.residual_sequence(ref, b)
    lk = [(0, 0), (3, 2), (6, 6), (5, 0), (0, 4), (4, 1), (2, 5), (1, 3)]
    colours = [(m, l, k) for (l, k) in lk for m in range(5)]
    B = _oracle_csr(ref, b, wb, coef, colours)
    s = 2 * b.gh + 1
    cols = []
    for (l, k) in lk:
        ii, jj = np.meshgrid(np.arange(l, im, s), np.arange(k, jm, s), indexing="ij")
        base = (5 * jj + 5 * jm * ii).ravel()
        cols.append((base[:, None] + np.arange(5)[None, :]).ravel())
    _compare_csr(A, B, "C5 recipe 1024x512, 40 of 245 seed vectors", restrict_cols=np.concatenate(cols))


def test_dz_loop_at_c4(gpu, ref):
    """C4: the Dz / Dz2 colour loop of BROADCAST_npz.py:1231-1246 at 500 x 150 on the device against the reference loop for a
    sample of colours (values) and in full for the integer lists"""
    from broadcast_b200.resident import Block, dz_coo
    from test_parity_gpu import _dz_args
    kind, im, jm = CONFIGS["C1"]
    a = H.make_case(kind, im, jm, gpu, with_w=True)
    b = H.make_case(kind, im, jm, ref, with_w=True)
    blk = Block(a)
    blk.apply_bcs()
    out = dz_coo(blk)
    wb, _ = H.residual_sequence(ref, b)
    gh = b.gh
    s = 2 * gh + 1
    n5 = 5 * im * jm
    nb = 25 * s * s * im * jm
    J1, I1, K1 = np.zeros(nb), np.zeros(nb, np.int32), np.zeros(nb, np.int32)
    J2, I2, K2 = np.zeros(nb), np.zeros(nb, np.int32), np.zeros(nb, np.int32)
    wd = b.zeros_state()
    dz, dz2 = b.zeros_state(), b.zeros_state()
    colours = [(m, l, k) for m in range(5) for (l, k) in ((0, 0), (3, 2), (6, 6), (5, 0), (0, 4), (2, 5))]
    segs = []
    for (m, l, k) in colours:
        wd *= 0.0
        ref["f_misc"].testvector(wd, m, l, k, gh, im, jm)
        ww = wb.copy(order="F")
        cases.apply_bcs_lin(b, ww, wd, ref["f_bnd"], ref["f_lin"])
        ref["f_dz"].coeffs_5p_dz(dz, ww, wd, *_dz_args(b))
        ref["f_dz"].coeffs_5p_dz2(dz2, ww, wd, *_dz_args(b))
        ref["f_misc"].computejacobianfromdz(J1, I1, K1, dz, m, l, k, gh, im, jm)
        ref["f_misc"].computejacobianfromdz(J2, I2, K2, dz2, m, l, k, gh, im, jm)
        base = k * n5 + l * n5 * s + m * n5 * s * s
        segs.append(np.arange(base, base + n5))
    sel = np.concatenate(segs)
    for wh, (J, I, K) in ((1, (J1, I1, K1)), (2, (J2, I2, K2))):
        jac, ia, ja = (t.cpu().numpy() for t in out[wh])
        assert np.array_equal(ia[sel], I[sel]) and np.array_equal(ja[sel], K[sel])
        e = np.abs(jac[sel] - J[sel]).max() / np.abs(J[sel]).max()
        _log({"test": "dz_loop", "config": "C4", "which": wh, "err_over_max": float(e), "colours": len(colours)})
        assert e < 1e-12, e
