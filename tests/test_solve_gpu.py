"""Row f4 of SURVEY.md section 8, the step after the assembly: the Newton correction A dw = res solved on the device (csrc/solve.cu:
warp-per-row SpMV, block-Jacobi preconditioner from the 5 x 5 diagonal blocks, restarted GMRES with device-resident Krylov basis)
against scipy's sparse LU on the host -- the role PETSc / MUMPS play in the reference (misc/PETSc_func.py:137-152, 247-263) -- and the
adjoint system on the transposed CSR (cylinder.py:1090-1177)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _csr_to_scipy(ip, idx, dat, n):
    import scipy.sparse as sp
    return sp.csr_matrix((dat.cpu().numpy(), idx.cpu().numpy(), ip.cpu().numpy()), shape=(n, n))


@pytest.fixture(scope="module")
def system(gpu):
    """relaxed Jacobian of a small boundary-layer block as device CSR + the same matrix in scipy"""
    import torch
    from broadcast_b200.resident import Block, jacobian_hybrid
    c = H.make_case("bl", 48, 28, gpu, with_w=True)
    blk = Block(c)
    blk.apply_bcs()
    A0 = _csr_to_scipy(*jacobian_hybrid(blk).to_csr(), 5 * c.im * c.jm)
    # pseudo-time term coefdiag = cflm1 * vol (BROADCAST_npz.py:1067), cflm1 from the size of the diagonal
    vol = torch.as_tensor(np.ascontiguousarray(c.vol[c.gh:-c.gh, c.gh:-c.gh].T), device=blk.device)
    cflm1 = 0.2 * float(np.median(np.abs(A0.diagonal()))) / float(vol.median())
    coef = (cflm1 * vol).contiguous()
    ip, idx, dat = jacobian_hybrid(blk, coefdiag=coef).to_csr()
    n = 5 * c.im * c.jm
    return dict(case=c, blk=blk, coef=coef, csr=(ip, idx, dat), A=_csr_to_scipy(ip, idx, dat, n), n=n)


def test_spmv_matches_scipy(system):
    import torch
    from broadcast_b200.resident import csr_spmv
    ip, idx, dat = system["csr"]
    x = torch.as_tensor(np.random.default_rng(0).standard_normal(system["n"]), device=dat.device)
    y = csr_spmv(ip, idx, dat, x).cpu().numpy()
    yr = system["A"] @ x.cpu().numpy()
    assert np.abs(y - yr).max() <= 1e-13 * np.abs(yr).max()


def test_block_jacobi_is_the_inverse_of_the_diagonal_blocks(system):
    from broadcast_b200.resident import block_jacobi
    ip, idx, dat = system["csr"]
    dinv = block_jacobi(ip, idx, dat).cpu().numpy()
    A = system["A"].tocsr()
    for cell in (0, 17, system["n"] // 5 - 1):
        D = A[5 * cell:5 * cell + 5, 5 * cell:5 * cell + 5].toarray()
        Di = dinv[:, cell].reshape(5, 5)
        assert np.abs(Di @ D - np.eye(5)).max() < 1e-10


def test_gmres_matches_sparse_lu(system):
    import scipy.sparse.linalg as spla
    import torch
    from broadcast_b200.resident import gmres, csr_spmv
    ip, idx, dat = system["csr"]
    b = np.random.default_rng(1).standard_normal(system["n"])
    xr = spla.spsolve(system["A"].tocsc(), b)
    x, info = gmres(ip, idx, dat, torch.as_tensor(b, device=dat.device), restart=40, maxit=6000, rtol=1e-11)
    assert info["converged"], info
    r = b - csr_spmv(ip, idx, dat, x).cpu().numpy()
    assert abs(np.linalg.norm(r) / np.linalg.norm(b) - info["relres"]) <= 1e-3 * info["relres"] + 1e-16   # the reported true residual
    assert np.abs(x.cpu().numpy() - xr).max() <= 1e-7 * np.abs(xr).max(), (info, np.abs(x.cpu().numpy() - xr).max() / np.abs(xr).max())
    # an unpreconditioned run on the same (badly row-scaled) matrix must not beat the preconditioned one
    _, plain = gmres(ip, idx, dat, torch.as_tensor(b, device=dat.device), restart=40, maxit=info["matvecs"], rtol=1e-11, precond=False)
    assert (not plain["converged"]) or plain["matvecs"] >= info["matvecs"]


def test_adjoint_system_on_the_transposed_csr(system):
    """the adjoint solve of the sensitivity drivers: A^T x = b through bcd_csr_transpose_* and the same GMRES"""
    import scipy.sparse.linalg as spla
    import torch
    from broadcast_b200.resident import gmres, csr_transpose
    ip, idx, dat = system["csr"]
    tp, ti, td = csr_transpose(ip, idx, dat, system["n"])
    b = np.random.default_rng(2).standard_normal(system["n"])
    xr = spla.spsolve(system["A"].T.tocsc(), b)
    x, info = gmres(tp, ti, td, torch.as_tensor(b, device=dat.device), restart=40, maxit=6000, rtol=1e-11)
    assert info["converged"], info
    assert np.abs(x.cpu().numpy() - xr).max() <= 1e-7 * np.abs(xr).max()


def test_newton_step_against_the_reference_loop(gpu, ref):
    """one whole Newton iteration on the device (fills, residual, Jacobian -> CSR, solve, update) against the reference's sequence on
    the oracle: 245-colour loop -> remove_zero_jac -> csr_matrix -> LU solve of iterNewton -> w += dw (BROADCAST_npz.py:1068-1172)"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch
    from broadcast_b200.resident import Block, newton_step
    im, jm = 30, 22
    a = H.make_case("bl", im, jm, gpu, with_w=True)
    b = H.make_case("bl", im, jm, ref, with_w=True)
    gh = a.gh
    wb, rb = H.residual_sequence(ref, b)
    cflm1 = 40.0
    coef = np.asfortranarray(cflm1 * b.vol[gh:-gh, gh:-gh])
    jac, ia, ja = H.jacobian_sequence(ref, b, wb, None, coef)
    keep = np.abs(jac) > 2e-16
    n = 5 * im * jm
    A = sp.csr_matrix((jac[keep], (ia[keep], ja[keep])), shape=(n, n))
    dw_ref = spla.spsolve(A.tocsc(), np.ravel(rb[gh:-gh, gh:-gh, :])).reshape(im, jm, 5)
    blk = Block(a)
    w0 = blk.w.clone()
    dw, info = newton_step(blk, coefdiag=coef, rtol=1e-12)
    assert info["converged"], info
    err = np.abs(dw.cpu().numpy() - dw_ref).max() / np.abs(dw_ref).max()
    assert err < 1e-7, (err, info)
    # w += dw on the interior cells
    upd = (blk.w - w0)[:, gh:gh + jm, gh:gh + im].permute(2, 1, 0).cpu().numpy()
    assert np.abs(upd - dw.cpu().numpy()).max() <= 1e-15 * max(1.0, float(blk.w.abs().max()))


def test_newton_loop_follows_the_reference_protocol(gpu):
    """resident.newton_loop = the iteration protocol of BROADCAST_npz.py:1007-1172 on the device: the first pass is newton_step
    with coefdiag = vol / dt (relaxation factor 1), dt from the CFL number and the first cell height; later passes scale the
    relaxation by the residual ratios of the first three equations; every solve converges at a moderate CFL number"""
    import torch
    from broadcast_b200.resident import Block, newton_loop, newton_step
    c = H.make_case("bl", 48, 30, gpu, with_w=True)
    gh, im, jm = c.gh, c.im, c.jm
    cfl = 1.0
    dt = cfl * float(c.yc[gh, gh + 1] - c.yc[gh, gh]) / (1.0 / float(c.phys["mach"]) + 1.0)
    blk = Block(c)
    hist = newton_loop(blk, cfl=cfl, nit=3, rtol=1e-10)
    assert len(hist) == 3 and all(h[4] > 0 for h in hist), [h[4] for h in hist]
    assert abs(hist[0][3] * dt - 1.0) < 1e-12                                  # first pass: cflm1 = 1 / dt
    n0, i0 = hist[0][1], hist[0][2]
    for it, norm, ninf, cflm1, _ in hist[1:]:
        r = max((norm[:3] / n0[:3]).max(), (ninf[:3] / i0[:3]).max())
        assert abs(cflm1 * dt / r - 1.0) < 1e-12
    # the first pass alone, through newton_step on a fresh block: same state afterwards
    blk2 = Block(c)
    vol = blk2.vol[gh:gh + jm, gh:gh + im].contiguous()
    newton_step(blk2, coefdiag=(1.0 / dt) * vol, rtol=1e-10)
    blk3 = Block(c)
    newton_loop(blk3, cfl=cfl, nit=1, rtol=1e-10)
    assert torch.equal(blk2.w, blk3.w)
