"""output formats of the Jacobian path (CPU tensors): COO -> CSR with scipy semantics, PETSc binary AIJ round trip, NPZ keys"""
import numpy as np
import scipy.sparse as sp
import torch

from broadcast_b200 import formats


def test_coo_to_csr_matches_scipy_with_duplicates():
    rng = np.random.default_rng(0)
    n, nnz = 60, 900
    ia = rng.integers(0, n, nnz); ja = rng.integers(0, n, nnz); v = rng.standard_normal(nnz)
    ip, idx, dat = formats.coo_to_csr(torch.from_numpy(v), torch.from_numpy(ia), torch.from_numpy(ja), n, n)
    A = sp.csr_matrix((v, (ia, ja)), shape=(n, n)); A.sort_indices()
    assert np.array_equal(ip.numpy(), A.indptr) and np.array_equal(idx.numpy(), A.indices)
    assert np.allclose(dat.numpy(), A.data, rtol=0, atol=1e-15)
    # a row block (slab): rows 20..39
    sel = (ia >= 20) & (ia < 40)
    ip2, idx2, dat2 = formats.coo_to_csr(torch.from_numpy(v[sel]), torch.from_numpy(ia[sel]), torch.from_numpy(ja[sel]), 20, n, row0=20)
    B = A[20:40]
    assert np.array_equal(ip2.numpy(), B.indptr) and np.array_equal(idx2.numpy(), B.indices)


def test_filter_and_divide_by_volume():
    im, jm, gh = 4, 3, 3
    vol = torch.arange(1, (im + 2 * gh) * (jm + 2 * gh) + 1, dtype=torch.float64).reshape(jm + 2 * gh, im + 2 * gh)
    ia = torch.tensor([0, 7, 5 * jm * 2 + 5 * 1 + 3, 2]); ja = torch.tensor([1, 2, 3, 4])
    jac = torch.tensor([1.0, 1e-17, 2.0, -3.0], dtype=torch.float64)
    v, r, c = formats.filter_divide(jac, ia, ja, vol=vol, jm=jm, gh=gh)
    assert r.tolist() == [0, 5 * jm * 2 + 8, 2]
    # reference formula: vol[IA // (5 jm) + gh, (IA % (5 jm)) // 5 + gh] with vol indexed [i, j]
    volF = vol.numpy().T
    want = [1.0 / volF[0 + gh, 0 + gh], 2.0 / volF[2 + gh, 1 + gh], -3.0 / volF[0 + gh, 0 + gh]]
    assert np.allclose(v.numpy(), want)


def test_petsc_binary_roundtrip(tmp_path):
    A = sp.random(40, 40, density=0.1, format="csr", random_state=1)
    A.sort_indices()
    for cplx in (True, False):
        p = str(tmp_path / ("J%d" % cplx))
        formats.write_petsc_aij(p, A.indptr, A.indices, A.data, 40, complex_scalar=cplx)
        raw = np.fromfile(p, dtype=">i4", count=4)
        assert raw.tolist() == [1211216, 40, 40, A.nnz]
        ip, idx, dat, shape = formats.read_petsc_aij(p, complex_scalar=cplx)
        assert shape == (40, 40) and np.array_equal(ip, A.indptr) and np.array_equal(idx, A.indices)
        assert np.array_equal(dat.real, A.data)
        import os
        assert os.path.getsize(p) == 16 + 4 * 40 + 4 * A.nnz + (16 if cplx else 8) * A.nnz


def test_petsc_binary_64bit_indices_and_chunks(tmp_path):
    """the layout of a --with-64-bit-indices PETSc (what C5's 6.09 G non-zeros need): int64 header / row lengths / columns,
    written in small chunks; 32-bit counts are refused instead of wrapped"""
    A = sp.random(300, 300, density=0.05, format="csr", random_state=2)
    A.sort_indices()
    p = str(tmp_path / "J64")
    formats.write_petsc_aij(p, A.indptr, A.indices, A.data, 300, complex_scalar=False, index64=True, chunk=97)
    raw = np.fromfile(p, dtype=">i8", count=4)
    assert raw.tolist() == [1211216, 300, 300, A.nnz]
    ip, idx, dat, shape = formats.read_petsc_aij(p, complex_scalar=False, index64=True)
    assert shape == (300, 300) and np.array_equal(ip, A.indptr) and np.array_equal(idx, A.indices) and np.array_equal(dat, A.data)
    import os
    assert os.path.getsize(p) == 32 + 8 * 300 + 8 * A.nnz + 8 * A.nnz
    # chunked 32-bit file == unchunked one
    q1, q2 = str(tmp_path / "a"), str(tmp_path / "b")
    formats.write_petsc_aij(q1, A.indptr, A.indices, A.data, 300, chunk=53)
    formats.write_petsc_aij(q2, A.indptr, A.indices, A.data, 300)
    assert open(q1, "rb").read() == open(q2, "rb").read()
    big = np.array([0, 2 ** 31], dtype=np.int64)
    import pytest
    with pytest.raises(ValueError):
        formats.write_petsc_aij(str(tmp_path / "c"), big, np.zeros(0, dtype=np.int32), np.zeros(0), 5, index64=False)


def test_npz_keys(tmp_path):
    f = str(tmp_path / "run")
    np.savez(f + ".npz", im=3, jm=2)
    formats.fill_npz(f, np.zeros((2, 2)), np.ones((2, 2)), [0, 1], [1, 0], [0.5, 0.25])
    formats.fill_npz_3d(f, [0], [0], [1.0], [1], [1], [2.0])
    d = np.load(f + ".npz")
    for k in ("im", "jm", "ResidualEndOfRun", "FlowSolutionEndOfRun", "IA", "JA", "Aij", "IAdz", "JAdz", "Aijdz", "IAdz2", "JAdz2", "Aijdz2"):
        assert k in d.files
