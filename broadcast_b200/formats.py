"""Output side of the Jacobian path: what the reference's Python glue does after the colour loop.

* ``remove_zero_jac``  BROADCAST_npz.py:129-135 (keep ``|v| > 2e-16``)
* division by the row cell's volume, BROADCAST_npz.py:1206-1209 (a pure-Python loop over nnz in the reference)
* COO -> CSR with duplicate summation, misc/PETSc_func.py:85 (``scipy.sparse.csr_matrix((A,(I,J)))``)
* PETSc binary AIJ files ``Jacsurvol`` / ``Dz`` / ``Dz2`` that biglobal_cyl.py:47-52 and resolvent_all.py:612-617 load
  with ``PETSc.Viewer().createBinary`` (big-endian: classid 1211216, M, N, nnz, row lengths, column indices, values;
  complex builds store (re, im) pairs)
* the ``IA / JA / Aij`` (+ ``IAdz ...``) keys of the run's ``.npz`` file, BROADCAST_npz.py:311-333

The filter, the division and the CSR construction run on the device with torch ops (sort / unique / index_add): they are
data movement, not kernels of the hot path; the writers are host I/O.
"""
from __future__ import annotations

import numpy as np
import torch

MAT_FILE_CLASSID = 1211216


def filter_divide(jac, ia, ja, *, thresh=2e-16, vol=None, jm=None, gh=None, ioff=0):
    """remove_zero_jac then Jacvol[k] = Jac[k] / vol[IA[k] // (5 jm) + gh, (IA[k] % (5 jm)) // 5 + gh] on the device.
    ``vol``: torch tensor (jm+2gh, im+2gh) = memory image of the Fortran array (local slab columns offset by ``ioff``)."""
    keep = jac.abs() > thresh
    jac, ia, ja = jac[keep], ia[keep].to(torch.int64), ja[keep].to(torch.int64)
    if vol is not None:
        ci = torch.div(ia, 5 * jm, rounding_mode="floor") - ioff + gh
        cj = torch.div(ia % (5 * jm), 5, rounding_mode="floor") + gh
        jac = jac / vol[cj, ci]
    return jac, ia, ja


def coo_to_csr(jac, ia, ja, nrows, ncols, row0=0):
    """CSR (indptr int64, indices int32/int64, data) of the rows [row0, row0 + nrows) with duplicates summed and columns
    sorted within a row -- scipy's csr_matrix((A,(I,J))) semantics -- on the device."""
    ia = ia.to(torch.int64) - row0
    key = ia * ncols + ja.to(torch.int64)
    key, order = torch.sort(key, stable=True)
    data = jac[order]
    ukey, inv = torch.unique_consecutive(key, return_inverse=True)
    out = torch.zeros(ukey.numel(), dtype=data.dtype, device=data.device)
    out.index_add_(0, inv, data)
    rows = torch.div(ukey, ncols, rounding_mode="floor")
    cols = ukey - rows * ncols
    counts = torch.bincount(rows, minlength=nrows)
    indptr = torch.zeros(nrows + 1, dtype=torch.int64, device=data.device)
    indptr[1:] = torch.cumsum(counts, 0)
    return indptr, cols, out


def write_petsc_aij(path, indptr, indices, data, ncols, complex_scalar=True, index64=None, chunk=1 << 24):
    """PETSc binary AIJ (Mat) file, loadable with ``PETSc.Mat().load(PETSc.Viewer().createBinary(path, 'r'))``.
    The reference's stability drivers run a complex-scalar PETSc (biglobal_cyl.py, resolvent_all.py): values are
    written as complex128 unless ``complex_scalar`` is False.
    ``index64``: the layout of a PETSc configured ``--with-64-bit-indices`` (header, row lengths and column indices are
    big-endian int64); chosen automatically when the matrix does not fit 32-bit counts (C5: 6.09 G non-zeros).  The arrays are
    written in chunks of ``chunk`` entries, so no big-endian copy of the whole matrix is ever held."""
    indptr = np.asarray(indptr, dtype=np.int64)
    m = indptr.size - 1
    nnz = int(indptr[-1])
    if index64 is None:
        index64 = nnz >= 2 ** 31 or m >= 2 ** 31 or ncols >= 2 ** 31
    if not index64 and (nnz >= 2 ** 31 or m >= 2 ** 31):
        raise ValueError("PETSc's default binary format stores 32-bit counts: pass index64=True (a --with-64-bit-indices PETSc)")
    it = ">i8" if index64 else ">i4"

    def put(fh, arr, dtype):
        arr = np.asarray(arr)
        for a in range(0, arr.shape[0], chunk):
            arr[a:a + chunk].astype(dtype).tofile(fh)

    with open(path, "wb") as fh:
        np.array([MAT_FILE_CLASSID, m, ncols, nnz], dtype=it).tofile(fh)
        for a in range(0, m, chunk):
            np.diff(indptr[a:min(a + chunk, m) + 1]).astype(it).tofile(fh)
        put(fh, indices, it)
        put(fh, data, ">c16" if complex_scalar else ">f8")


def read_petsc_aij(path, complex_scalar=True, index64=False):
    """inverse of write_petsc_aij -> (indptr, indices, data, (M, N))"""
    it = ">i8" if index64 else ">i4"
    with open(path, "rb") as fh:
        hdr = np.fromfile(fh, dtype=it, count=4)
        if hdr[0] != MAT_FILE_CLASSID:
            raise ValueError("not a PETSc binary Mat file (or the other index width)")
        m, n, nnz = int(hdr[1]), int(hdr[2]), int(hdr[3])
        rowlen = np.fromfile(fh, dtype=it, count=m).astype(np.int64)
        indices = np.fromfile(fh, dtype=it, count=nnz).astype(np.int64 if index64 else np.int32)
        data = np.fromfile(fh, dtype=">c16" if complex_scalar else ">f8", count=nnz)
    indptr = np.concatenate(([0], np.cumsum(rowlen)))
    return indptr, indices, data.astype(np.complex128 if complex_scalar else np.float64), (m, n)


def fill_npz(filename, w, res, ia, ja, jacvol):
    """BROADCAST_npz.py:311-320 (fillNPZ): adds the end-of-run state and the Jacobian lists to ``filename``.npz"""
    import os
    dic = dict(np.load(filename + ".npz")) if os.path.exists(filename + ".npz") else {}
    dic["ResidualEndOfRun"] = res
    dic["FlowSolutionEndOfRun"] = w
    dic["IA"] = np.asarray(ia)
    dic["JA"] = np.asarray(ja)
    dic["Aij"] = np.asarray(jacvol)
    np.savez(filename + ".npz", **dic)


def fill_npz_3d(filename, iadz, jadz, jacdz, iadz2, jadz2, jacdz2):
    """BROADCAST_npz.py:323-333 (fillNPZ_3D)"""
    dic = dict(np.load(filename + ".npz"))
    dic.update(IAdz=np.asarray(iadz), JAdz=np.asarray(jadz), Aijdz=np.asarray(jacdz), IAdz2=np.asarray(iadz2), JAdz2=np.asarray(jadz2),
               Aijdz2=np.asarray(jacdz2))
    np.savez(filename + ".npz", **dic)
