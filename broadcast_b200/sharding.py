"""i-slab sharding of one structured block over the GPUs of a box.

The reference numbers Jacobian rows i-major (``ia = e-1 + 5(j-1) + 5 jm (i-1)``, misc/ComputeJacobian.f90:526), so a
contiguous range of columns ``i`` is a contiguous block of matrix rows -- PETSc's ``mpiaij`` ownership range
(misc/PETSc_func.py:87-90).  Every rank owns ``im/N`` columns plus ``gh`` halo columns on each slab-internal
edge; the only data-path exchange is the neighbour swap of those ``gh`` columns of ``w`` before an evaluation
(residual or Jacobian: the colour seeds are analytic, so nothing is exchanged inside the colour loop).

This module is host logic only (numpy / torch.distributed); it runs unchanged on CPU tensors with the gloo
backend, which is how tests/test_sharding_cpu.py covers it without a GPU.
"""
from __future__ import annotations

import numpy as np

from .cases import Case


def slab_range(im: int, rank: int, world: int):
    """global (1-based, inclusive) column range [lo, hi] owned by ``rank``"""
    base, rem = divmod(im, world)
    lo = rank * base + min(rank, rem) + 1
    n = base + (1 if rank < rem else 0)
    return lo, lo + n - 1


def slab_of(case: Case, rank: int, world: int):
    """The i-slab of ``case`` owned by ``rank`` as a Case of its own (local indices), with its gh halo columns, plus the
    slab descriptor (ioff, im_global, edges) the device entry points need (include/broadcast_b200.h, bcd_slab_begin)."""
    if world == 1:
        return case, (0, case.im, 0)
    if case.periodic_i:
        raise NotImplementedError("i-slabs of an i-periodic block (O-mesh): shard over colours instead (SURVEY.md 8(e))")
    gh, im, jm = case.gh, case.im, case.jm
    lo, hi = slab_range(im, rank, world)
    n = hi - lo + 1
    if n < 2 * gh + 1:
        raise ValueError(f"slab of {n} columns is narrower than the stencil ({2 * gh + 1}): use fewer ranks")
    cs = slice(lo - 1, hi + 2 * gh)          # storage columns of cells lo-gh .. hi+gh
    ns = slice(lo - 1, hi + 2 * gh + 1)
    first, last = rank == 0, rank == world - 1
    F = np.asfortranarray
    bcs = []
    for bc in case.bcs:
        kind = bc[0]
        itf = np.array(bc[2], dtype=float)
        if kind == "inflow":
            if first:
                bcs.append(bc)
        elif kind == "outflow":
            if last:
                it = itf.copy(); it[0, 0] = n; it[1, 0] = n
                bcs.append((kind, bc[1], F(it)))
        elif kind == "noref":   # global i range [1-gh, im]: every slab fills its own columns and its halo
            it = itf.copy()
            it[0, 0] = 1 - gh
            it[1, 0] = n if last else n + gh
            g0 = lo - gh           # global cell index of local column 1-gh
            wbd = bc[3][g0 - (1 - gh): g0 - (1 - gh) + (int(it[1, 0]) - int(it[0, 0]) + 1), :]
            bcs.append((kind, bc[1], F(it), F(wbd)))
        elif kind == "wall":
            it = itf.copy(); it[0, 0] = 1 - gh; it[1, 0] = n + gh
            bcs.append((kind, bc[1], F(it)))
        else:
            raise NotImplementedError(kind)
    sl = Case(name=f"{case.name}_slab{rank}of{world}", im=n, jm=jm, gh=gh, phys=case.phys, k2=case.k2, k4=case.k4,
              x0=F(case.x0[ns]), y0=F(case.y0[ns]), nx=F(case.nx[ns]), ny=F(case.ny[ns]), xc=F(case.xc[cs]), yc=F(case.yc[cs]),
              vol=F(case.vol[cs]), volf=F(case.volf[cs]), w=F(case.w[cs]), bcs=bcs, periodic_i=False, scheme=case.scheme)
    edges = (0 if first else 1) | (0 if last else 2)
    return sl, (lo - 1, im, edges)


class HaloExchange:
    """neighbour exchange of the gh halo columns of a state held as a torch tensor of shape (planes, jm+2gh, im+2gh)
    (the memory image of the Fortran array (im+2gh, jm+2gh, planes)): NCCL send/recv over NVLink on the GPUs, gloo on CPU.
    All rows (ghost rows included) are exchanged, so corner ghosts of a slab are its neighbour's boundary ghosts."""

    def __init__(self, gh: int, rank: int, world: int, group=None):
        self.gh, self.rank, self.world, self.group = gh, rank, world, group
        self._buf = {}

    def _b(self, key, like, shape):
        b = self._buf.get(key)
        if b is None or b.shape != shape or b.device != like.device or b.dtype != like.dtype:
            import torch
            b = torch.empty(shape, dtype=like.dtype, device=like.device)
            self._buf[key] = b
        return b

    def bytes_per_exchange(self, w) -> int:
        sides = (1 if self.rank > 0 else 0) + (1 if self.rank < self.world - 1 else 0)
        return sides * w.shape[0] * w.shape[1] * self.gh * w.element_size()

    def __call__(self, w):
        if self.world == 1:
            return
        import torch.distributed as dist
        gh = self.gh
        ni = w.shape[2]
        im = ni - 2 * gh
        shape = (w.shape[0], w.shape[1], gh)
        ops = []
        if self.rank > 0:
            sl, rl = self._b("sl", w, shape), self._b("rl", w, shape)
            sl.copy_(w[:, :, gh:2 * gh])                 # my first owned columns -> left neighbour's right halo
            ops += [dist.P2POp(dist.isend, sl, self.rank - 1, self.group), dist.P2POp(dist.irecv, rl, self.rank - 1, self.group)]
        if self.rank < self.world - 1:
            sr, rr = self._b("sr", w, shape), self._b("rr", w, shape)
            sr.copy_(w[:, :, im:im + gh])                # my last owned columns -> right neighbour's left halo
            ops += [dist.P2POp(dist.isend, sr, self.rank + 1, self.group), dist.P2POp(dist.irecv, rr, self.rank + 1, self.group)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        if self.rank > 0:
            w[:, :, 0:gh].copy_(self._buf["rl"])
        if self.rank < self.world - 1:
            w[:, :, im + gh:im + 2 * gh].copy_(self._buf["rr"])


def colour_range(ncolours: int, rank: int, world: int):
    """[c0, c1) of the ``ncolours`` = (2gh+1)^2 colour passes owned by ``rank``: colour sharding for grids too small for
    i-slabs and for the i-periodic O-mesh.  Every rank holds the whole state and evaluates its passes on the
    full grid; nothing is exchanged, the COO entries of different ranks are disjoint column subsets."""
    base, rem = divmod(ncolours, world)
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


def merge_colour_shards(parts, n, thresh=2e-16):
    """host merge of per-rank COO triples (jac, ia, ja) produced under colour sharding into one scipy CSR matrix of order n:
    zero filter of BROADCAST_npz.py:129-135 on each part, then duplicate summation (scipy csr semantics, as the reference)"""
    import scipy.sparse as sp
    js, rs, cs = [], [], []
    for jac, ia, ja in parts:
        jac, ia, ja = np.asarray(jac), np.asarray(ia), np.asarray(ja)
        keep = np.abs(jac) > thresh
        js.append(jac[keep]); rs.append(ia[keep]); cs.append(ja[keep])
    return sp.csr_matrix((np.concatenate(js), (np.concatenate(rs), np.concatenate(cs))), shape=(n, n))


def gather_row_blocks(parts):
    """host gather of per-rank CSR row blocks [(indptr, indices, data), ...] (rank order = row order) into one CSR triple"""
    indptr = [np.asarray(parts[0][0], dtype=np.int64)]
    for ip, _, _ in parts[1:]:
        ip = np.asarray(ip, dtype=np.int64)
        indptr.append(ip[1:] + indptr[-1][-1])
    return np.concatenate(indptr), np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts])
