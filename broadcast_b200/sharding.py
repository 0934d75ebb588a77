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

import os

import numpy as np

from .cases import Case


def slab_range(im: int, rank: int, world: int, bounds=None):
    """global (1-based, inclusive) column range [lo, hi] owned by ``rank``; ``bounds`` (world + 1 increasing column counts from 0 to
    im) replaces the even split (the tapered pipeline of resident.StreamedBlock)"""
    if bounds is not None:
        if len(bounds) != world + 1 or bounds[0] != 0 or bounds[-1] != im or any(b <= a for a, b in zip(bounds[:-1], bounds[1:])):
            raise ValueError("bounds must be world + 1 increasing column counts from 0 to im")
        return int(bounds[rank]) + 1, int(bounds[rank + 1])
    base, rem = divmod(im, world)
    lo = rank * base + min(rank, rem) + 1
    n = base + (1 if rank < rem else 0)
    return lo, lo + n - 1


def slab_of(case: Case, rank: int, world: int, bounds=None):
    """The i-slab of ``case`` owned by ``rank`` as a Case of its own (local indices), with its gh halo columns, plus the
    slab descriptor (ioff, im_global, edges) the device entry points need (include/broadcast_b200.h, bcd_slab_begin)."""
    if world == 1:
        return case, (0, case.im, 0)
    gh, im, jm = case.gh, case.im, case.jm
    if case.periodic_i:
        if bounds is not None:
            raise NotImplementedError("uneven slabs of an i-periodic block")
        return _periodic_slab_of(case, rank, world)
    lo, hi = slab_range(im, rank, world, bounds)
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


def row_window_of(case: Case, la: int, lb: int):
    """Rows la .. lb (1-based, inclusive) of ``case`` as a Case of their own: a window in j.  Its gh ghost rows on a cut side hold
    the parent's real rows (the caller copies them in), the boundary list is clipped to the window and re-expressed in local rows,
    and a window that does not start at the wall runs the ``_nowall`` form of the scheme.  The kernels treat a j-cut like a physical
    side without a fill (sensor gradients of the first ghost row extrapolated), so the residual of the FIRST and LAST rows of a cut
    window differs from the parent's: callers keep a margin of at least one row (resident.RowStreamedBlock) and use the rows
    inside it, which are bit-identical to the parent's."""
    gh, im, jm = case.gh, case.im, case.jm
    if case.periodic_i:
        raise NotImplementedError("row windows of an i-periodic block")
    if not (1 <= la <= lb <= jm):
        raise ValueError("row window out of range")
    n = lb - la + 1
    first, last = la == 1, lb == jm
    cs = (slice(None), slice(la - 1, lb + 2 * gh))          # storage rows of cell rows la-gh .. lb+gh
    ns = (slice(None), slice(la - 1, lb + 2 * gh + 1))
    F = np.asfortranarray
    bcs = []
    for bc in case.bcs:
        kind = bc[0]
        itf = np.array(bc[2], dtype=float)
        if kind == "inflow":           # Ilo, rows 1 .. jm: the window's rows, table rows sliced
            it = itf.copy(); it[0, 1] = 1; it[1, 1] = n
            bcs.append((kind, bc[1], F(it), F(bc[3][la - 1:lb])))
        elif kind == "outflow":        # Ihi, rows 1 .. jm + gh (corner ownership of the top ghost rows)
            it = itf.copy(); it[0, 1] = 1; it[1, 1] = n + gh if last else n
            bcs.append((kind, bc[1], F(it)))
        elif kind == "noref":          # Jhi
            if last:
                it = itf.copy(); it[0, 1] = n; it[1, 1] = n
                bcs.append((kind, bc[1], F(it), bc[3]))
        elif kind == "wall":           # Jlo
            if first:
                bcs.append(bc)
        else:
            raise NotImplementedError(kind)
    scheme = case.scheme if first else case.scheme.replace("_2d", "_nowall_2d") if "nowall" not in case.scheme else case.scheme
    return Case(name=f"{case.name}_rows{la}to{lb}", im=im, jm=n, gh=gh, phys=case.phys, k2=case.k2, k4=case.k4,
                x0=F(case.x0[ns]), y0=F(case.y0[ns]), nx=F(case.nx[ns]), ny=F(case.ny[ns]), xc=F(case.xc[cs]), yc=F(case.yc[cs]),
                vol=F(case.vol[cs]), volf=F(case.volf[cs]), w=F(case.w[cs]), bcs=bcs, periodic_i=False, scheme=scheme)


def _periodic_slab_of(case: Case, rank: int, world: int):
    """i-slab of an i-periodic block (O-mesh, card_cyl2d.py): the join across the cut (jn_match_2d, cylinder.py:499-527) becomes
    the halo exchange between the LAST and the FIRST slab, so every slab has two slab-internal edges (edges = 3) and no join.
    The j-side fills keep their window over the slab's own columns; the exchange follows them (``slab_periodic``), because the
    join they replace copies the ghost rows too.  The metric halo columns are slices of the global arrays, whose ghost columns
    already hold the periodic copies (cases.make_cyl_case)."""
    gh, im, jm = case.gh, case.im, case.jm
    lo, hi = slab_range(im, rank, world)
    n = hi - lo + 1
    if n < 2 * gh + 1:
        raise ValueError(f"slab of {n} columns is narrower than the stencil ({2 * gh + 1}): use fewer ranks")
    cs = slice(lo - 1, hi + 2 * gh)
    ns = slice(lo - 1, hi + 2 * gh + 1)
    F = np.asfortranarray
    bcs = []
    for bc in case.bcs:
        kind = bc[0]
        if kind == "jn":
            continue
        itf = np.array(bc[2], dtype=float)
        if int(itf[0, 0]) != 1 or int(itf[1, 0]) != im or bc[1] not in ("Jlo", "Jhi"):
            raise NotImplementedError("periodic i-slabs: j-side fills over the columns 1 .. im only")
        it = itf.copy(); it[0, 0] = 1; it[1, 0] = n
        if kind == "noref":
            bcs.append((kind, bc[1], F(it), F(bc[3][lo - 1:hi, :])))
        elif kind in ("wall", "symmetry", "antisymmetry"):
            bcs.append((kind, bc[1], F(it)))
        elif kind in ("wall_iso", "pressure"):
            bcs.append((kind, bc[1], F(it)) + tuple(bc[3:]))
        else:
            raise NotImplementedError(kind)
    sl = Case(name=f"{case.name}_slab{rank}of{world}", im=n, jm=jm, gh=gh, phys=case.phys, k2=case.k2, k4=case.k4,
              x0=F(case.x0[ns]), y0=F(case.y0[ns]), nx=F(case.nx[ns]), ny=F(case.ny[ns]), xc=F(case.xc[cs]), yc=F(case.yc[cs]),
              vol=F(case.vol[cs]), volf=F(case.volf[cs]), w=F(case.w[cs]), bcs=bcs, periodic_i=False, scheme=case.scheme,
              slab_periodic=True)
    return sl, (lo - 1, im, 3)


class HaloExchange:
    """neighbour exchange of the gh halo columns of a state held as a torch tensor of shape (planes, jm+2gh, im+2gh)
    (the memory image of the Fortran array (im+2gh, jm+2gh, planes)): NCCL send/recv over NVLink on the GPUs, gloo on CPU.
    All rows (ghost rows included) are exchanged, so corner ghosts of a slab are its neighbour's boundary ghosts."""

    def __init__(self, gh: int, rank: int, world: int, group=None, periodic: bool = False):
        self.gh, self.rank, self.world, self.group = gh, rank, world, group
        self.left = rank - 1 if rank > 0 else (world - 1 if periodic and world > 1 else None)
        self.right = rank + 1 if rank < world - 1 else (0 if periodic and world > 1 else None)
        self._buf = {}

    def _b(self, key, like, shape):
        b = self._buf.get(key)
        if b is None or b.shape != shape or b.device != like.device or b.dtype != like.dtype:
            import torch
            b = torch.empty(shape, dtype=like.dtype, device=like.device)
            self._buf[key] = b
        return b

    def bytes_per_exchange(self, w) -> int:
        sides = (1 if self.left is not None else 0) + (1 if self.right is not None else 0)
        return sides * w.shape[0] * w.shape[1] * self.gh * w.element_size()

    def __call__(self, w):
        if self.world == 1:
            return
        import torch.distributed as dist
        gh = self.gh
        ni = w.shape[2]
        im = ni - 2 * gh
        shape = (w.shape[0], w.shape[1], gh)
        sends, recvs = [], []
        if self.left is not None:
            sl, rl = self._b("sl", w, shape), self._b("rl", w, shape)
            sl.copy_(w[:, :, gh:2 * gh])                 # my first owned columns -> left neighbour's right halo
            sends.append((sl, self.left, 0)); recvs.append((rl, self.left, 1))
        if self.right is not None:
            sr, rr = self._b("sr", w, shape), self._b("rr", w, shape)
            sr.copy_(w[:, :, im:im + gh])                # my last owned columns -> right neighbour's left halo
            sends.append((sr, self.right, 1)); recvs.append((rr, self.right, 0))
        # tags tell the two messages of a pair of ranks apart when the same neighbour sits on both sides (two periodic slabs):
        # a message sent "to the left" (tag 0) is received by its target as coming "from the right" (tag 0)
        # (NCCL ignores tags and matches the messages of a pair in issue order: receive in the order the neighbour sends)
        if self.left is not None and self.left == self.right:
            recvs.reverse()
        ops = [dist.P2POp(dist.isend, b, peer, self.group, tag) for b, peer, tag in sends]
        ops += [dist.P2POp(dist.irecv, b, peer, self.group, tag) for b, peer, tag in recvs]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        if self.left is not None:
            w[:, :, 0:gh].copy_(self._buf["rl"])
        if self.right is not None:
            w[:, :, im + gh:im + 2 * gh].copy_(self._buf["rr"])


class PeerHalo:
    """Halo exchange by PEER STORES over NVLink (csrc/halo.cu): every rank's push kernel writes its gh edge columns straight into
    its neighbours' mailboxes and releases a flag; the unpack kernel waits for its own flags and fills the halo columns of ``w``.
    No NCCL call, no torch op and no host synchronisation on the data path -- two launches per exchange, capturable in a CUDA
    graph (``StepGraph``).  ``torch.distributed`` is used once, at construction, to hand the 64-byte IPC handles around.
    ``periodic``: the first and the last slab are neighbours too (i-periodic O-mesh, cylinder.py:499-527)."""

    def __init__(self, gh: int, rank: int, world: int, w, group=None, periodic: bool = False):
        import ctypes
        import torch.distributed as dist
        from . import _lib
        self.gh, self.rank, self.world = gh, rank, world
        self.lib = _lib.lib()
        self.lib.bcd_halo_mailbox.restype = ctypes.c_void_p
        self.rows = int(w.shape[0] * w.shape[1])
        self.ni = int(w.shape[2])
        self.handle = ctypes.c_void_p(None)
        self.left = rank - 1 if rank > 0 else (world - 1 if periodic and world > 1 else None)
        self.right = rank + 1 if rank < world - 1 else (0 if periodic and world > 1 else None)
        if world == 1:
            return
        mine = (ctypes.c_ubyte * 64)()
        _lib.check(self.lib.bcd_halo_create(ctypes.byref(self.handle), gh, ctypes.c_longlong(self.rows), mine), "bcd_halo_create")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(mine), group=group)
        self.lib.bcd_halo_peer.restype = ctypes.c_void_p
        for side, nb in ((0, self.left), (1, self.right)):
            if nb is None:
                continue
            if side == 1 and nb == self.left:      # two slabs of a periodic block: the same neighbour on both sides, mapped once
                import torch
                _lib.check(self.lib.bcd_halo_connect(self.handle, 1, None, ctypes.c_void_p(self.lib.bcd_halo_peer(self.handle, 0)),
                                                     torch.cuda.current_device()), "bcd_halo_connect")
                continue
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[nb])
            _lib.check(self.lib.bcd_halo_connect(self.handle, side, buf, ctypes.c_void_p(None), -1), "bcd_halo_connect")
        dist.barrier(group=group)      # every mailbox is mapped before the first push

    @classmethod
    def local(cls, ws, gh: int, periodic: bool = False):
        """the halos of several slabs held by ONE process (tensors ``ws`` in slab order, on one or several devices), connected
        through plain device pointers instead of IPC handles.  The exchanges of the slabs must be issued on DIFFERENT streams
        (an unpack spins until its neighbours' pushes have run)."""
        import ctypes
        import torch
        from . import _lib
        L = _lib.lib()
        L.bcd_halo_mailbox.restype = ctypes.c_void_p
        L.bcd_halo_peer.restype = ctypes.c_void_p
        n = len(ws)
        hs = []
        for r, w in enumerate(ws):
            h = cls.__new__(cls)
            h.gh, h.rank, h.world, h.lib = gh, r, n, L
            h.rows, h.ni = int(w.shape[0] * w.shape[1]), int(w.shape[2])
            h.left = r - 1 if r > 0 else (n - 1 if periodic and n > 1 else None)
            h.right = r + 1 if r < n - 1 else (0 if periodic and n > 1 else None)
            h.handle = ctypes.c_void_p(None)
            with torch.cuda.device(w.device):
                _lib.check(L.bcd_halo_create(ctypes.byref(h.handle), gh, ctypes.c_longlong(h.rows), None), "bcd_halo_create")
            hs.append(h)
        for r, (h, w) in enumerate(zip(hs, ws)):
            with torch.cuda.device(w.device):
                for side, nb in ((0, h.left), (1, h.right)):
                    if nb is not None:
                        _lib.check(L.bcd_halo_connect(h.handle, side, None, ctypes.c_void_p(L.bcd_halo_mailbox(hs[nb].handle)),
                                                      ws[nb].device.index), "bcd_halo_connect")
        return hs

    def bytes_per_exchange(self, w) -> int:
        sides = (1 if self.left is not None else 0) + (1 if self.right is not None else 0)
        return sides * self.rows * self.gh * w.element_size()

    def __call__(self, w):
        if self.world == 1:
            return
        import ctypes
        import torch
        from . import _lib
        assert w.shape[0] * w.shape[1] == self.rows and w.shape[2] == self.ni and w.is_contiguous()
        st = ctypes.c_void_p(torch.cuda.current_stream(w.device).cuda_stream)
        _lib.check(self.lib.bcd_halo_exchange(self.handle, ctypes.c_void_p(w.data_ptr()), ctypes.c_longlong(self.rows), self.ni, self.gh, st),
                   "bcd_halo_exchange")

    def error(self) -> int:
        return int(self.lib.bcd_halo_error(self.handle)) if self.world > 1 else 0

    def close(self):
        if self.handle:
            self.lib.bcd_halo_destroy(self.handle)
            self.handle = None


class StepGraph:
    """The per-step sequence [halo exchange, boundary fills, residual] of one Block captured ONCE into a CUDA graph
    (bcd_graph_begin / _end) and replayed with one host call per step: at 8 GPUs the step is bounded by launch latency, not by
    device time (72 us of Python-issued copies and launches on a 315 us kernel)."""

    def __init__(self, blk, halo=None):
        import ctypes
        import torch
        from . import _lib
        self.blk, self.halo, self.lib = blk, halo, _lib.lib()
        self.exec = ctypes.c_void_p(None)

        after = bool(getattr(blk.case, "slab_periodic", False))   # periodic slabs: the exchange replaces the join, which follows the fills

        # BROADCAST_B200_STEP_OVERLAP=1: the step forks inside the graph -- [exchange, boundary fills] on a side stream WHILE the inner
        # tiles of the residual (which read no ghost cell and no halo column) run on the bulk kernel; the ring of tiles follows both.
        # Measured on 8 B200 (profiles/r2_c_summary.md): 0.3072 ms against 0.3062 ms for the plain sequence, same checksum -- the
        # ~38 us a step takes beyond its residual kernel are node-to-node latencies and the kernel's own ramp / tail, not the
        # exchange; what the fork hides the extra ring launch costs again.  Off by default.
        overlap = halo is not None and os.environ.get("BROADCAST_B200_STEP_OVERLAP", "0") == "1"
        self.overlap = overlap
        side = torch.cuda.Stream(device=blk.device) if overlap else None

        def seq():
            if overlap:
                main = torch.cuda.current_stream(blk.device)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    if not after:
                        halo(blk.w)
                    blk.apply_bcs()
                    if after:
                        halo(blk.w)
                blk.residual_part(1)
                main.wait_stream(side)
                blk.residual_part(2)
                return
            if halo is not None and not after:
                halo(blk.w)
            blk.apply_bcs()
            if halo is not None and after:
                halo(blk.w)
            blk.residual()
        self._seq = seq
        seq()                                   # scratch buffers, kernel attributes, descriptor caches: not capturable
        torch.cuda.current_stream(blk.device).synchronize()
        # a capture cannot run on the legacy default stream: the graph is captured (and replayed) on the caller's current stream
        # when that is a real stream, else on a stream of its own that is ordered with the default stream around every replay
        cur = torch.cuda.current_stream(blk.device)
        self.own = torch.cuda.Stream(device=blk.device) if cur.cuda_stream == 0 else None
        with torch.cuda.stream(self.own if self.own is not None else cur):
            st = blk._stream()
            _lib.check(self.lib.bcd_graph_begin(st), "bcd_graph_begin")
            try:
                seq()
            finally:
                rc = self.lib.bcd_graph_end(st, ctypes.byref(self.exec))
            _lib.check(rc, "bcd_graph_end")

    def __call__(self):
        import ctypes
        import torch
        from . import _lib
        cur = torch.cuda.current_stream(self.blk.device)
        if cur.cuda_stream == 0:
            if self.own is None:
                self.own = torch.cuda.Stream(device=self.blk.device)
            self.own.wait_stream(cur)
            _lib.check(self.lib.bcd_graph_launch(self.exec, ctypes.c_void_p(self.own.cuda_stream)), "bcd_graph_launch")
            cur.wait_stream(self.own)
        else:
            _lib.check(self.lib.bcd_graph_launch(self.exec, ctypes.c_void_p(cur.cuda_stream)), "bcd_graph_launch")

    def close(self):
        if self.exec:
            self.lib.bcd_graph_destroy(self.exec)
            self.exec = None


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
