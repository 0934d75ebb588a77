"""Device-resident mode: one structured block held in HBM (torch tensors are only the memory
holders), driven through the ``bcd_*`` device-pointer entry points of the C ABI.

A ``Block`` mirrors what the reference drivers keep in their NPZ tree for one run
(BROADCAST_npz.py:247-262): geometry, state, boundary-condition tables and physics scalars; its
methods are the driver steps of BROADCAST_npz.py:1011-1137 / cylinder.py:841-985 without host
round trips.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .cases import Case


def _t(a: np.ndarray, device) -> torch.Tensor:
    """numpy Fortran-ordered (ni,nj[,np]) -> torch tensor (np,nj,ni) with identical memory layout"""
    a = np.asfortranarray(a, dtype=np.float64)
    return torch.from_numpy(np.ascontiguousarray(a.T)).to(device)


def to_numpy(t: torch.Tensor) -> np.ndarray:
    return np.asfortranarray(t.detach().cpu().numpy().T)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(None)


def _interf(a) -> np.ndarray:
    a = np.asarray(a)
    return np.array([a[0, 0], a[0, 1], a[1, 0], a[1, 1]], dtype=np.int32)


class Block:
    def __init__(self, case: Case, device="cuda:0", slab=None):
        """``slab`` = (ioff, im_global, edges) when ``case`` is an i-slab of a larger block (broadcast_b200.sharding)."""
        if not torch.cuda.is_available():
            raise _lib.BroadcastB200Error("resident mode needs a CUDA device (no CPU fallback)")
        self.lib = _lib.lib()
        self.slab = tuple(int(v) for v in slab) if slab is not None else None
        self.ioff, self.im_global = (self.slab[0], self.slab[1]) if self.slab else (0, case.im)
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.case = case
        self.im, self.jm, self.gh = case.im, case.jm, case.gh
        self.nx, self.ny = _t(case.nx, device), _t(case.ny, device)
        self.vol, self.volf = _t(case.vol, device), _t(case.volf, device)
        self.w = _t(case.w, device)
        self.res = torch.zeros_like(self.w)
        self.out10 = torch.zeros(16, dtype=torch.float64, device=device)
        self.wall = 0 if "nowall" in case.scheme else 1
        p = case.phys
        self._phys = [ctypes.c_double(float(v)) for v in (p["cp"], p["cv"], p["prandtl"], p["gam"], p["rgaz"], p["cs"], p["muref"],
                                                          p["tref"], p["cs"], case.k2, case.k4)]
        self.gam = float(p["gam"])
        # boundary tables
        self.bcs = []
        for bc in case.bcs:
            kind = bc[0]
            if kind == "inflow":
                fld = np.asfortranarray(bc[3])
                self.bcs.append(("inflow", bc[1].encode(), _interf(bc[2]), torch.from_numpy(np.ascontiguousarray(fld.T)).to(device),
                                 fld.shape[0]))
            elif kind == "noref":
                wbd = np.asfortranarray(bc[3])
                self.bcs.append(("noref", bc[1].encode(), _interf(bc[2]), torch.from_numpy(np.ascontiguousarray(wbd.T)).to(device),
                                 wbd.shape[0]))
            elif kind in ("outflow", "wall", "symmetry", "antisymmetry"):
                self.bcs.append((kind, bc[1].encode(), _interf(bc[2])))
            elif kind in ("wall_blow_profile", "wall_iso_profile"):   # (kind, loc, interf, profile[, rgaz]); profile passive
                prof = np.ascontiguousarray(np.asarray(bc[3], dtype=np.float64).ravel())
                self.bcs.append((kind, bc[1].encode(), _interf(bc[2]), torch.from_numpy(prof).to(device), prof.size,
                                 float(bc[4]) if len(bc) > 4 else 0.0))
            elif kind in ("wall_iso", "pressure"):     # (kind, loc, interf, twall, rgaz) / (kind, loc, interf, pext, noref)
                self.bcs.append((kind, bc[1].encode(), _interf(bc[2]), float(bc[3]), float(bc[4])))
            elif kind == "jn":
                self.bcs.append(("jn", [(_interf(prr), _interf(prd), np.asarray(tr, dtype=np.int32)) for prr, prd, tr in bc[1:]]))
            else:
                raise ValueError(kind)

    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc, what):
        _lib.check(rc, what)

    def call(self, name, *args):
        """one device entry point, inside this block's slab context (bcd_slab_begin / bcd_slab_end)"""
        fn = getattr(self.lib, name)
        if self.slab is None:
            rc = fn(*args)
        else:
            _lib.check(self.lib.bcd_slab_begin(*self.slab), "bcd_slab_begin")
            try:
                rc = fn(*args)
            finally:
                self.lib.bcd_slab_end()
        _lib.check(rc, name)

    def upload_state(self, w_host: np.ndarray):
        self.w.copy_(torch.from_numpy(np.ascontiguousarray(np.asfortranarray(w_host).T)))

    def apply_bcs(self, w=None, wd=None, ndir=0):
        """ghost fill in driver order; with ``wd`` ([ndir,5,nj,ni]) the linearised fills (w and wd ghosts)."""
        L, st = self.lib, self._stream()
        im, jm, gh = self.im, self.jm, self.gh
        if w is None and wd is None:
            # primal fill of the block's own state: the whole list in ONE call (bcd_apply_bcs), descriptors built once
            d = self.__dict__.get("_bc_desc_cache")
            if d is None:
                d = self._bc_desc_cache = _bc_descs(self)
            self.call("bcd_apply_bcs", _p(self.w), _p(self.nx), _p(self.ny), ctypes.c_double(self.gam), gh, im, jm, d[0], d[1], st)
            return
        w = self.w if w is None else w
        I = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        for bc in self.bcs:
            kind = bc[0]
            if kind == "inflow":
                self._ck(L.bcd_bc_supandsubinlet(_p(w), _p(wd), ndir, bc[1], I(bc[2]), _p(bc[3]), _p(self.nx), _p(self.ny),
                                                 ctypes.c_double(self.gam), im, jm, bc[4], gh, st), "bcd_bc_supandsubinlet")
            elif kind == "noref":
                self._ck(L.bcd_bc_no_reflexion(_p(w), _p(wd), ndir, _p(bc[3]), bc[1], I(bc[2]), _p(self.nx), _p(self.ny),
                                               ctypes.c_double(self.gam), gh, im, jm, bc[4], st), "bcd_bc_no_reflexion")
            elif kind == "outflow":
                self._ck(L.bcd_bc_extrapolate_o2(_p(w), _p(wd), ndir, bc[1], I(bc[2]), im, jm, gh, st), "bcd_bc_extrapolate_o2")
            elif kind == "wall":
                self._ck(L.bcd_bc_wall_viscous_adia(_p(w), _p(wd), ndir, bc[1], ctypes.c_double(self.gam), I(bc[2]), gh, im, jm, st),
                         "bcd_bc_wall_viscous_adia")
            elif kind == "wall_iso":
                self._ck(L.bcd_bc_wall_viscous_iso(_p(w), _p(wd), ndir, ctypes.c_double(bc[3]), bc[1], ctypes.c_double(self.gam),
                                                   ctypes.c_double(bc[4]), I(bc[2]), gh, im, jm, st), "bcd_bc_wall_viscous_iso")
            elif kind in ("symmetry", "antisymmetry"):
                self._ck(L.bcd_bc_symmetry(_p(w), _p(wd), ndir, bc[1], I(bc[2]), _p(self.nx), _p(self.ny), gh, im, jm,
                                           int(kind == "antisymmetry"), st), "bcd_bc_symmetry")
            elif kind in ("wall_blow_profile", "wall_iso_profile"):
                self._ck(L.bcd_bc_wall_profile(_p(w), _p(wd), ndir, int(kind == "wall_blow_profile"), _p(bc[3]), ctypes.c_void_p(None), bc[1],
                                               ctypes.c_double(self.gam), ctypes.c_double(0.0), ctypes.c_double(bc[5]), ctypes.c_double(0.0),
                                               I(bc[2]), gh, im, jm, bc[4], st), "bcd_bc_wall_profile")
            elif kind == "pressure":
                self._ck(L.bcd_bc_pressure(_p(w), _p(wd), ndir, bc[1], I(bc[2]), ctypes.c_double(bc[3]), int(bc[4] != 0.0),
                                           ctypes.c_double(self.gam), _p(self.nx), _p(self.ny), im, jm, gh, st), "bcd_bc_pressure")
            elif kind == "jn":
                targets = [w] if wd is None else [wd, w]
                for t in targets:
                    em = 5 * (ndir if (t is wd and ndir > 1) else 1)
                    for prr, prd, tr in bc[1]:
                        self._ck(L.bcd_jn_match(_p(t), I(prr), gh, gh, gh, gh, im, jm, _p(t), I(prd), gh, gh, gh, gh, im, jm, I(tr), em,
                                                st), "bcd_jn_match")

    def residual(self, generic=False, w=None, out=None, variant=None):
        """variant: 0 default (the 32 x 9 tile kernel k_residual_fast), 1 generic four-kernel pipeline, 2 tile kernel + TMA/persistent,
        3 first-generation tile kernel, 4 = 0 by name, 5 j-marching persistent kernel k_residual_march (rings fed by TMA / bulk copies)"""
        w = self.w if w is None else w
        out = self.res if out is None else out
        v = int(variant) if variant is not None else (1 if generic else 0)
        self.call("bcd_residual", _p(out), _p(w), _p(self.nx), _p(self.ny), _p(self.vol), _p(self.volf), self.gh, *self._phys,
                  self.im, self.jm, self.wall, v, self._stream())
        return out

    def residual_part(self, part, stream=None):
        """part 1 = inner tiles (no ghost / halo reads), part 2 = the outer ring of tiles, 0 = all (bcd_residual_part)"""
        st = ctypes.c_void_p(stream.cuda_stream) if stream is not None else self._stream()
        self.call("bcd_residual_part", _p(self.res), _p(self.w), _p(self.nx), _p(self.ny), _p(self.vol), _p(self.volf), self.gh,
                  *self._phys, self.im, self.jm, self.wall, int(part), st)
        return self.res

    def step_overlapped(self, halo=None):
        """halo exchange + boundary fills on a side stream WHILE the inner tiles of the residual run; the ring of tiles that
        reads ghosts / halo columns follows both.  Same result as halo(); step(), bit for bit."""
        main = torch.cuda.current_stream(self.device)
        side = self.__dict__.setdefault("_side_stream", torch.cuda.Stream(device=self.device))
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if halo is not None:
                halo(self.w)
            self.apply_bcs()
        self.residual_part(1)
        main.wait_stream(side)
        return self.residual_part(2)

    def tangent(self, wd, ndir, out, w=None, rect=None):
        w = self.w if w is None else w
        r = None
        if rect is not None:
            r = np.asarray(rect, dtype=np.int32)
        self.call("bcd_tangent", _p(out), _p(w), _p(wd), ndir, _p(self.nx), _p(self.ny), _p(self.vol), _p(self.volf), self.gh,
                  *self._phys, self.im, self.jm, self.wall,
                  r.ctypes.data_as(ctypes.c_void_p) if r is not None else ctypes.c_void_p(None), self._stream())
        return out

    def dz_tangent(self, wd0, wd, out1=None, out2=None, w=None, rect=None):
        """tangent of the Dz (out1) and Dz2 (out2) operator rows w.r.t. the base flow, both in ONE pass over device arrays
        (f_lindz.coeffs_5p_dz_d / coeffs_5p_dz2_d, BROADCAST_npz_sens.py:1768-1769); wd0 = base-flow variation, wd = mode"""
        w = self.w if w is None else w
        r = np.asarray(rect, dtype=np.int32) if rect is not None else None
        null = ctypes.c_void_p(None)
        self.call("bcd_dz_tangent", _p(out1) if out1 is not None else null, _p(out2) if out2 is not None else null, _p(w), _p(wd0),
                  _p(wd), _p(self.nx), _p(self.ny), _p(self.vol), self.gh, *self._phys[:9], self.im, self.jm,
                  r.ctypes.data_as(ctypes.c_void_p) if r is not None else null, self._stream())
        return out1, out2

    def norms(self, res=None, reduce=None):
        res = self.res if res is None else res
        self._ck(self.lib.bcd_norm_sums(_p(self.out10), _p(res), self.im, self.jm, self.gh, self._stream()), "bcd_norm_sums")
        if reduce is not None:   # slabs: sum of squares / 10th powers over the ranks (one all-reduce of 10 doubles)
            reduce(self.out10)
        h = self.out10.cpu().numpy()
        return np.sqrt(h[:5]), h[5:10] ** 0.1

    def step(self):
        """one explicit-stage evaluation: boundary fill then residual (BROADCAST_npz.py:854-875)"""
        self.apply_bcs()
        return self.residual()

    def step_from_host(self, w_pinned: torch.Tensor, res_pinned: torch.Tensor, halo=None):
        """plugin-level step with HOST buffers: H2D of the state, (halo exchange,) boundary fill + residual on the
        device, D2H of the residual.  Geometry and BC tables stay resident (they belong to the mesh)."""
        self.w.copy_(w_pinned, non_blocking=True)
        if halo is not None:
            halo(self.w)
        self.step()
        res_pinned.copy_(self.res, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()


class StreamedBlock:
    """Plugin-level residual step on HOST buffers, pipelined over ``nslab`` i-slabs of the block on ONE GPU: the pitched
    host-to-device copy of slab k+1 (its columns plus gh halo columns straight from the host array) and the
    device-to-host copy of slab k-1's residual overlap the boundary fill + residual kernels of slab k (three streams,
    both copy engines busy).  Same results as ``Block.step_from_host`` (slab-internal edges compute their gradients)."""

    @staticmethod
    def tapered_bounds(im: int, nslab: int, unit: int = 32):
        """column boundaries of ``nslab`` slabs whose widths double from both ends towards the middle (multiples of ``unit``): the
        step then starts its first device-to-host copy after a SHORT first upload and ends with a SHORT last download -- with even
        slabs the pipeline costs (nslab + 1) slab transfers, one of them with only one direction of the link busy at each end"""
        half = nslab // 2
        w = [2 ** min(k, nslab - 1 - k) for k in range(nslab)]
        # cap the doubling so that the widest slabs are at most ~4x the even width (device memory per slab, pipeline granularity)
        cap = max(1, 4 * sum(w) // nslab) if nslab > 2 else max(w)
        w = [min(x, cap) for x in w]
        tot = sum(w)
        cols = [max(unit, int(round(im * x / tot / unit)) * unit) for x in w]
        cols[half] += im - sum(cols)                      # the remainder goes to a middle slab
        if cols[half] < unit or any(c < 2 * 3 + 1 for c in cols):
            return None
        b = [0]
        for c in cols:
            b.append(b[-1] + c)
        return b

    def __init__(self, case: Case, nslab: int = 8, device="cuda:0", first: int = 0, count: int = None, bounds=None):
        """``first`` / ``count``: this object drives only the slabs first .. first+count-1 of the ``nslab`` slabs of ``case`` (one
        rank of a multi-GPU run pipelining ITS part of the block: the host arrays then hold the columns of those slabs plus gh
        halo columns on each side, i.e. the rank's own slab image).  ``bounds``: uneven slab boundaries (``tapered_bounds``)."""
        from . import sharding
        self.case, self.device = case, torch.device(device)
        self.gh, self.im, self.jm = case.gh, case.im, case.jm
        count = nslab - first if count is None else count
        self.col0 = sharding.slab_range(case.im, first, nslab, bounds)[0] - 1     # storage column of the host arrays' first column
        self.slabs = []
        for k in range(first, first + count):
            sl, desc = sharding.slab_of(case, k, nslab, bounds)
            lo, hi = sharding.slab_range(case.im, k, nslab, bounds)
            self.slabs.append((Block(sl, device, slab=desc if nslab > 1 else None), lo - self.col0, hi - self.col0))
        # one stream per ROLE (host-to-device copies, kernels, device-to-host copies) and one event pair per slab: the H2D queue
        # never waits behind a D2H copy, both copy engines stay busy for the whole step (measured: 45.9 GB/s each way at once)
        self.s_in, self.s_k, self.s_out = (torch.cuda.Stream(device=self.device) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(nslab)]
        self.ev_k = [torch.cuda.Event() for _ in range(nslab)]
        self.lib = _lib.lib()

    def bytes_per_step(self):
        nj = self.jm + 2 * self.gh
        h2d = sum((b.im + 2 * self.gh) * nj * 5 * 8 for b, _, _ in self.slabs)
        d2h = sum(b.im * nj * 5 * 8 for b, _, _ in self.slabs)
        return h2d, d2h

    def step_from_host(self, w_pinned: torch.Tensor, res_pinned: torch.Tensor):
        """``w_pinned`` / ``res_pinned``: pinned host tensors (5, jm+2gh, im+2gh) = memory image of the Fortran arrays"""
        gh, nj, ni = self.gh, self.jm + 2 * self.gh, int(w_pinned.shape[2])
        rows = 5 * nj
        main = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_k, self.s_out):
            s.wait_stream(main)
        LL, VP = ctypes.c_longlong, ctypes.c_void_p
        for k, (b, lo, hi) in enumerate(self.slabs):      # all host-to-device copies, back to back
            nl = b.im + 2 * gh
            src = w_pinned.data_ptr() + (lo - 1) * 8                          # storage column of cell lo-gh
            _lib.check(self.lib.bcd_memcpy2d(_p(b.w), LL(nl * 8), VP(src), LL(ni * 8), LL(nl * 8), LL(rows), 1, VP(self.s_in.cuda_stream)),
                       "bcd_memcpy2d")
            self.ev_in[k].record(self.s_in)
        for k, (b, lo, hi) in enumerate(self.slabs):
            # kernels of slab k as soon as its state has landed ...
            self.s_k.wait_event(self.ev_in[k])
            with torch.cuda.stream(self.s_k):
                b.apply_bcs()
                b.residual()
            self.ev_k[k].record(self.s_k)
            # ... and its residual back as soon as it exists (issued right away: the host needs ~0.2 ms per slab to issue the
            # kernels, a device-to-host copy issued after ALL kernels would start milliseconds late)
            nl = b.im + 2 * gh
            self.s_out.wait_event(self.ev_k[k])
            dst = res_pinned.data_ptr() + (lo - 1 + gh) * 8                   # owned columns only
            _lib.check(self.lib.bcd_memcpy2d(VP(dst), LL(ni * 8), VP(b.res.data_ptr() + gh * 8), LL(nl * 8), LL(b.im * 8), LL(rows), 2,
                                             VP(self.s_out.cuda_stream)), "bcd_memcpy2d")
        main.wait_stream(self.s_out)
        main.synchronize()


class RowStreamedBlock:
    """Plugin-level residual step on HOST buffers pipelined over row windows of the block: ``nwin`` windows in j, so that every
    host-link transfer is CONTIGUOUS -- rows are the slow index of the Fortran planes: a window is one run of memory per plane, five
    runs per copy, where an i-slab is 10 270 pitched rows of 8 KB.  Measured on B200 (profiles/r2_c_summary.md): with both directions
    busy the pitched copies of ``StreamedBlock`` move 29.6 GB/s each way, contiguous ones 43 GB/s.
    Window k owns the output rows [a, b]; it is computed on the rows [a - M, b + M] (clipped to the block; M = ``margin``) plus gh
    ghost rows that arrive with the same copy, as a block of its
    own (``sharding.row_window_of``: clipped boundary list, ``_nowall`` scheme away from the wall).  Only the owned rows travel back.
    Same result as ``Block.step_from_host``, bit for bit (tests/test_parity_gpu.py)."""

    def __init__(self, case: Case, nwin: int = 8, device="cuda:0", margin: int = 2, taper: bool = False):
        """``margin``: rows computed beyond the owned ones on a cut side: at least TWO (the sensor gradient of a cut window's first
        ghost row is extrapolated, which reaches the two outermost rows: a margin of one row is NOT bit-identical on the GPU,
        margins 2 and gh + 2 are, tests/test_parity_gpu.py).  ``taper``: window heights doubling from both ends towards the middle
        (short first upload, short last download); measured slower than even windows at C5 (18.5 vs 17.7 ms: the thin end windows
        cost more than the shorter fill and drain save), so off by default."""
        from . import sharding
        self.case, self.device = case, torch.device(device)
        self.gh, self.im, self.jm = case.gh, case.im, case.jm
        M = int(margin)
        if M < 2:
            raise ValueError("a cut window needs a margin of at least two rows")
        bounds = StreamedBlock.tapered_bounds(case.jm, nwin, unit=8) if (taper and nwin > 2) else None
        self.wins = []
        for k in range(nwin):
            a, b = sharding.slab_range(case.jm, k, nwin, bounds)
            la, lb = max(1, a - M), min(case.jm, b + M)
            self.wins.append((Block(sharding.row_window_of(case, la, lb), device), a, b, la, lb))
        self.s_in, self.s_k, self.s_out = (torch.cuda.Stream(device=self.device) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(nwin)]
        self.ev_k = [torch.cuda.Event() for _ in range(nwin)]
        self.lib = _lib.lib()

    def bytes_per_step(self):
        ni = self.im + 2 * self.gh
        h2d = sum((lb - la + 1 + 2 * self.gh) * ni * 5 * 8 for _, _, _, la, lb in self.wins)
        d2h = sum((b - a + 1) * ni * 5 * 8 for _, a, b, _, _ in self.wins)
        return h2d, d2h

    def step_from_host(self, w_pinned: torch.Tensor, res_pinned: torch.Tensor):
        """``w_pinned`` / ``res_pinned``: pinned host tensors (5, jm+2gh, im+2gh) = memory image of the Fortran arrays.  The rows
        1 .. jm of ``res_pinned`` are written (all columns; the ghost columns receive the zeros of the device buffer)."""
        gh, nj, ni = self.gh, self.jm + 2 * self.gh, self.im + 2 * self.gh
        main = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_k, self.s_out):
            s.wait_stream(main)
        LL, VP = ctypes.c_longlong, ctypes.c_void_p
        plane = nj * ni * 8
        for k, (blk, a, b, la, lb) in enumerate(self.wins):      # all host-to-device copies, back to back
            nr = lb - la + 1 + 2 * gh                           # storage rows la-1 .. lb+2gh-1 of every plane: one run of memory
            src = w_pinned.data_ptr() + (la - 1) * ni * 8
            _lib.check(self.lib.bcd_memcpy2d(_p(blk.w), LL(nr * ni * 8), VP(src), LL(plane), LL(nr * ni * 8), LL(5), 1,
                                             VP(self.s_in.cuda_stream)), "bcd_memcpy2d")
            self.ev_in[k].record(self.s_in)
        for k, (blk, a, b, la, lb) in enumerate(self.wins):
            self.s_k.wait_event(self.ev_in[k])
            with torch.cuda.stream(self.s_k):
                blk.apply_bcs()
                blk.residual()
            self.ev_k[k].record(self.s_k)
            self.s_out.wait_event(self.ev_k[k])
            nr = lb - la + 1 + 2 * gh
            src = blk.res.data_ptr() + (a - la + gh) * ni * 8                 # owned rows only
            dst = res_pinned.data_ptr() + (a - 1 + gh) * ni * 8
            _lib.check(self.lib.bcd_memcpy2d(VP(dst), LL(plane), VP(src), LL(nr * ni * 8), LL((b - a + 1) * ni * 8), LL(5), 2,
                                             VP(self.s_out.cuda_stream)), "bcd_memcpy2d")
        main.wait_stream(self.s_out)
        main.synchronize()


def local_halo_exchange(blocks):
    """halo exchange between the i-slab Blocks of ONE process (rank order = list order): device-to-device copies of the
    gh columns next to every slab-internal edge (peer copies over NVLink when the blocks live on different GPUs)."""
    for a, b in zip(blocks[:-1], blocks[1:]):
        gh = a.gh
        b.w[:, :, 0:gh].copy_(a.w[:, :, a.im:a.im + gh], non_blocking=True)                    # a's last owned -> b's left halo
        a.w[:, :, a.im + gh:a.im + 2 * gh].copy_(b.w[:, :, gh:2 * gh], non_blocking=True)      # b's first owned -> a's right halo


# ----------------------------------------------------------------------------------------------
# whole colour loop on the device
# ----------------------------------------------------------------------------------------------
class _BcDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("loc", ctypes.c_char * 4), ("window", ctypes.c_int32 * 4), ("prd", ctypes.c_int32 * 4),
                ("tr", ctypes.c_int32 * 2), ("lm", ctypes.c_int32), ("table", ctypes.c_void_p), ("param", ctypes.c_double * 2)]


_KIND = {"inflow": 1, "noref": 2, "outflow": 3, "wall": 4, "jn": 5, "wall_iso": 6, "symmetry": 7, "antisymmetry": 8, "pressure": 9, "wall_blow_profile": 10,
         "wall_iso_profile": 11}
SCATTER = {"jv": 0, "jv_relaxed": 1, "dz": 2, "jv_relaxed_withjn": 3, "jv_withjn": 4, "jv_dbyvol": 5, "jv_relaxed_dbyvol": 6}


def _bc_descs(blk: "Block"):
    out = []
    for bc in blk.bcs:
        kind = bc[0]
        if kind == "jn":
            for prr, prd, tr in bc[1]:
                d = _BcDesc()
                d.kind = 5
                d.window[:] = [int(v) for v in prr]
                d.prd[:] = [int(v) for v in prd]
                d.tr[:] = [int(v) for v in tr]
                out.append(d)
        else:
            d = _BcDesc()
            d.kind = _KIND[kind]
            d.loc = bc[1]
            d.window[:] = [int(v) for v in bc[2]]
            if kind in ("inflow", "noref"):
                d.table = bc[3].data_ptr()
                d.lm = int(bc[4])
            elif kind in ("wall_iso", "pressure"):
                d.param[:] = [bc[3], bc[4]]
            elif kind in ("wall_blow_profile", "wall_iso_profile"):
                d.table = bc[3].data_ptr()
                d.lm = int(bc[4])
                d.param[:] = [0.0, bc[5]]
            out.append(d)
    arr = (_BcDesc * len(out))(*out)
    return arr, len(out)


def jacobian_coo(blk: "Block", coefdiag=None, kind=None, rect=None, out=None, compact=False, colours=None):
    """The reference's COO lists (Jac, IA, JA; slot order of misc/ComputeJacobian.f90:524) assembled
    on the device by the colour loop of BROADCAST_npz.py:1068-1127 with 5 directions per pass.
    ``blk.w`` must hold the state with its ghosts filled (``blk.apply_bcs()``).
    ``colours`` = (c0, c1): colour sharding, only the passes c0 <= l*(2gh+1)+k < c1 (sharding.colour_range); the slots
    of the other colours stay zero."""
    im, jm, gh = blk.im, blk.jm, blk.gh
    s = 2 * gh + 1
    nb = 25 * s * s * im * jm
    if rect is not None and compact:
        nb = 25 * s * s * max(0, rect[1] - rect[0] + 1) * max(0, rect[3] - rect[2] + 1)
    if nb >= 2 ** 31:
        raise _lib.BroadcastB200Error("the reference COO layout overflows 32-bit slots at this size; use the block-CSR assembly")
    if kind is None:
        kind = "jv_relaxed_withjn" if blk.case.periodic_i else "jv_relaxed"
    if out is None:
        jac = torch.zeros(nb, dtype=torch.float64, device=blk.device)
        ia = torch.zeros(nb, dtype=torch.int32, device=blk.device)
        ja = torch.zeros(nb, dtype=torch.int32, device=blk.device)
    else:
        jac, ia, ja = out
    if coefdiag is None and "relaxed" in kind:
        coefdiag = torch.zeros((jm, im), dtype=torch.float64, device=blk.device)
    elif coefdiag is not None and not isinstance(coefdiag, torch.Tensor):
        coefdiag = _t(coefdiag, blk.device)
    descs, n = _bc_descs(blk)
    r = np.asarray(rect, dtype=np.int32) if rect is not None else None
    if colours is not None:
        _lib.check(_lib.lib().bcd_colour_range(int(colours[0]), int(colours[1])), "bcd_colour_range")
    try:
        blk.call("bcd_jacobian_coo", _p(jac), _p(ia), _p(ja), _p(blk.w), _p(blk.nx), _p(blk.ny), _p(blk.vol), _p(blk.volf), gh, *blk._phys,
                 im, jm, blk.wall, descs, n, SCATTER[kind], _p(coefdiag),
                 r.ctypes.data_as(ctypes.c_void_p) if r is not None else ctypes.c_void_p(None), 1 if compact else 0, blk._stream())
    finally:
        if colours is not None:
            _lib.lib().bcd_colour_range(0, 0)
    return jac, ia, ja


def dz_coo(blk: "Block", which=(1, 2)):
    """COO lists of the spanwise operators Dz (which 1) and Dz2 (which 2) by the colour loop of
    BROADCAST_npz.py:1231-1246 on the device.  Returns {1: (jac, ia, ja), 2: (jac, ia, ja)}."""
    im, jm, gh = blk.im, blk.jm, blk.gh
    s = 2 * gh + 1
    nb = 25 * s * s * im * jm
    if nb >= 2 ** 31:
        raise _lib.BroadcastB200Error("the reference COO layout overflows 32-bit slots at this size")
    out = {}
    for wh in (1, 2):
        if wh in which:
            out[wh] = (torch.zeros(nb, dtype=torch.float64, device=blk.device), torch.zeros(nb, dtype=torch.int32, device=blk.device),
                       torch.zeros(nb, dtype=torch.int32, device=blk.device))
    descs, n = _bc_descs(blk)
    a = [(_p(t) for t in out[wh]) if wh in out else (ctypes.c_void_p(None),) * 3 for wh in (1, 2)]
    args = [x for trip in a for x in trip]
    blk.call("bcd_dz_coo", *args, _p(blk.w), _p(blk.nx), _p(blk.ny), _p(blk.vol), _p(blk.volf), gh, *blk._phys[:9], im, jm, descs, n,
             blk._stream())
    return out


def dz_tangent_coo(blk: "Block", wmoder, wmodei=None, which=(1, 2)):
    """COO lists of d/dw [Dz(w) mode] (which 1) and d/dw [Dz2(w) mode] (which 2) by the colour loop of the sensitivity driver
    (BROADCAST_npz_sens.py:1741-1800) on the device.  Returns {1: (jac_r, jac_i or None, ia, ja), 2: ...}."""
    im, jm, gh = blk.im, blk.jm, blk.gh
    s = 2 * gh + 1
    nb = 25 * s * s * im * jm
    if nb >= 2 ** 31:
        raise _lib.BroadcastB200Error("the reference COO layout overflows 32-bit slots at this size")
    null = ctypes.c_void_p(None)
    out, args = {}, []
    for wh in (1, 2):
        if wh in which:
            jr = torch.zeros(nb, dtype=torch.float64, device=blk.device)
            ji = torch.zeros(nb, dtype=torch.float64, device=blk.device) if wmodei is not None else None
            ia, ja = (torch.zeros(nb, dtype=torch.int32, device=blk.device) for _ in range(2))
            out[wh] = (jr, ji, ia, ja)
            args += [_p(jr), _p(ji) if ji is not None else null, _p(ia), _p(ja)]
        else:
            args += [null] * 4
    descs, n = _bc_descs(blk)
    blk.call("bcd_dz_tangent_coo", *args, _p(blk.w), _p(wmoder), _p(wmodei) if wmodei is not None else null, _p(blk.nx), _p(blk.ny),
             _p(blk.vol), gh, *blk._phys[:9], im, jm, descs, n, blk._stream())
    return out


def remove_zero_jac(jac, ia, ja, thresh=2e-16):
    """BROADCAST_npz.py:129-135 on the device"""
    keep = jac.abs() > thresh
    return jac[keep], ia[keep], ja[keep]


# ----------------------------------------------------------------------------------------------
# hybrid assembly: direct block kernels on the regular interior + colour loop on the boundary strips
# ----------------------------------------------------------------------------------------------
def jacobian_slots(blk: "Block"):
    off = np.zeros((29, 2), dtype=np.int32)
    n = blk.lib.bcd_jacobian_slots(off.ctypes.data_as(ctypes.c_void_p))
    return off[:n]


class HybridJacobian:
    """fixed-pattern block Jacobian: ``blocks[slot, e, m, j-1, i-1]`` for the regular interior rows
    (gh+1..im-gh x gh+1..jm-gh; other rows of ``blocks`` are unused) and reference-ordered COO lists for
    the four boundary strips."""

    def __init__(self, blk, blocks, offsets, strips, region, strip_rects=(), counts=None, count_thresh=None):
        self.blk, self.blocks, self.offsets, self.strips, self.region = blk, blocks, offsets, strips, region
        self.strip_rects = list(strip_rects)
        # per-row counts of the regular rows taken by the assembly kernel itself (jacobian_hybrid(count_thresh=...))
        self.counts, self.count_thresh = counts, count_thresh

    def to_coo(self, thresh=2e-16):
        """filtered COO (remove_zero_jac semantics) on the device, reference numbering of rows/columns"""
        blk = self.blk
        im, jm = blk.im, blk.jm
        i0, i1, j0, j1 = self.region
        vals, rows, cols = [], [], []
        if i1 >= i0 and j1 >= j0:
            dev = blk.device
            I = torch.arange(i0, i1 + 1, device=dev, dtype=torch.int64)[None, :] + blk.ioff   # global column index of the local cells
            J = torch.arange(j0, j1 + 1, device=dev, dtype=torch.int64)[:, None]
            for s, (di, dj) in enumerate(self.offsets):
                for e in range(5):
                    ia = (e + 5 * (J - 1) + 5 * jm * (I - 1)).expand(j1 - j0 + 1, i1 - i0 + 1)
                    for m in range(5):
                        v = self.blocks[s, e, m, j0 - 1:j1, i0 - 1:i1]
                        keep = v.abs() > thresh
                        ja = (m + 5 * (J + int(dj) - 1) + 5 * jm * (I + int(di) - 1)).expand_as(v)
                        vals.append(v[keep])
                        rows.append(ia[keep])
                        cols.append(ja[keep])
        for jac, ia, ja in self.strips:
            keep = jac.abs() > thresh
            vals.append(jac[keep])
            rows.append(ia[keep].to(torch.int64))
            cols.append(ja[keep].to(torch.int64))
        return torch.cat(vals), torch.cat(rows), torch.cat(cols)

    def to_csr(self, thresh=2e-16, divide_by_vol=False, out=None, slack=False):
        """device-side CSR row block of this (slab's) rows: (indptr, indices, data) torch tensors, columns ascending in a row,
        optionally divided by the row cell's volume (``Jacsurvol``, BROADCAST_npz.py:1206-1210).  Hand-written kernels
        (``to_csr_device``); ``to_csr_torch`` is the same result by torch ops (the cross-check, ~150x slower at C1).
        ``out`` = (indices, data) buffers of a previous assembly: reused when they hold the new non-zero count."""
        return self.to_csr_device(thresh, divide_by_vol, out, slack)

    def to_csr_torch(self, thresh=2e-16, divide_by_vol=False):
        """``to_csr`` by torch ops (masked selects per block plane, sort-based COO -> CSR): cross-check of the kernels"""
        from . import formats
        blk = self.blk
        v, r, c = self.to_coo(thresh)
        if divide_by_vol:
            ci = torch.div(r, 5 * blk.jm, rounding_mode="floor") - blk.ioff + blk.gh
            cj = torch.div(r % (5 * blk.jm), 5, rounding_mode="floor") + blk.gh
            v = v / blk.vol[cj, ci]
        n = 5 * blk.im_global * blk.jm
        return formats.coo_to_csr(v, r, c, 5 * blk.im * blk.jm, n, row0=5 * blk.jm * blk.ioff)

    def to_csr_device(self, thresh=2e-16, divide_by_vol=False, out=None, slack=False):
        """the same CSR row block as ``to_csr`` by hand-written kernels (csrc/csr.cu): per-row counts, warp-shuffle scan, ballot-
        compacted fill in column order, warp rank sort of the strip rows -- no sort of the whole matrix, no Python loop over the
        725 block planes.  Returns (indptr int64, indices int32, data float64) device tensors."""
        blk = self.blk
        dev = blk.device
        n = 5 * blk.im * blk.jm
        indptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
        counted = self.counts is not None and self.count_thresh == thresh
        counts = self.counts if counted else torch.empty(n + 1, dtype=torch.int32, device=dev)
        bsum = torch.empty(n // 2048 + 2, dtype=torch.int64, device=dev)
        region = np.asarray(self.region, dtype=np.int32)
        ns = len(self.strips)
        PP = ctypes.c_void_p * max(ns, 1)
        sj = PP(*[t[0].data_ptr() for t in self.strips])
        si = PP(*[t[1].data_ptr() for t in self.strips])
        sk = PP(*[t[2].data_ptr() for t in self.strips])
        slen = (ctypes.c_longlong * max(ns, 1))(*[t[0].numel() for t in self.strips])
        srect = np.asarray(self.strip_rects, dtype=np.int32).reshape(-1) if ns else np.zeros(4, dtype=np.int32)
        VP = ctypes.c_void_p
        if counted:
            blk.call("bcd_hybrid_csr_indptr_counted", _p(indptr), _p(counts), _p(bsum), region.ctypes.data_as(VP), ns, sj, si, slen,
                     ctypes.c_double(thresh), blk.gh, blk.im, blk.jm, blk._stream())
            self.counts = None     # the fill below uses the array as its cursor work space
        else:
            blk.call("bcd_hybrid_csr_indptr", _p(indptr), _p(counts), _p(bsum), _p(self.blocks), region.ctypes.data_as(VP), ns, sj, si, slen,
                     ctypes.c_double(thresh), blk.gh, blk.im, blk.jm, blk._stream())
        nnz = int(indptr[-1].item())
        if out is not None and out[0].numel() >= nnz and out[1].numel() >= nnz:
            # the pattern is value dependent, so the count moves a little from one Newton iterate to the next: the buffers of the
            # previous assembly serve as long as they are large enough (a cudaMalloc of tens of GB costs more than the fill)
            indices, data = out[0][:nnz], out[1][:nnz]
            self.csr_storage = out
        else:
            cap = nnz + (nnz // 100 + 1024 if out is not None or slack else 0)   # head room for the next iterate's count
            self.csr_storage = (torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.float64, device=dev))
            indices, data = self.csr_storage[0][:nnz], self.csr_storage[1][:nnz]
        blk.call("bcd_hybrid_csr_fill", _p(indices), _p(data), _p(counts), _p(indptr), _p(self.blocks), region.ctypes.data_as(VP), ns,
                 srect.ctypes.data_as(VP), sj, si, sk, slen, ctypes.c_double(thresh), _p(blk.vol if divide_by_vol else None), blk.gh,
                 blk.im, blk.jm, blk._stream())
        return indptr, indices, data

    def to_scipy_csr(self, thresh=2e-16):
        import scipy.sparse as sp
        v, r, c = self.to_coo(thresh)
        n = 5 * self.blk.im_global * self.blk.jm
        return sp.csr_matrix((v.cpu().numpy(), (r.cpu().numpy(), c.cpu().numpy())), shape=(n, n))


class BandedAssembly:
    """Jacobian -> zero filter -> CSR (/ vol) of a block too large for its 29 x 25 block planes to sit in HBM next to the CSR itself
    (C5 on one GPU: 97 GB of block values + 73 GB of CSR): the block is walked in ``nband`` i-bands, every band an i-slab of the
    block in GLOBAL numbering (the machinery of the multi-GPU path, sharding.slab_of / bcd_slab_begin, with the halo columns copied
    from the resident state instead of exchanged); ONE band-sized buffer of block values is reused by all bands, so the full block
    array never exists and every CSR value is written once.  The result is the list of CSR row blocks in row order -- what the
    reference's PETSc path consumes rank by rank (misc/PETSc_func.py:71-95); ``gather`` concatenates them on the device."""

    def __init__(self, case: Case, nband: int, device="cuda:0"):
        from . import sharding
        self.case, self.nband, self.device = case, nband, torch.device(device)
        self.bands = []
        for b in range(nband):
            sl, desc = sharding.slab_of(case, b, nband)
            lo, hi = sharding.slab_range(case.im, b, nband)
            self.bands.append((Block(sl, device, slab=desc if nband > 1 else None), lo, hi))
        nmax = max(b.im for b, _, _ in self.bands)
        self._buf = torch.empty(29 * 25 * case.jm * nmax, dtype=torch.float64, device=self.device)
        self._csr = [None] * nband      # (indices, data) storage of every band's row block, reused by the next assembly
        # the boundary strips are assembled ONCE on the whole block (49 latency-bound colour passes: run per band they cost nband
        # times as much) and every band's CSR conversion picks its rows out of these lists
        self.whole = Block(case, device) if nband > 1 else None

    def assemble_csr(self, w, coefdiag=None, divide_by_vol=True, thresh=2e-16, reuse=True):
        """``w``: the state of the whole block (device tensor (5, jm+2gh, im+2gh), ghosts filled).  ``coefdiag``: optional (jm, im)
        device tensor.  Returns [(indptr int64, indices int32, data float64), ...], one row block per band.  With ``reuse`` the
        index / value arrays of the previous assembly are overwritten when they are large enough (the Newton loop assembles the
        same block again and again); pass ``reuse=False`` to get arrays of your own."""
        gh, jm = self.case.gh, self.case.jm
        out = []
        strips = None
        if self.whole is not None:
            wb = self.whole
            wb.w.copy_(w)
            im = wb.im
            rects = [(1, im, 1, gh), (1, im, jm - gh + 1, jm)]
            if not self.case.periodic_i:     # the bands of an i-periodic block have no irregular columns: the cut is a slab edge
                rects += [(1, gh, gh + 1, jm - gh), (im - gh + 1, im, gh + 1, jm - gh)]
            rects = [r for r in rects if r[1] >= r[0] and r[3] >= r[2]]
            kind = "jv_relaxed_withjn" if self.case.periodic_i else "jv_relaxed"
            cdw = coefdiag
            if cdw is None:
                cdw = getattr(self, "_zero_cd", None)
                if cdw is None:
                    cdw = self._zero_cd = torch.zeros((jm, im), dtype=torch.float64, device=self.device)
            if getattr(self, "_strip_out", None) is None:
                s_ = 2 * gh + 1
                self._strip_out = [tuple(torch.zeros(25 * s_ * s_ * (r[1] - r[0] + 1) * (r[3] - r[2] + 1), dtype=dt, device=self.device)
                                         for dt in (torch.float64, torch.int32, torch.int32)) for r in rects]
            strips = jacobian_strips(wb, rects, coefdiag=cdw, kind=kind, out=self._strip_out)
        for blk, lo, hi in self.bands:
            blk.w.copy_(w[:, :, lo - 1:hi + 2 * gh])          # the band's columns + gh halo columns on each side
            blocks = self._buf[:29 * 25 * jm * blk.im].view(29, 5, 5, jm, blk.im)
            cd = coefdiag[:, lo - 1:hi].contiguous() if coefdiag is not None else None
            H = jacobian_hybrid(blk, coefdiag=cd, blocks=blocks, count_thresh=thresh, strips=strips)
            b = len(out)
            ip, idx, dat = H.to_csr(thresh=thresh, divide_by_vol=divide_by_vol, out=self._csr[b] if reuse else None, slack=reuse)
            if reuse:
                self._csr[b] = H.csr_storage
            out.append((ip, idx, dat))
        return out

    @staticmethod
    def gather(parts):
        """one CSR triple (device) from the row blocks"""
        ip = [parts[0][0]]
        for p in parts[1:]:
            ip.append(p[0][1:] + ip[-1][-1])
        return torch.cat(ip), torch.cat([p[1] for p in parts]), torch.cat([p[2] for p in parts])


def csr_transpose(indptr, indices, data, ncols, row0=0, stream=None):
    """CSR of A^T (device tensors indptr int64 [ncols+1], indices int32, data) from the CSR row block (indptr int64, indices int32,
    data float64) of A holding the rows row0 .. row0 + nrows - 1: the adjoint operator the reference's adjoint / optimal-forcing
    drivers build with PETSc's createTranspose (cylinder.py:1090-1177).  Hand-written kernels (csrc/csr.cu: column counts, scan,
    atomic fill, warp rank sort per transposed row); the result has ascending column indices, like scipy's csr_matrix(A.T)."""
    L = _lib.lib()
    dev = data.device
    nrows = indptr.numel() - 1
    nnz = int(indices.numel())
    st = ctypes.c_void_p((stream or torch.cuda.current_stream(dev)).cuda_stream)
    tptr = torch.empty(ncols + 1, dtype=torch.int64, device=dev)
    counts = torch.empty(ncols + 1, dtype=torch.int32, device=dev)
    cursor = torch.empty(ncols + 1, dtype=torch.int32, device=dev)
    bsum = torch.empty(ncols // 2048 + 2, dtype=torch.int64, device=dev)
    LL = ctypes.c_longlong
    _lib.check(L.bcd_csr_transpose_indptr(_p(tptr), _p(counts), _p(bsum), _p(indices), LL(nnz), LL(ncols), st), "bcd_csr_transpose_indptr")
    tind = torch.empty(nnz, dtype=torch.int32, device=dev)
    tdat = torch.empty(nnz, dtype=torch.float64, device=dev)
    _lib.check(L.bcd_csr_transpose_fill(_p(tind), _p(tdat), _p(cursor), _p(tptr), _p(indptr), _p(indices), _p(data), LL(nrows), LL(row0),
                                        LL(ncols), st), "bcd_csr_transpose_fill")
    return tptr, tind, tdat


def jacobian_hybrid(blk: "Block", coefdiag=None, kind=None, blocks=None, interior="faces", strip_buffers=None, count_thresh=None,
                    strips=None):
    """Jacobian of the current state: interior rows by the direct block kernels (one launch per
    structural column offset, no colouring), boundary strips (gh rows/columns along each side) by the
    reference colour loop restricted to those rows.
    The strip COO buffers (and the zero ``coefdiag`` used when none is given) belong to the block and are REUSED by its next
    assembly (every slot is rewritten by every assembly): the strip colour loop is replayed as a CUDA graph keyed on its
    pointers, fresh buffers would mean a new capture per call.  Pass ``strip_buffers="fresh"`` to get buffers of your own."""
    im, jm, gh = blk.im, blk.jm, blk.gh
    if kind is None:
        kind = "jv_relaxed_withjn" if blk.case.periodic_i else "jv_relaxed"
    relaxed = "relaxed" in kind
    cd = None
    if coefdiag is not None:
        cd = coefdiag if isinstance(coefdiag, torch.Tensor) else _t(coefdiag, blk.device)
    elif relaxed:
        cd = getattr(blk, "_zero_coefdiag", None)
        if cd is None:
            cd = blk._zero_coefdiag = torch.zeros((jm, im), dtype=torch.float64, device=blk.device)
    edges = blk.slab[2] if blk.slab else 0
    ilo = 1 if edges & 1 else gh + 1          # a slab-internal edge has no irregular rows: the block kernels run up to it
    ihi = im if edges & 2 else im - gh
    region = (ilo, ihi, gh + 1, jm - gh)
    if blocks is None:
        blocks = torch.zeros((29, 5, 5, jm, im), dtype=torch.float64, device=blk.device)
    counts = None
    if count_thresh is not None and interior == "faces":
        # the assembly kernel also counts the entries |v| > count_thresh of every regular row (what to_csr would otherwise do in a
        # pass of its own over the block values)
        counts = torch.empty(5 * im * jm + 1, dtype=torch.int32, device=blk.device)
        blk.call("bcd_jacobian_interior_counted", _p(blocks), _p(counts), ctypes.c_double(count_thresh), _p(blk.w), _p(blk.nx), _p(blk.ny),
                 _p(blk.vol), _p(blk.volf), gh, *blk._phys, im, jm, _p(cd if relaxed else None), ctypes.c_void_p(None), blk._stream())
    else:
        blk.call("bcd_jacobian_interior" if interior == "faces" else "bcd_jacobian_interior_ad", _p(blocks), _p(blk.w), _p(blk.nx), _p(blk.ny),
                 _p(blk.vol), _p(blk.volf), gh, *blk._phys, im, jm, _p(cd if relaxed else None), ctypes.c_void_p(None), blk._stream())
    rects = [(1, im, 1, gh), (1, im, jm - gh + 1, jm)]
    if not edges & 1:
        rects.append((1, gh, gh + 1, jm - gh))
    if not edges & 2:
        rects.append((im - gh + 1, im, gh + 1, jm - gh))
    rects = [r for r in rects if r[1] >= r[0] and r[3] >= r[2]]
    if strips is not None:
        # strip lists computed elsewhere in GLOBAL numbering (BandedAssembly: once on the whole block; the CSR conversion of this
        # row block takes its own rows out of them): only the regular rows are assembled here
        strips = list(strips)
        if len(rects) > len(strips):
            raise ValueError("fewer strip lists than irregular sides of this block")
        # the conversion sorts the rows of `rects` and reads one rectangle per list: empty rectangles for the lists of other bands
        rects = rects + [(1, 0, 1, 0)] * (len(strips) - len(rects))
        return HybridJacobian(blk, blocks, jacobian_slots(blk), strips, region, strip_rects=rects, counts=counts,
                              count_thresh=count_thresh)
    if strip_buffers == "fresh":
        strip_buffers = None
    elif strip_buffers is None:
        cache = blk.__dict__.setdefault("_strip_buffers", {})
        key = tuple(rects)
        if key not in cache:
            s_ = 2 * gh + 1
            cache[key] = [(torch.zeros(25 * s_ * s_ * (r[1] - r[0] + 1) * (r[3] - r[2] + 1), dtype=torch.float64, device=blk.device),
                           torch.zeros(25 * s_ * s_ * (r[1] - r[0] + 1) * (r[3] - r[2] + 1), dtype=torch.int32, device=blk.device),
                           torch.zeros(25 * s_ * s_ * (r[1] - r[0] + 1) * (r[3] - r[2] + 1), dtype=torch.int32, device=blk.device))
                          for r in rects]
        strip_buffers = cache[key]
    strips = jacobian_strips(blk, rects, coefdiag=cd, kind=kind, out=strip_buffers)
    return HybridJacobian(blk, blocks, jacobian_slots(blk), strips, region, strip_rects=rects, counts=counts, count_thresh=count_thresh)


def jacobian_strips(blk: "Block", rects, coefdiag=None, kind="jv_relaxed", out=None):
    """reference colour loop restricted to the rows of up to four rectangles, all in the same 49 passes (bcd_jacobian_strips);
    returns one compact COO triple (jac, ia, ja) per rectangle"""
    gh = blk.gh
    s = 2 * gh + 1
    if not rects:
        return []
    if out is None:
        out = []
        for r in rects:
            nb = 25 * s * s * (r[1] - r[0] + 1) * (r[3] - r[2] + 1)
            out.append((torch.zeros(nb, dtype=torch.float64, device=blk.device), torch.zeros(nb, dtype=torch.int32, device=blk.device),
                        torch.zeros(nb, dtype=torch.int32, device=blk.device)))
    n = len(rects)
    ra = np.asarray(rects, dtype=np.int32).reshape(-1)
    PP = ctypes.c_void_p * n
    jp, ip, kp = PP(*[t[0].data_ptr() for t in out]), PP(*[t[1].data_ptr() for t in out]), PP(*[t[2].data_ptr() for t in out])
    if coefdiag is None and "relaxed" in kind:
        coefdiag = torch.zeros((blk.jm, blk.im), dtype=torch.float64, device=blk.device)
    descs, nb_ = _bc_descs(blk)
    blk.call("bcd_jacobian_strips", n, ra.ctypes.data_as(ctypes.c_void_p), jp, ip, kp, _p(blk.w), _p(blk.nx), _p(blk.ny), _p(blk.vol),
             _p(blk.volf), gh, *blk._phys, blk.im, blk.jm, blk.wall, descs, nb_, SCATTER[kind], _p(coefdiag), blk._stream())
    return out


# ----------------------------------------------------------------------------------------------
# the step after the assembly: Newton correction on the device (csrc/solve.cu)
# ----------------------------------------------------------------------------------------------
def csr_spmv(indptr, indices, data, x, out=None, stream=None):
    """y = A x for a device CSR row block (indptr int64, indices int32, data float64): hand-written warp-per-row kernel"""
    L = _lib.lib()
    nrows = indptr.numel() - 1
    y = out if out is not None else torch.empty(nrows, dtype=torch.float64, device=data.device)
    st = ctypes.c_void_p((stream or torch.cuda.current_stream(data.device)).cuda_stream)
    _lib.check(L.bcd_csr_spmv(_p(y), _p(indptr), _p(indices), _p(data), _p(x), ctypes.c_longlong(nrows), st), "bcd_csr_spmv")
    return y


def block_jacobi(indptr, indices, data, col0=0, stream=None):
    """inverse 5 x 5 diagonal blocks of a CSR row block, (25, ncell) device tensor (bcd_block_jacobi_setup); raises if a block is
    singular"""
    L = _lib.lib()
    n = indptr.numel() - 1
    if n % 5:
        raise ValueError("the row block must hold whole cells (5 rows each)")
    ncell = n // 5
    dinv = torch.empty((25, ncell), dtype=torch.float64, device=data.device)
    nbad = torch.zeros(1, dtype=torch.int32, device=data.device)
    st = ctypes.c_void_p((stream or torch.cuda.current_stream(data.device)).cuda_stream)
    _lib.check(L.bcd_block_jacobi_setup(_p(dinv), _p(nbad), _p(indptr), _p(indices), _p(data), ctypes.c_longlong(ncell),
                                        ctypes.c_longlong(col0), st), "bcd_block_jacobi_setup")
    if int(nbad.item()):
        raise _lib.BroadcastB200Error(f"{int(nbad.item())} singular diagonal blocks")
    return dinv


def gmres(indptr, indices, data, b, x0=None, restart=40, maxit=2000, rtol=1e-10, precond=True, side="left", stream=None):
    """Solve A x = b on the device (whole square matrix on this GPU): restarted GMRES preconditioned by the inverse diagonal blocks
    (``precond``: True, False, or a tensor from ``block_jacobi``) on the ``side`` "left" (default: the iteration then controls
    M^-1 (b - A x), which is free of the cell-size scaling of the rows) or "right".  Returns (x, info) with info = dict(matvecs,
    converged, relres = true ||b - A x|| / ||b||, relres_iter = the controlled one).  The reference's counterpart is the
    PETSc / MUMPS LU solve of misc/PETSc_func.py:137-152, 247-263."""
    L = _lib.lib()
    L.bcd_gmres_work_doubles.restype = ctypes.c_longlong
    dev = data.device
    n = indptr.numel() - 1
    if b.numel() != n:
        raise ValueError("right-hand side and matrix sizes differ")
    dinv = None
    if isinstance(precond, torch.Tensor):
        dinv = precond
    elif precond:
        dinv = block_jacobi(indptr, indices, data, stream=stream)
    nw = int(L.bcd_gmres_work_doubles(ctypes.c_longlong(n), int(restart)))
    if nw < 0:
        raise ValueError("restart must be in 1 .. 63")
    work = torch.empty(nw, dtype=torch.float64, device=dev)
    x = torch.zeros(n, dtype=torch.float64, device=dev) if x0 is None else x0.to(dtype=torch.float64, device=dev, copy=True).contiguous()
    b = b.contiguous()
    info = (ctypes.c_int32 * 2)()
    relres = (ctypes.c_double * 2)()
    st = ctypes.c_void_p((stream or torch.cuda.current_stream(dev)).cuda_stream)
    _lib.check(L.bcd_gmres(_p(x), _p(b), _p(indptr), _p(indices), _p(data), _p(dinv), ctypes.c_longlong(n), int(restart), int(maxit),
                           ctypes.c_double(rtol), {"left": 1, "right": 0}[side], _p(work), ctypes.c_longlong(nw), info, relres, st),
               "bcd_gmres")
    return x, {"matvecs": int(info[0]), "converged": bool(info[1]), "relres": float(relres[0]), "relres_iter": float(relres[1])}


def newton_step(blk: "Block", coefdiag=None, restart=40, maxit=2000, rtol=1e-10, update=True):
    """One Newton iteration of the reference's loop (BROADCAST_npz.py:1018-1172) with every stage on the device: boundary fills,
    residual, relaxed Jacobian -> zero filter -> CSR (not divided by the volume: the matrix of iterNewton), solve A dw = res,
    w += dw.  ``coefdiag``: the pseudo-time term cflm1 * vol (:1067) as a (jm, im) device tensor or an (im, jm) numpy array, or None.
    Returns (dw as an (im, jm, 5) tensor, info)."""
    gh, im, jm = blk.gh, blk.im, blk.jm
    blk.apply_bcs()
    blk.residual()
    cd = None
    if coefdiag is not None:
        cd = coefdiag if isinstance(coefdiag, torch.Tensor) else _t(coefdiag, blk.device)
    H = jacobian_hybrid(blk, coefdiag=cd, count_thresh=2e-16)
    indptr, indices, data = H.to_csr(thresh=2e-16, divide_by_vol=False)
    # right-hand side in the row numbering of the matrix: row = e + 5 (j - 1) + 5 jm (i - 1) (misc/ComputeJacobian.f90:526) =
    # ravel(res[gh:-gh, gh:-gh, :]) of the Fortran array (BROADCAST_npz.py:1157)
    rhs = blk.res[:, gh:gh + jm, gh:gh + im].permute(2, 1, 0).contiguous().view(-1)
    dw, info = gmres(indptr, indices, data, rhs, restart=restart, maxit=maxit, rtol=rtol)
    dw3 = dw.view(im, jm, 5)
    if update and info["converged"]:
        blk.w[:, gh:gh + jm, gh:gh + im] += dw3.permute(2, 1, 0)
    info["nnz"] = int(data.numel())
    return dw3, info


def newton_loop(blk: "Block", cfl: float, nit: int, restart=40, maxit=4000, rtol=1e-8, verbose=False):
    """The reference's fixed-point (pseudo-transient Newton) loop, BROADCAST_npz.py:1007-1172, with every stage on the device:
    dt = cfl (yc(1,2) - yc(1,1)) / (1 / Mach + 1); per iteration: boundary fills, residual, norms; the relaxation follows the residual,
    cflm1 = max(norm / norm0, ninf / ninf0 over the first three equations) / dt (:1059-1061), coefdiag = cflm1 vol (:1067);
    Jacobian -> zero filter -> CSR; dw = A^-1 res (here: GMRES + block Jacobi on the resident matrix instead of MUMPS);
    w += dw -- or, when the correction is not finite or the solve did not converge, the previous correction is taken back and
    the CFL number halved (:1164-1169).  Returns the history [(iteration, norm[5], ninf[5], cflm1, matvecs), ...] (matvecs <= 0: the
    step was rejected).  This is the reference's iteration PROTOCOL; the cards run it at CFL = 1e10 (card_bl2d_fv_npz.py:56: the
    un-relaxed Jacobian, a direct solver's job), which the block-Jacobi iteration does not reach -- use moderate CFL numbers here and
    the CSR / PETSc output with the reference's LU for the final Newton iterations."""
    c = blk.case
    gh, im, jm = blk.gh, blk.im, blk.jm
    dt = cfl * float(c.yc[gh, gh + 1] - c.yc[gh, gh]) / (1.0 / float(c.phys["mach"]) + 1.0)
    dtm1 = 1.0 / dt
    vol = blk.vol[gh:gh + jm, gh:gh + im].contiguous()
    hist, norm0m1, ninf0m1, dw_old = [], None, None, None
    for it in range(1, nit + 1):
        blk.apply_bcs()
        blk.residual()
        norm, ninf = blk.norms()
        if it == 1:
            norm0m1, ninf0m1 = 1.0 / np.maximum(norm, 1e-15), 1.0 / np.maximum(ninf, 1e-15)
        r = float(max((norm[:3] * norm0m1[:3]).max(), (ninf[:3] * ninf0m1[:3]).max()))
        cflm1 = r * dtm1
        try:
            dw, info = newton_step(blk, coefdiag=cflm1 * vol, restart=restart, maxit=maxit, rtol=rtol, update=False)
            ok = info["converged"] and bool(torch.isfinite(dw).all())
        except _lib.BroadcastB200Error:      # a state that has left the physical range (singular diagonal blocks): rejected like a NaN
            dw, info, ok = None, {"matvecs": 0}, False
        if ok:
            blk.w[:, gh:gh + jm, gh:gh + im] += dw.permute(2, 1, 0)
            dw_old = dw.clone()
        else:
            if dw_old is not None:
                blk.w[:, gh:gh + jm, gh:gh + im] -= dw_old.permute(2, 1, 0)
            dtm1 *= 2.0          # cfl = cfl / 2
        hist.append((it, norm, ninf, cflm1, info["matvecs"] if ok else -info["matvecs"]))
        if verbose:
            print(f"newton {it}: |res|_2 = {norm}, 1/cfl = {cflm1:.3e}, matvecs = {info['matvecs']}, {'ok' if ok else 'rejected'}", flush=True)
    return hist
