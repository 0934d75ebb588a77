"""ctypes view of the resident C-ABI context (include/broadcast_b200.h: bcast_ctx_*), the interface a C or Fortran host binds
(INTEGRATION.md section 3).  Host numpy arrays in, host numpy arrays out; device memory is owned by the library, not by torch.
Used by the tests to check the context against the Python resident layer and the oracle; a Python driver would normally use
broadcast_b200.resident.Block instead."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .cases import Case
from .resident import _BcDesc, _KIND, _interf

VP, D, LL = ctypes.c_void_p, ctypes.c_double, ctypes.c_longlong


def _a(x):
    return x.ctypes.data_as(VP)


class Context:
    def __init__(self, case: Case):
        self.lib = _lib.lib()
        self.case = case
        self.im, self.jm, self.gh = case.im, case.jm, case.gh
        p = case.phys
        self.h = VP()
        phys = [D(float(v)) for v in (p["cp"], p["cv"], p["prandtl"], p["gam"], p["rgaz"], p["cs"], p["muref"], p["tref"], p["cs"],
                                      case.k2, case.k4)]
        _lib.check(self.lib.bcast_ctx_create(ctypes.byref(self.h), self.im, self.jm, self.gh, *phys, 0 if "nowall" in case.scheme else 1),
                   "bcast_ctx_create")
        _lib.check(self.lib.bcast_ctx_set_geometry(self.h, _a(case.nx), _a(case.ny), _a(case.vol), _a(case.volf)), "bcast_ctx_set_geometry")
        descs, self._keep = [], []
        for bc in case.bcs:
            kind = bc[0]
            if kind == "jn":
                for prr, prd, tr in bc[1:]:
                    d = _BcDesc()
                    d.kind = 5
                    d.window[:] = [int(v) for v in _interf(prr)]
                    d.prd[:] = [int(v) for v in _interf(prd)]
                    d.tr[:] = [int(v) for v in np.asarray(tr)]
                    descs.append(d)
            else:
                d = _BcDesc()
                d.kind = _KIND[kind]
                d.loc = bc[1].encode()
                d.window[:] = [int(v) for v in _interf(bc[2])]
                if kind in ("inflow", "noref"):
                    t = np.asfortranarray(bc[3], dtype=np.float64)
                    self._keep.append(t)
                    d.table = t.ctypes.data
                    d.lm = int(t.shape[0])
                elif kind in ("wall_iso", "pressure"):
                    d.param[:] = [float(bc[3]), float(bc[4])]
                elif kind in ("wall_blow_profile", "wall_iso_profile"):
                    t = np.ascontiguousarray(np.asarray(bc[3], dtype=np.float64).ravel())
                    self._keep.append(t)
                    d.table = t.ctypes.data
                    d.lm = int(t.size)
                    d.param[:] = [0.0, float(bc[4]) if len(bc) > 4 else 0.0]
                descs.append(d)
        arr = (_BcDesc * max(len(descs), 1))(*descs)
        _lib.check(self.lib.bcast_ctx_set_bcs(self.h, arr, len(descs)), "bcast_ctx_set_bcs")

    def close(self):
        if self.h:
            self.lib.bcast_ctx_destroy(self.h)
            self.h = VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_state(self, w):
        w = np.asfortranarray(w, dtype=np.float64)
        _lib.check(self.lib.bcast_ctx_upload_state(self.h, _a(w)), "bcast_ctx_upload_state")

    def state(self):
        w = self.case.zeros_state()
        _lib.check(self.lib.bcast_ctx_download_state(self.h, _a(w)), "bcast_ctx_download_state")
        return w

    def residual(self):
        _lib.check(self.lib.bcast_ctx_residual(self.h), "bcast_ctx_residual")
        res = self.case.zeros_state()
        _lib.check(self.lib.bcast_ctx_download_residual(self.h, _a(res)), "bcast_ctx_download_residual")
        return res

    def norms(self):
        n2, ninf = np.zeros(5), np.zeros(5)
        _lib.check(self.lib.bcast_ctx_norms(self.h, _a(n2), _a(ninf)), "bcast_ctx_norms")
        return n2, ninf

    def jacobian_csr(self, coefdiag=None, divide_by_vol=False, thresh=2e-16, scatter_kind=-1):
        nnz = LL(0)
        cd = np.asfortranarray(coefdiag, dtype=np.float64) if coefdiag is not None else None
        _lib.check(self.lib.bcast_ctx_jacobian_csr(self.h, _a(cd) if cd is not None else VP(None), int(bool(divide_by_vol)), D(thresh),
                                                   int(scatter_kind), ctypes.byref(nnz)), "bcast_ctx_jacobian_csr")
        n = 5 * self.im * self.jm
        indptr, indices, data = np.zeros(n + 1, np.int64), np.zeros(nnz.value, np.int32), np.zeros(nnz.value, np.float64)
        _lib.check(self.lib.bcast_ctx_download_csr(self.h, _a(indptr), _a(indices), _a(data)), "bcast_ctx_download_csr")
        return indptr, indices, data
