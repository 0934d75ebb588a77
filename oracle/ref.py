"""ctypes loader for ``oracle/_ref/libbroadcast_ref.so``  --  TEST INFRASTRUCTURE ONLY.

The library is the reference's own Fortran (onera/Broadcast ``srcfv/prepro``, ``srcfv/tangent``,
``srcfv/dz``, ``misc/ComputeJacobian.f90`` ...) machine-translated to C by ``oracle/f90_to_c.py``
and compiled by ``oracle/build_ref.py``.  Calls take the FULL Fortran argument list, in Fortran
order (f2py's optional trailing dimension arguments are all explicit here).

Arrays must be Fortran-ordered; real arrays float64 (modified in place when the Fortran dummy
is ``intent(inout)``), integer arrays are converted to int32 copies when needed (``interf`` is
passed as a float array by the reference drivers and cast by f2py: BROADCAST_npz.py:706-731).
"""
from __future__ import annotations

import ctypes
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")


def available(fast: bool = False) -> bool:
    name = "libbroadcast_ref_fast.so" if fast else "libbroadcast_ref.so"
    return os.path.exists(os.path.join(REFDIR, name)) and os.path.exists(os.path.join(REFDIR, "manifest.json"))


class RefLib:
    def __init__(self, fast: bool = False):
        name = "libbroadcast_ref_fast.so" if fast else "libbroadcast_ref.so"
        self.path = os.path.join(REFDIR, name)
        self.lib = ctypes.CDLL(self.path)
        with open(os.path.join(REFDIR, "manifest.json")) as fh:
            self.manifest = {m["name"]: m for m in json.load(fh)}
        for name, m in self.manifest.items():
            setattr(self, name, self._make(name, m))

    def _make(self, name, m):
        cfun = getattr(self.lib, name)
        cfun.restype = None
        argtypes = []
        for a in m["args"]:
            if a["dims"] is not None:
                argtypes.append(ctypes.c_void_p)
            elif a["type"] == "char":
                argtypes.append(ctypes.c_char_p)
            elif a["type"] == "int":
                argtypes.append(ctypes.c_int)
            else:
                argtypes.append(ctypes.c_double)
        cfun.argtypes = argtypes
        spec = m["args"]

        def call(*args):
            if len(args) != len(spec):
                raise TypeError(f"{name}: expected {len(spec)} arguments ({[s['name'] for s in spec]}), got {len(args)}")
            cargs = []
            keep = []
            for a, s in zip(args, spec):
                if s["dims"] is not None:
                    want = np.float64 if s["type"] == "real" else np.int32
                    arr = a
                    if not isinstance(arr, np.ndarray) or arr.dtype != want or not arr.flags.f_contiguous:
                        if s["intent"] in ("inout", "out") and s["type"] == "real":
                            raise TypeError(f"{name}: argument {s['name']} must be a Fortran-ordered float64 ndarray")
                        arr = np.asfortranarray(np.asarray(a), dtype=want)
                        if s["intent"] in ("inout", "out"):
                            # integer inout arrays (ia, ja): must already be int32 F-contiguous
                            raise TypeError(f"{name}: argument {s['name']} must be an int32 ndarray")
                    keep.append(arr)
                    cargs.append(arr.ctypes.data)
                elif s["type"] == "char":
                    cargs.append(a.encode() if isinstance(a, str) else a)
                elif s["type"] == "int":
                    cargs.append(int(a))
                else:
                    cargs.append(float(a))
            cfun(*cargs)

        call.__name__ = name
        call.__doc__ = f"{name}({', '.join(s['name'] for s in spec)})  [{m['src']}:{m['first_line']}-{m['last_line']}]"
        return call


_cache = {}


def load(fast: bool = False) -> RefLib:
    if fast not in _cache:
        _cache[fast] = RefLib(fast)
    return _cache[fast]
