"""ctypes binding of ``libbroadcast_b200.so`` (the C ABI of include/broadcast_b200.h).

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but
every compute call raises ``BroadcastB200Error`` if the library reports an error (e.g. no device).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbroadcast_b200.so")


class BroadcastB200Error(RuntimeError):
    pass


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BroadcastB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(broadcast_b200 has no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.bc_last_error.restype = ctypes.c_char_p
        _lib.bc_launch_count.restype = ctypes.c_longlong
    return _lib


def _conv(a):
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data)
    if isinstance(a, str):
        return ctypes.c_char_p(a.encode())
    if isinstance(a, bytes):
        return ctypes.c_char_p(a)
    if isinstance(a, (int, np.integer)):
        return ctypes.c_int(int(a))
    if isinstance(a, (float, np.floating)):
        return ctypes.c_double(float(a))
    if a is None:
        return ctypes.c_void_p(None)
    raise TypeError(f"cannot pass {type(a)} through the C ABI")


# arguments that are 64-bit integers in the header
_INT64_ARGS = {
    "computejacobianfromjv": (10,), "computejacobianfromjv_relaxed": (10,), "computejacobianfromjv_relaxed_withjn": (10,),
    "computejacobianfromjv_withjn": (10,), "computejacobianfromdz": (10,), "computejacobianfromjv_relaxed_withjnandcheck": (10,),
}


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().bc_last_error()
        raise BroadcastB200Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def call_host(name: str, *args):
    """backend for broadcast_b200.f2py_api: full Fortran argument list -> bc_<name>."""
    fn = getattr(lib(), "bc_" + name)
    fn.restype = ctypes.c_int
    keep = list(args)  # keep converted temporaries alive
    cargs = []
    wide = _INT64_ARGS.get(name, ())
    for n, a in enumerate(args):
        if n in wide:
            cargs.append(ctypes.c_int64(int(a)))
        else:
            cargs.append(_conv(a))
    rc = fn(*cargs)
    del keep
    check(rc, "bc_" + name)


def launch_count() -> int:
    return int(lib().bc_launch_count())


def device_count() -> int:
    return int(lib().bc_device_count())
