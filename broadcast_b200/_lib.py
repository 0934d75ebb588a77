"""ctypes binding of ``libbroadcast_b200.so`` (the C ABI of include/broadcast_b200.h).

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but
every compute call raises ``BroadcastB200Error`` if the library reports an error (e.g. no device).
"""
from __future__ import annotations

import ctypes
import os
import re

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
    """untyped marshalling for the resident (bcd_*) entry points, whose callers build ctypes values themselves"""
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


# ---- argument kinds of every bc_* entry point, read from include/broadcast_b200.h --------------------------------
# The reference drivers pass scalars as whatever numpy gave them: 0-d arrays (cp = dic['Cp'], BROADCAST_npz.py:433-443),
# numpy scalars, Python ints for real dummies (k4 = 1) and floats for integer ones (gh = 3.0, card_cyl2d.py:67).  f2py
# casts them to the Fortran dummy's type; so does call_host, from the declared C type of each parameter.
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "broadcast_b200.h")
_DECL = re.compile(r"^int\s+(bc_\w+)\s*\(([^;]*?)\)\s*;", re.M | re.S)
_signatures = None


def _kind_of(param: str) -> str:
    p = re.sub(r"/\*.*?\*/", " ", param, flags=re.S).strip()
    if "*" in p:
        return "str" if re.match(r"(const\s+)?char\b", p) else "ptr"
    t = p.rsplit(None, 1)[0] if " " in p else p
    t = t.replace("const", "").strip()
    if t == "double":
        return "f64"
    if t in ("int64_t", "long long"):
        return "i64"
    if t in ("int", "int32_t"):
        return "i32"
    raise BroadcastB200Error(f"{HEADER_PATH}: cannot classify parameter {param!r}")


def signatures() -> dict:
    """{entry point: tuple of 'ptr' | 'str' | 'f64' | 'i32' | 'i64'} for every `int bc_*(...)` of the header"""
    global _signatures
    if _signatures is None:
        with open(HEADER_PATH) as f:
            text = re.sub(r"//[^\n]*", "", f.read())
        sig = {}
        for name, params in _DECL.findall(text):
            params = params.strip()
            sig[name] = () if params in ("", "void") else tuple(_kind_of(q) for q in params.split(","))
        _signatures = sig
    return _signatures


def _as_int(a, what):
    if isinstance(a, np.ndarray):
        if a.size != 1:
            raise ValueError(f"{what}: expected an integer scalar, got an array of shape {a.shape}")
        a = a.reshape(()).item()
    v = int(a)
    if v != a:
        raise ValueError(f"{what}: expected an integer, got {a!r}")
    return v


def _as_float(a, what):
    if isinstance(a, np.ndarray):
        if a.size != 1:
            raise ValueError(f"{what}: expected a real scalar, got an array of shape {a.shape}")
        a = a.reshape(()).item()
    return float(a)


def marshal(name: str, args):
    """ctypes values of `args` for entry point `name`, converted to the declared parameter types (or ValueError/TypeError)"""
    kinds = signatures()[name]
    if len(kinds) != len(args):
        raise TypeError(f"{name} takes {len(kinds)} arguments, {len(args)} given")
    out = []
    for n, (k, a) in enumerate(zip(kinds, args)):
        what = f"{name} argument {n + 1}"
        if k == "ptr":
            if a is None:
                out.append(ctypes.c_void_p(None))
            elif isinstance(a, np.ndarray):
                out.append(ctypes.c_void_p(a.ctypes.data))
            else:
                raise TypeError(f"{what}: expected a numpy array, got {type(a).__name__}")
        elif k == "str":
            out.append(ctypes.c_char_p(a if isinstance(a, bytes) else str(a).encode()))
        elif k == "f64":
            out.append(ctypes.c_double(_as_float(a, what)))
        elif k == "i64":
            out.append(ctypes.c_int64(_as_int(a, what)))
        else:
            out.append(ctypes.c_int(_as_int(a, what)))
    return out


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().bc_last_error()
        raise BroadcastB200Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def call_host(name: str, *args):
    """backend for broadcast_b200.f2py_api: full Fortran argument list -> bc_<name>."""
    fn = getattr(lib(), "bc_" + name)
    fn.restype = ctypes.c_int
    keep = list(args)  # keep converted temporaries alive
    rc = fn(*marshal("bc_" + name, args))
    del keep
    check(rc, "bc_" + name)


def launch_count() -> int:
    return int(lib().bc_launch_count())


def device_count() -> int:
    return int(lib().bc_device_count())
