"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/broadcast_b200.h declares; calls fail loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "broadcast_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bcd?_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from broadcast_b200 import _lib
    lib = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import broadcast_b200 as bb
    wd = np.zeros((16, 16, 5), order="F")
    with pytest.raises(bb.BroadcastB200Error):
        bb.f_misc.testvector(wd, 0, 0, 0, 3, 10, 10)


def test_signature_layer_validates_like_f2py():
    import broadcast_b200 as bb
    w = np.zeros((16, 16, 5))  # C-ordered: f2py rejects it for intent(inout)
    with pytest.raises(ValueError):
        bb.f_bnd.bc_wall_viscous_adia_2d(w, "Jlo", 1.4, np.array([[1, 1], [10, 1]]), 3, 10, 10)
    w = np.zeros((16, 16, 5), order="F")
    with pytest.raises(ValueError):
        bb.f_bnd.bc_wall_viscous_adia_2d(w, "Klo", 1.4, np.array([[1, 1], [10, 1]]), 3, 10, 10)
