"""The product's block-Jacobian math (broadcast_b200/csrc/facejac.cuh is host+device code) compiled for the HOST
and checked against the reference's 245-colour loop run on oracle/_ref: every interior 5x5 block, to 1e-12 of the
largest Jacobian entry.  This is the CPU-side proof of the semi-analytic face linearisation; on the GPU the same
templates run inside k_face_packages / k_jac_assemble (tests/test_parity_gpu.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "facejac_host.cpp")
SO = os.path.join(HERE, "host", "libfacejac_host.so")
OFFSETS = [(-3, 0), (-2, -2), (-2, -1), (-2, 0), (-2, 1), (-2, 2), (-1, -2), (-1, -1), (-1, 0), (-1, 1), (-1, 2), (0, -3), (0, -2),
           (0, -1), (0, 0), (0, 1), (0, 2), (0, 3), (1, -2), (1, -1), (1, 0), (1, 1), (1, 2), (2, -2), (2, -1), (2, 0), (2, 1), (2, 2),
           (3, 0)]


@pytest.fixture(scope="module")
def hostlib():
    deps = [SRC] + [os.path.join(HERE, "..", "broadcast_b200", "csrc", f) for f in ("facejac.cuh", "scheme.cuh", "grid.cuh", "dual.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    return ctypes.CDLL(SO)


@pytest.mark.parametrize("entry", ["fj_host_blocks", "fj_host_blocks_table"])
@pytest.mark.parametrize("kind,im,jm", [("bl", 24, 16), ("cyl", 28, 16)])
def test_face_linearisation_blocks_match_reference_colour_loop(ref, hostlib, kind, im, jm, entry):
    c = H.make_case(kind, im, jm, ref, with_w=True)
    w, _ = H.residual_sequence(ref, c)
    jac, ia, ja = H.jacobian_sequence(ref, c, w, None, None)
    n = 5 * im * jm
    A = sp.csr_matrix((jac, (ia, ja)), shape=(n, n)).toarray()
    vals = np.zeros((29, 25, jm, im))
    p, gh = c.phys, c.gh
    D = ctypes.c_double
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = getattr(hostlib, entry)(P(vals), P(w), P(c.nx), P(c.ny), P(c.vol), P(c.volf), gh, D(p["cp"]), D(p["cv"]), D(p["prandtl"]),
                                D(p["gam"]), D(p["rgaz"]), D(p["cs"]), D(p["muref"]), D(p["tref"]), D(p["cs"]), D(c.k2), D(c.k4), im, jm)
    assert rc == 0
    scale = np.abs(A).max()
    worst, seen = 0.0, np.zeros((n, n), dtype=bool)
    for i in range(gh + 1, im - gh + 1):
        for j in range(gh + 1, jm - gh + 1):
            r0 = 5 * (j - 1) + 5 * jm * (i - 1)
            for s, (di, dj) in enumerate(OFFSETS):
                c0 = 5 * (j + dj - 1) + 5 * jm * (i + di - 1)
                worst = max(worst, np.abs(vals[s, :, j - 1, i - 1].reshape(5, 5) - A[r0:r0 + 5, c0:c0 + 5]).max())
                seen[r0:r0 + 5, c0:c0 + 5] = True
            # nothing of an interior row lies outside the 29 structural blocks
            assert not np.any(A[r0:r0 + 5][~seen[r0:r0 + 5]])
    assert worst < 1e-12 * scale, worst / scale
