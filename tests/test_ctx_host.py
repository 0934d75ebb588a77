"""A plain C host (tests/host/ctx_host.c, built with gcc against include/broadcast_b200.h and libbroadcast_b200.so) drives the
resident C-ABI context: CPU test = the header is valid C and every bcast_ctx_* symbol links; GPU test = the C program's residual,
norms and CSR Jacobian equal the Python paths array for array."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host", "ctx_host.c")
EXE = os.path.join(HERE, "host", "ctx_host")
LIBDIR = os.path.join(ROOT, "broadcast_b200")


def build():
    deps = [SRC, os.path.join(ROOT, "include", "broadcast_b200.h"), os.path.join(LIBDIR, "libbroadcast_b200.so")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Werror", SRC, "-o", EXE, "-L" + LIBDIR, "-lbroadcast_b200",
                               "-Wl,-rpath," + LIBDIR])
    return EXE


def test_c_host_compiles_and_links():
    exe = build()
    out = subprocess.run([exe], capture_output=True)          # no arguments: returns 1 before touching the device
    assert out.returncode == 1


def write_case(path, c, coef):
    from broadcast_b200.resident import _KIND, _interf
    p = c.phys
    recs, tables = [], []
    for bc in c.bcs:
        kind = bc[0]
        if kind == "jn":
            for prr, prd, tr in bc[1:]:
                recs.append([5, 0, 0, 0, 0, *_interf(prr), *_interf(prd), *np.asarray(tr, dtype=np.int32), 0])
                tables.append(None)
        else:
            t = np.asfortranarray(bc[3], dtype=np.float64) if kind in ("inflow", "noref") else None
            loc = [ord(ch) for ch in bc[1]] + [0]
            recs.append([_KIND[kind], *loc, *_interf(bc[2]), 0, 0, 0, 0, 0, 0, t.shape[0] if t is not None else 0])
            tables.append(t)
    with open(path, "wb") as fh:
        np.array([c.im, c.jm, c.gh, 0 if "nowall" in c.scheme else 1, len(recs)], dtype=np.int32).tofile(fh)
        np.array([p["cp"], p["cv"], p["prandtl"], p["gam"], p["rgaz"], p["cs"], p["muref"], p["tref"], p["cs"], c.k2, c.k4]).tofile(fh)
        for a in (c.nx, c.ny, c.vol, c.volf, c.w, coef):
            fh.write(np.asfortranarray(a, dtype=np.float64).tobytes(order="F"))
        for r, t in zip(recs, tables):
            np.array(r, dtype=np.int32).tofile(fh)
            np.zeros(2).tofile(fh)              # bc_desc_t.param (kinds 6 and 9 only)
            if t is not None:
                fh.write(t.tobytes(order="F"))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,im,jm", [("bl", 48, 26), ("cyl", 42, 30)])
def test_c_host_matches_the_python_paths(gpu, tmp_path, kind, im, jm):
    from broadcast_b200.cabi_ctx import Context
    exe = build()
    c = H.make_case(kind, im, jm, gpu, with_w=True)
    coef = np.asfortranarray(np.random.default_rng(3).uniform(0.5, 1.5, size=(im, jm)))
    fin, fout = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    write_case(fin, c, coef)
    out = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    raw = open(fout, "rb").read()
    nnz = int(np.frombuffer(raw, np.int64, 1)[0])
    off = 8
    n2 = np.frombuffer(raw, np.float64, 5, off); off += 40
    ninf = np.frombuffer(raw, np.float64, 5, off); off += 40
    nres = c.w.size
    res = np.frombuffer(raw, np.float64, nres, off).reshape(c.w.shape, order="F"); off += 8 * nres
    nrow = 5 * im * jm
    indptr = np.frombuffer(raw, np.int64, nrow + 1, off); off += 8 * (nrow + 1)
    indices = np.frombuffer(raw, np.int32, nnz, off); off += 4 * nnz
    data = np.frombuffer(raw, np.float64, nnz, off)
    ctx = Context(c)
    ctx.upload_state(c.w)
    assert np.array_equal(ctx.residual(), res)
    m2, minf = ctx.norms()
    assert np.allclose(n2, m2, rtol=1e-13) and np.allclose(ninf, minf, rtol=1e-13)     # atomics: summation order varies
    p2, i2, d2 = ctx.jacobian_csr(coefdiag=coef, divide_by_vol=True)
    assert nnz == len(d2) and np.array_equal(indptr, p2) and np.array_equal(indices, i2) and np.array_equal(data, d2)
    ctx.close()
