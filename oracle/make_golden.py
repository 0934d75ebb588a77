#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from oracle/_ref  --  TEST INFRASTRUCTURE ONLY.

oracle/_ref is the reference's own Fortran (read in place under /root/reference) machine-translated
to C by oracle/f90_to_c.py and compiled with gcc -O2 -ffp-contract=off.  The reference ships no golden
vectors, so these fixtures are "outputs of the reference itself run here":
they travel to the GPU box, where /root/reference does not exist, and pin both the numpy oracle and
the CUDA path.

    python oracle/make_golden.py          # rewrites tests/golden/*.npz (needs oracle/_ref built)

Each fixture stores the complete inputs of every call (so a test needs nothing else) and the outputs:
geometry, boundary fill, residual (wall and nowall variants), tangent of a seeded random direction
through the linearised boundary fills, a sample of colours through seed -> BC_d -> tangent -> scatter
(COO ia/ja/jac), norms, and Dz/Dz2 operator rows for one colour.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refmods  # noqa: E402
from broadcast_b200 import cases  # noqa: E402
import helpers as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
COLOURS = [(0, 0, 0), (1, 3, 2), (4, 6, 6), (2, 5, 0), (3, 0, 4)]
FIXTURES = [("bl", 24, 16), ("cyl", 28, 16)]


def make(kind, im, jm, R):
    c = H.make_case(kind, im, jm, R, with_w=True)
    gh = c.gh
    out = dict(kind=kind, im=im, jm=jm, gh=gh, colours=np.array(COLOURS, dtype=np.int32))
    for n in ("x0", "y0", "nx", "ny", "xc", "yc", "vol", "volf"):
        out["geom_" + n] = getattr(c, n)
    out["w_init"] = c.w
    w, res = H.residual_sequence(R, c)
    out["w_filled"], out["res"] = w, res
    _, out["res_nowall"] = H.residual_sequence(R, c, "flux_num_dnc5_nowall_2d")
    rng = np.random.default_rng(11)
    wd0 = np.asfortranarray(rng.standard_normal(w.shape))
    out["wd_in"] = wd0
    out["wd_filled"], out["resd"] = H.tangent_sequence(R, c, w, wd0)
    coef = np.asfortranarray(np.random.default_rng(12).uniform(0.5, 1.5, size=(im, jm)))
    out["coefdiag"] = coef
    out["coo_jac"], out["coo_ia"], out["coo_ja"] = H.jacobian_sequence(R, c, w, COLOURS, coef)
    n2, ninf = R["f_norm"].compute_norml2inf(res, im, jm, gh)
    out["norm_l2"], out["norm_inf"] = n2, ninf
    if "f_dz" in R:
        m, l, k = COLOURS[1]
        wd = c.zeros_state()
        R["f_misc"].testvector(wd, m, l, k, gh, im, jm)
        dz, dz2 = c.zeros_state(), c.zeros_state()
        a = c.scheme_args()
        dzargs = a[:18] + a[20:]  # coeffs_5p_dz takes no k2, k4
        R["f_dz"].coeffs_5p_dz(dz, w, wd, *dzargs)
        R["f_dz"].coeffs_5p_dz2(dz2, w, wd, *dzargs)
        out["dz_colour"], out["dz"], out["dz2"] = np.array([m, l, k], dtype=np.int32), dz, dz2
    return out


def make_dz_tangent(kind, im, jm, R):
    """outputs of the reference's srcfv/tangentdz/coeffs_5p_dz_d.f90 / coeffs_5p_dz2_d.f90 (f_lindz, BROADCAST_npz_sens.py:1768-1769)
    on the filled state of the fixture, a seeded base-flow variation wd0 and a seeded mode wd (both stored)"""
    c = H.make_case(kind, im, jm, R, with_w=True)
    w, _ = H.residual_sequence(R, c)
    rng = np.random.default_rng(21)
    wd = np.asfortranarray(rng.standard_normal(w.shape))
    wd0 = np.asfortranarray(rng.standard_normal(w.shape) * np.abs(w).max(axis=(0, 1)))
    a = c.scheme_args()
    dzargs = a[:18] + a[20:]
    out = dict(kind=kind, im=im, jm=jm, gh=c.gh, wd=wd, wd0=wd0)
    for name, key in (("coeffs_5p_dz_d", "dzd"), ("coeffs_5p_dz2_d", "dz2d")):
        z, zd = c.zeros_state(), c.zeros_state()
        getattr(R["f_lindz"], name)(z, zd, w, wd0, wd, *dzargs)
        out[key] = zd
    return out


def make_bcs(kind, im, jm, R):
    """outputs of the reference's isothermal-wall and symmetry fills and their tangents (srcfv/prepro/bc_wall_viscous_iso.f90,
    bc_symmetry.f90, srcfv/tangent/bc_wall_viscous_iso_d.f90, bc_symmetry_d.f90) on each side of the fixture's filled state"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_extra_bcs_cpu as T
    c = H.make_case(kind, im, jm, R, with_w=True)
    w0, _ = H.residual_sequence(R, c)
    rng = np.random.default_rng(31)
    d = np.asfortranarray(rng.standard_normal(w0.shape))
    out = dict(kind=kind, im=im, jm=jm, gh=c.gh, seed=31, twall=T.TWALL)   # wd_in = default_rng(seed).standard_normal(w.shape)
    for name in T.NAMES:
        for loc, interf in T.sides(c):
            w, wd = w0.copy(order="F"), d.copy(order="F")
            T.fill(R, name, c, w, loc, interf, wd)
            out[f"{name}_{loc}_w"], out[f"{name}_{loc}_wd"] = T.strip(w, loc, c.gh).copy(), T.strip(wd, loc, c.gh).copy()
    return out


def main():
    if not refmods.available():
        sys.path.insert(0, HERE)
        import build_ref
        if not build_ref.build(verbose=True):
            raise SystemExit("oracle/_ref is not built and /root/reference is absent")
    R = refmods.make()
    os.makedirs(OUT, exist_ok=True)
    only_dz_tangent = "--dz-tangent" in sys.argv   # add the f_lindz fixtures without rewriting the others
    only_bcs = "--bcs" in sys.argv                 # same for the isothermal-wall / symmetry fixtures
    for kind, im, jm in FIXTURES:
        d = make_bcs(kind, im, jm, R)
        p = os.path.join(OUT, "bcs", f"{kind}_{im}x{jm}.npz")
        os.makedirs(os.path.dirname(p), exist_ok=True)
        np.savez_compressed(p, **d)
        print(p, os.path.getsize(p) // 1024, "KiB")
        if only_bcs:
            continue
        d = make_dz_tangent(kind, im, jm, R)
        p = os.path.join(OUT, "lindz", f"{kind}_{im}x{jm}.npz")
        os.makedirs(os.path.dirname(p), exist_ok=True)
        np.savez_compressed(p, **d)
        print(p, os.path.getsize(p) // 1024, "KiB")
        if only_dz_tangent:
            continue
        d = make(kind, im, jm, R)
        p = os.path.join(OUT, f"{kind}_{im}x{jm}.npz")
        np.savez_compressed(p, **d)
        print(p, os.path.getsize(p) // 1024, "KiB")


if __name__ == "__main__":
    main()
