"""CPU tests of the checker itself (no GPU): oracle/_ref (the reference Fortran machine-translated to C and
run here) against the committed golden fixtures, the numpy restatement against the same fixtures, and the
derivative / assembly identities that pin the tangent and the colouring."""
import glob
import os

import numpy as np
import pytest

import helpers as H
from broadcast_b200 import cases

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
TOL = 1e-12


def _load(p):
    d = np.load(p)
    return {k: d[k] for k in d.files}


def _case_from_golden(g, mods):
    return H.make_case(str(g["kind"]), int(g["im"]), int(g["jm"]), mods, with_w=True)


def test_golden_fixtures_present():
    assert len(GOLD) >= 2


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_ref_reproduces_golden(ref, path):
    """oracle/_ref rebuilt on this machine reproduces the committed outputs (libm differences only)."""
    g = _load(path)
    c = _case_from_golden(g, ref)
    for n in ("x0", "y0", "nx", "ny", "xc", "yc", "vol", "volf"):
        assert np.array_equal(getattr(c, n), g["geom_" + n]), n
    assert np.array_equal(c.w, g["w_init"])
    w, res = H.residual_sequence(ref, c)
    assert np.all(H.rel_err(w, g["w_filled"]) < 1e-14)
    assert np.all(H.rel_err(res, g["res"]) < 1e-13)
    _, rnw = H.residual_sequence(ref, c, "flux_num_dnc5_nowall_2d")
    assert np.all(H.rel_err(rnw, g["res_nowall"]) < 1e-13)
    wd, resd = H.tangent_sequence(ref, c, w, np.asfortranarray(g["wd_in"]))
    assert np.all(H.rel_err(wd, g["wd_filled"]) < 1e-13)
    assert np.all(H.rel_err(resd, g["resd"]) < 1e-13)
    jac, ia, ja = H.jacobian_sequence(ref, c, w, [tuple(x) for x in g["colours"]], np.asfortranarray(g["coefdiag"]))
    assert np.array_equal(ia, g["coo_ia"]) and np.array_equal(ja, g["coo_ja"])
    assert np.abs(jac - g["coo_jac"]).max() < 1e-13 * np.abs(g["coo_jac"]).max()
    # spanwise operator rows of one colour (srcfv/dz/coeffs_5p_dz.F90, coeffs_5p_dz2.F90)
    m, l, k = (int(x) for x in g["dz_colour"])
    wd = c.zeros_state()
    ref["f_misc"].testvector(wd, m, l, k, c.gh, c.im, c.jm)
    a = c.scheme_args()
    dzargs = a[:18] + a[20:]
    for name, key in (("coeffs_5p_dz", "dz"), ("coeffs_5p_dz2", "dz2")):
        z = c.zeros_state()
        getattr(ref["f_dz"], name)(z, w, wd, *dzargs)
        assert np.abs(g[key]).max() > 0
        assert np.all(H.rel_err(z, g[key]) < 1e-13), name


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_numpy_restatement_matches_golden(path):
    """the hand-written numpy restatement (oracle/broadcast_oracle.py) against the reference's outputs"""
    from oracle import broadcast_oracle as O
    g = _load(path)
    im, jm, gh = int(g["im"]), int(g["jm"]), int(g["gh"])
    # geometry from the raw node coordinates
    geo = {n: np.asfortranarray(g["geom_" + n]).copy(order="F") for n in ("x0", "y0", "nx", "ny", "xc", "yc", "vol", "volf")}
    if str(g["kind"]) == "bl":  # (the cylinder fixture has periodic metric copies applied after computegeom)
        raw = {n: np.zeros_like(v) for n, v in geo.items()}
        raw["x0"][gh:gh + im + 1, :] = geo["x0"][gh:gh + im + 1, :]
        raw["y0"][:, gh:gh + jm + 1] = geo["y0"][:, gh:gh + jm + 1]
        O.computegeom_2d(raw["x0"], raw["y0"], raw["nx"], raw["ny"], raw["xc"], raw["yc"], raw["vol"], raw["volf"], im, jm, gh)
        for n in geo:
            assert np.array_equal(raw[n], geo[n]), n
    phys = cases.nondim_physics(4.5, 288.0, 3.4e6, lref_unit_reynolds=True) if str(g["kind"]) == "bl" else \
        cases.nondim_physics(0.3, 288.0, 46.8, lref_unit_reynolds=False)
    k2, k4 = (1.01, 1.0) if str(g["kind"]) == "bl" else (0.0, 1.0)
    args = (geo["x0"], geo["y0"], geo["nx"], geo["ny"], geo["xc"], geo["yc"], geo["vol"], geo["volf"], gh, phys["cp"], phys["cv"],
            phys["prandtl"], phys["gam"], phys["rgaz"], phys["cs"], phys["muref"], phys["tref"], phys["cs"], k2, k4, im, jm)
    w = np.asfortranarray(g["w_filled"])
    res = np.zeros_like(w)
    O.flux_num_dnc5_2d(res, w, *args)
    assert np.all(H.rel_err(res, g["res"]) < TOL), H.rel_err(res, g["res"])
    res = np.zeros_like(w)
    O.flux_num_dnc5_nowall_2d(res, w, *args)
    assert np.all(H.rel_err(res, g["res_nowall"]) < TOL)
    resd = np.zeros_like(w)
    O.flux_num_dnc5_2d_d(res, resd, w, np.asfortranarray(g["wd_filled"]), *args)
    assert np.all(H.rel_err(resd, g["resd"]) < TOL), H.rel_err(resd, g["resd"])


@pytest.mark.parametrize("kind,im,jm", [("bl", 24, 16), ("cyl", 28, 16)])
def test_tangent_is_the_derivative_of_the_residual(ref, kind, im, jm):
    """central finite difference of (boundary fill + residual) vs (linearised fill + tangent): the check the
    reference authors left commented at BROADCAST_npz.py:1091-1125"""
    c = H.make_case(kind, im, jm, ref, with_w=True)
    w0 = c.w.copy(order="F")
    rng = np.random.default_rng(4)
    v = np.asfortranarray(rng.standard_normal(w0.shape))
    gh = c.gh
    v[:gh] = 0; v[-gh:] = 0; v[:, :gh] = 0; v[:, -gh:] = 0   # interior direction; ghosts follow through the BCs

    def R(w):
        c.w = w
        return H.residual_sequence(ref, c)[1]
    eps = 1e-6
    fd = (R(np.asfortranarray(w0 + eps * v)) - R(np.asfortranarray(w0 - eps * v))) / (2 * eps)
    c.w = w0
    wf, _ = H.residual_sequence(ref, c)
    _, resd = H.tangent_sequence(ref, c, wf, v)
    err = H.rel_err(resd, fd)
    assert np.all(err < 2e-6), err


@pytest.mark.parametrize("kind,im,jm", [("bl", 24, 16), ("cyl", 28, 16)])
def test_assembled_jacobian_times_vector(ref, kind, im, jm):
    """245-colour loop -> COO -> CSR, then A v == -tangent(v) + coefdiag v for a random interior v: pins the
    seeds, the scatter's (vali, valj) rules and the slot layout (misc/ComputeJacobian.f90:357-374, 503-570, 847-926)"""
    import scipy.sparse as sp
    c = H.make_case(kind, im, jm, ref, with_w=True)
    wf, _ = H.residual_sequence(ref, c)
    coef = np.asfortranarray(np.random.default_rng(6).uniform(0.5, 1.5, size=(im, jm)))
    jac, ia, ja = H.jacobian_sequence(ref, c, wf, None, coef)
    n = 5 * im * jm
    A = sp.csr_matrix((jac, (ia, ja)), shape=(n, n))
    gh = c.gh
    rng = np.random.default_rng(7)
    v = c.zeros_state()
    v[gh:-gh, gh:-gh, :] = rng.standard_normal((im, jm, 5))
    _, resd = H.tangent_sequence(ref, c, wf, v)
    # row/column numbering: e-1 + 5 (j-1) + 5 jm (i-1)
    flat = lambda a: np.ascontiguousarray(a[gh:-gh, gh:-gh, :]).reshape(-1)   # (i, j, e) C-order == that numbering
    lhs = A @ flat(v)
    rhs = -flat(resd) + np.repeat(coef.reshape(-1), 5) * flat(v)
    assert np.abs(lhs - rhs).max() < 1e-11 * np.abs(rhs).max()


def test_uniform_flow_is_preserved_by_the_nowall_scheme(ref):
    im, jm, gh = 20, 14, 3
    a = cases._alloc(im, jm, gh)
    x = np.linspace(0.0, 2.0, im + 1)
    y = np.linspace(0.0, 1.0, jm + 1)
    a["x0"][gh:gh + im + 1, :] = x[:, None]
    a["y0"][:, gh:gh + jm + 1] = y[None, :]
    ref["f_geom"].computegeom_2d(a["x0"], a["y0"], a["nx"], a["ny"], a["xc"], a["yc"], a["vol"], a["volf"], im, jm, gh)
    p = cases.nondim_physics(0.3, 288.0, 1000.0, lref_unit_reynolds=False)
    a["w"][:, :, :] = np.array([1.0, 0.7, 0.2, 0.1, p["einf"]])[None, None, :]
    res = np.zeros_like(a["w"])
    ref["f_sch"].flux_num_dnc5_nowall_2d(res, a["w"], a["x0"], a["y0"], a["nx"], a["ny"], a["xc"], a["yc"], a["vol"], a["volf"], gh,
                                         p["cp"], p["cv"], p["prandtl"], p["gam"], p["rgaz"], p["cs"], p["muref"], p["tref"], p["cs"],
                                         1.01, 1.0, im, jm)
    assert np.abs(res).max() < 1e-13


def test_seed_and_scatter_integer_semantics(ref):
    """brute force: for every interior row cell the column cell chosen by the scatter is a seed cell of the
    colour, and every seed cell within the stencil half-width gh of the row (interior rule) is that cell"""
    im, jm, gh = 23, 17, 3
    s = 2 * gh + 1
    nb = 25 * s * s * im * jm
    resd = np.asfortranarray(np.ones((im + 2 * gh, jm + 2 * gh, 5)))
    for (m, l, k) in [(0, 0, 0), (3, 2, 5), (4, 6, 6), (1, 4, 3)]:
        jac, ia, ja = np.zeros(nb), np.zeros(nb, np.int32), np.zeros(nb, np.int32)
        ref["f_misc"].computejacobianfromjv(jac, ia, ja, resd, m, l, k, gh, im, jm)
        wd = np.asfortranarray(np.zeros((im + 2 * gh, jm + 2 * gh, 5)))
        ref["f_misc"].testvector(wd, m, l, k, gh, im, jm)
        seeds = {(i, j) for i in range(1, im + 1) for j in range(1, jm + 1) if wd[i + gh - 1, j + gh - 1, m] == 1.0}
        assert seeds == {(i, j) for i in range(l + 1, im + 1, s) for j in range(k + 1, jm + 1, s)}
        n = 5 * im * jm
        base = k * n + l * n * s + m * n * s * s
        for e in range(1, 6):
            for j in range(1, jm + 1):
                for i in range(1, im + 1):
                    slot = base + (i - 1) + (j - 1) * im + (e - 1) * im * jm
                    near = [(a, b) for (a, b) in seeds if abs(a - i) <= gh and abs(b - j) <= gh]
                    if jac[slot] != 0.0:
                        assert ia[slot] == e - 1 + 5 * (j - 1) + 5 * jm * (i - 1)
                        col = ja[slot]
                        assert col % 5 == m
                        cj, ci = (col // 5) % jm + 1, col // (5 * jm) + 1
                        assert (ci, cj) in seeds
                        if near:
                            assert (ci, cj) == near[0] and len(near) == 1
                    else:
                        assert not near and ia[slot] == 5 * im * jm - 1 and ja[slot] == 0


def test_fill_skipping_rule_of_the_colour_loops(ref):
    """The device colour loops skip the linearised boundary fill of a side for the colours that have no seed row / column
    among the first three interior lines of that side (csrc/jacobian.cu, active_bcs).  Checked on the reference itself: for
    every colour the ghost tangents the reference's linearised fills write on such a side are exactly zero."""
    import helpers as H
    from broadcast_b200 import cases
    for kind, im, jm in (("bl", 23, 17), ("cyl", 28, 16)):
        c = H.make_case(kind, im, jm, ref, with_w=True)
        w, _ = H.residual_sequence(ref, c)
        gh = c.gh
        s = 2 * gh + 1
        near_lo = lambda q: q <= gh - 1
        near_hi = lambda q, n: n >= q + 1 and (n - (q + 1)) % s <= gh - 1
        skipped = 0
        for l in range(s):
            for k in range(s):
                for m in (0, 4):
                    wd = c.zeros_state()
                    ref["f_misc"].testvector(wd, m, l, k, gh, im, jm)
                    ww = w.copy(order="F")
                    cases.apply_bcs_lin(c, ww, wd, ref["f_bnd"], ref["f_lin"])
                    ghosts = {"Ilo": wd[:gh], "Ihi": wd[-gh:], "Jlo": wd[:, :gh], "Jhi": wd[:, -gh:]}
                    active = {"Ilo": near_lo(l), "Ihi": near_hi(l, im), "Jlo": near_lo(k), "Jhi": near_hi(k, jm)}
                    for side, a in active.items():
                        if kind == "cyl" and side in ("Ilo", "Ihi"):
                            continue      # periodic cut: the join copies interior tangents, always applied
                        if not a:
                            # the whole ghost strip of the side, corners included (a corner is filled by the LATER of its two
                            # sides from the ghosts of the earlier one, which are zero as well)
                            other = ("Jlo", "Jhi") if side[0] == "I" else ("Ilo", "Ihi")
                            g_ = ghosts[side]
                            if kind == "cyl" and side[0] == "J":
                                g_ = g_[gh:-gh]       # the cut's ghost columns are copies of interior cells of the far side
                            assert not np.any(g_[gh:-gh] if side[0] == "J" else g_[:, gh:-gh]), (kind, side, l, k, m)
                            if not any(active[o] for o in other):
                                assert not np.any(g_), (kind, side, l, k, m, "corners")
                            skipped += 1
        assert skipped > 0
