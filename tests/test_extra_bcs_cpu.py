"""Pins the checker for the boundary fills of SURVEY.md 8(f3) (isothermal wall, symmetry plane): oracle/_ref =
srcfv/prepro/bc_wall_viscous_iso.f90, bc_symmetry.f90 and their Tapenade tangents.  The reference ships no vectors for them, so:
the tangent routine against central differences of the primal one, the defining invariants of each fill (mirror density / reflected
velocity / wall temperature), and the committed golden outputs (tests/golden/bcs, oracle/make_golden.py --bcs) that the GPU tests
compare the product with."""
import glob
import os

import numpy as np
import pytest

import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "bcs", "*.npz")))
TWALL = 1.3
NAMES = ["iso", "sym", "anti", "pres", "presnr", "blow", "isoprof"]
GAMD, RGAZD = 0.3, -0.2      # tangents of the gas constants handed to the profile-wall routines (their shipped tangents are active in them)


def profiles(c, name, loc, interf):
    """seeded wall profile along the line (blowing velocity / wall temperature) and its tangent"""
    n = int(interf[1, 1] - interf[0, 1] + 1) if loc[0] == "I" else int(interf[1, 0] - interf[0, 0] + 1)
    rng = np.random.default_rng(17)
    x = rng.uniform(0.5, 1.5, n)
    return (1e-2 * (x - 1.0), rng.standard_normal(n) * 1e-2) if name == "blow" else (TWALL * x, rng.standard_normal(n))


def sides(c):
    """(loc, interf) of the four sides of the block, corner ownership as BROADCAST_npz.py:706-731"""
    im, jm = c.im, c.jm
    return [("Ilo", np.array([[1, 1], [1, jm]])), ("Ihi", np.array([[im, 1], [im, jm]])),
            ("Jlo", np.array([[1, 1], [im, 1]])), ("Jhi", np.array([[1, jm], [im, jm]]))]


def strip(a, loc, gh):
    """the ghost layers a fill at `loc` may write (the fixtures store these only; everything else must stay as it was)"""
    return {"Ilo": a[:gh], "Ihi": a[-gh:], "Jlo": a[:, :gh], "Jhi": a[:, -gh:]}[loc]


def check_against_golden(g, name, loc, gh, w, wd, w_in, wd_in, tol):
    assert np.all(H.rel_err(strip(w, loc, gh), g[f"{name}_{loc}_w"]) < tol), (name, loc)
    assert np.all(H.rel_err(strip(wd, loc, gh), g[f"{name}_{loc}_wd"]) < tol), (name, loc)
    for a, b in ((w, w_in), (wd, wd_in)):          # nothing outside the ghost strip of that side is touched
        a, b = a.copy(), b.copy()
        strip(a, loc, gh)[...] = 0.0
        strip(b, loc, gh)[...] = 0.0
        assert np.array_equal(a, b), (name, loc)


def fill(mods, name, c, w, loc, interf, wd=None, gas=(0.0, 0.0), prof_shift=0.0):
    """one fill (primal, or tangent when wd is given); gas / prof_shift displace gam, rgaz and the wall profile of the PRIMAL
    profile-wall routines (for the finite-difference check of their active scalar inputs)"""
    p = c.phys
    if prof_shift:
        pr, prd = profiles(c, name, loc, interf)
        pr = pr + prof_shift * prd
        if name == "blow":
            mods["f_bnd"].bc_wall_blow_profile_2d(w, pr, loc, p["gam"] + gas[0], interf, c.gh, c.im, c.jm)
        else:
            mods["f_bnd"].bc_wall_viscous_iso_profile_2d(w, pr, loc, p["gam"] + gas[0], p["rgaz"] + gas[1], interf, c.gh, c.im, c.jm)
        return
    if name == "iso":
        if wd is None:
            mods["f_bnd"].bc_wall_viscous_iso_2d(w, TWALL, loc, p["gam"], p["rgaz"], interf, c.gh, c.im, c.jm)
        else:
            mods["f_lin"].bc_wall_viscous_iso_2d_d(w, wd, TWALL, loc, p["gam"], p["rgaz"], interf, c.gh, c.im, c.jm)
    elif name == "blow":
        pr, prd = profiles(c, name, loc, interf)
        if wd is None:
            mods["f_bnd"].bc_wall_blow_profile_2d(w, pr, loc, p["gam"] + gas[0], interf, c.gh, c.im, c.jm)
        else:
            mods["f_lin"].bc_wall_blow_profile_2d_d(w, wd, pr, prd, loc, p["gam"], GAMD, interf, c.gh, c.im, c.jm)
    elif name == "isoprof":
        pr, prd = profiles(c, name, loc, interf)
        if wd is None:
            mods["f_bnd"].bc_wall_viscous_iso_profile_2d(w, pr, loc, p["gam"] + gas[0], p["rgaz"] + gas[1], interf, c.gh, c.im, c.jm)
        else:
            mods["f_lin"].bc_wall_viscous_iso_profile_2d_d(w, wd, pr, prd, loc, p["gam"], GAMD, p["rgaz"], RGAZD, interf, c.gh, c.im, c.jm)
    elif name in ("sym", "anti"):
        rt = "bc_symmetry_2d" if name == "sym" else "bc_antisymmetry_2d"
        if wd is None:
            getattr(mods["f_bnd"], rt)(w, loc, interf, c.nx, c.ny, c.gh, c.im, c.jm)
        else:
            getattr(mods["f_lin"], rt + "_d")(w, wd, loc, interf, c.nx, c.ny, c.gh, c.im, c.jm)
    else:   # pressure outlet, plain ("pres") or with the characteristic blend ("presnr")
        g = c.gh
        q = c.w[g, g]
        pext = 0.97 * (p["gam"] - 1.0) * (q[4] - 0.5 * (q[1] ** 2 + q[2] ** 2 + q[3] ** 2) / q[0])
        noref = name == "presnr"
        if wd is None:
            mods["f_bnd"].bc_pressure_2d(w, loc, interf, pext, noref, p["gam"], c.nx, c.ny, c.im, c.jm, c.gh)
        else:
            mods["f_lin"].bc_pressure_2d_d(w, wd, loc, interf, pext, noref, p["gam"], c.nx, c.ny, c.im, c.jm, c.gh)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("kind,im,jm", [("bl", 20, 12), ("cyl", 24, 12)])
def test_tangent_fill_is_the_derivative_of_the_primal_fill(ref, name, kind, im, jm):
    c = H.make_case(kind, im, jm, ref, with_w=True)
    w0, _ = H.residual_sequence(ref, c)
    rng = np.random.default_rng(2)
    d = np.asfortranarray(rng.standard_normal(w0.shape) * 1e-2 * np.abs(w0).max(axis=(0, 1)))
    for loc, interf in sides(c):
        if name == "presnr" and (kind, loc) != ("bl", "Ihi"):
            continue      # the wave-direction switches are piecewise constant: differences only where no switch is near its jump
        w, wd = w0.copy(order="F"), d.copy(order="F")
        fill(ref, name, c, w, loc, interf, wd)
        wp, wm = np.asfortranarray(w0 + 1e-5 * d), np.asfortranarray(w0 - 1e-5 * d)
        if name in ("blow", "isoprof"):      # the profile, gam and rgaz move along with w
            fill(ref, name, c, wp, loc, interf, gas=(1e-5 * GAMD, 1e-5 * RGAZD), prof_shift=1e-5)
            fill(ref, name, c, wm, loc, interf, gas=(-1e-5 * GAMD, -1e-5 * RGAZD), prof_shift=-1e-5)
        else:
            fill(ref, name, c, wp, loc, interf)
            fill(ref, name, c, wm, loc, interf)
        fd = (wp - wm) / 2e-5
        assert np.all(H.rel_err(fd, wd) < 1e-7), (name, loc, H.rel_err(fd, wd))
        w1 = w0.copy(order="F")
        fill(ref, name, c, w1, loc, interf)
        assert np.array_equal(w1, w)          # the tangent routine also writes the primal ghosts, identically


def test_fill_invariants(ref):
    c = H.make_case("bl", 20, 12, ref, with_w=True)
    w0, _ = H.residual_sequence(ref, c)
    g, p = c.gh, c.phys
    # symmetry at Jlo: ghost de mirrors row de-1; density equal, wall-normal momentum reversed (the wall is y = 0: normal = e_y)
    w = w0.copy(order="F")
    fill(ref, "sym", c, w, "Jlo", sides(c)[2][1])
    for de in range(1, g + 1):
        gi, ii = g - de, g + de - 1
        assert np.array_equal(w[g:-g, gi, 0], w[g:-g, ii, 0])
        assert np.allclose(w[g:-g, gi, 2], -w[g:-g, ii, 2], rtol=1e-13, atol=0)
        assert np.allclose(w[g:-g, gi, 1], w[g:-g, ii, 1], rtol=1e-13, atol=1e-300)
        assert np.array_equal(w[g:-g, gi, 3], w0[g:-g, gi, 3])        # rho w is not written
    # isothermal wall at Jlo: mean of the first ghost and first interior density is pw / (rgaz twall)
    w = w0.copy(order="F")
    fill(ref, "iso", c, w, "Jlo", sides(c)[2][1])
    def pres(q):
        return (p["gam"] - 1.0) * (q[..., 4] - 0.5 * (q[..., 1] ** 2 + q[..., 2] ** 2 + q[..., 3] ** 2) / q[..., 0])
    pw = 1.125 * pres(w0[g:-g, g]) - 0.125 * pres(w0[g:-g, g + 1])
    assert np.allclose(0.5 * (w[g:-g, g - 1, 0] + w0[g:-g, g, 0]), pw / (p["rgaz"] * TWALL), rtol=1e-13)
    assert np.allclose(w[g:-g, g - 1, 1] / w[g:-g, g - 1, 0], -w0[g:-g, g, 1] / w0[g:-g, g, 0], rtol=1e-13, atol=1e-300)


def test_bc_golden_present():
    assert len(GOLD) >= 2


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_ref_reproduces_the_bc_golden(ref, path):
    g = np.load(path)
    c = H.make_case(str(g["kind"]), int(g["im"]), int(g["jm"]), ref, with_w=True)
    w0, _ = H.residual_sequence(ref, c)
    for name in NAMES:
        for loc, interf in sides(c):
            d = np.asfortranarray(np.random.default_rng(int(g["seed"])).standard_normal(w0.shape))
            w, wd = w0.copy(order="F"), d.copy(order="F")
            fill(ref, name, c, w, loc, interf, wd)
            check_against_golden(g, name, loc, c.gh, w, wd, w0, d, 1e-13)


def test_profile_wall_tangents_accept_the_driver_arity(ref):
    """BROADCAST_npz_sens.py:1763 calls flinwall(w, wd, velprof, velprofd, 'Jlo', gam, interf, gh, im, jm): no gamd, although the shipped
    Tapenade routine has one.  The signature layer (shared by the product and the oracle) takes both; a missing gamd / rgazd is 0."""
    c = H.make_case("bl", 20, 12, ref, with_w=True)
    w0, _ = H.residual_sequence(ref, c)
    p = c.phys
    loc, interf = sides(c)[2]
    d = np.asfortranarray(np.random.default_rng(1).standard_normal(w0.shape))
    for name in ("blow", "isoprof"):
        pr, prd = profiles(c, name, loc, interf)
        outs = []
        for driver_style in (True, False):
            w, wd = w0.copy(order="F"), d.copy(order="F")
            if name == "blow":
                args = (p["gam"],) if driver_style else (p["gam"], 0.0)
                ref["f_lin"].bc_wall_blow_profile_2d_d(w, wd, pr, prd, loc, *args, interf, c.gh, c.im, c.jm)
            else:
                args = (p["gam"], p["rgaz"]) if driver_style else (p["gam"], 0.0, p["rgaz"], 0.0)
                ref["f_lin"].bc_wall_viscous_iso_profile_2d_d(w, wd, pr, prd, loc, *args, interf, c.gh, c.im, c.jm)
            outs.append((w, wd))
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        assert not np.array_equal(outs[0][1], d)


@pytest.mark.parametrize("loc", ["Ilo", "Ihi", "Jlo", "Jhi"])
def test_bc_general_is_the_dirichlet_fill_it_names(ref, loc):
    """bc_general_2d (srcfv/borders/bc_general.F90): ghost layer de of line cell l = field(l, de, :), nothing else touched; its tangent
    zeroes the same ghosts and leaves w alone -- the defining property pins the oracle's binding (interface array in Fortran order)"""
    import helpers as H
    c = H.make_case("bl", 19, 13, ref, with_w=True)
    gh, im, jm = c.gh, c.im, c.jm
    rng = np.random.default_rng(1)
    lm = jm if loc[0] == "I" else im
    if loc[0] == "I":
        i = 1 if loc == "Ilo" else im
        itf = np.array([[i, 1], [i, jm]], dtype=float)
    else:
        j = 1 if loc == "Jlo" else jm
        itf = np.array([[1, j], [im, j]], dtype=float)
    field = np.asfortranarray(rng.standard_normal((lm, gh, 5)))
    w = c.w.copy(order="F")
    ref["f_bnd"].bc_general_2d(w, loc, itf, field, gh, im, jm)
    exp = c.w.copy(order="F")
    ghost = np.zeros(exp.shape[:2], dtype=bool)
    for de in range(1, gh + 1):
        sl = {"Ilo": (gh - de, slice(gh, gh + jm)), "Ihi": (gh + im - 1 + de, slice(gh, gh + jm)),
              "Jlo": (slice(gh, gh + im), gh - de), "Jhi": (slice(gh, gh + im), gh + jm - 1 + de)}[loc]
        exp[sl] = field[:, de - 1, :]
        ghost[sl] = True
    assert np.array_equal(w, exp)
    wd = np.asfortranarray(rng.standard_normal(w.shape))
    wd0, w0 = wd.copy(order="F"), w.copy(order="F")
    ref["f_lin"].bc_general_2d_d(w, wd, loc, itf, field, gh, im, jm)
    assert np.array_equal(w, w0) and not wd[ghost].any() and np.array_equal(wd[~ghost], wd0[~ghost])
