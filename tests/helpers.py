"""Driver call sequences of the reference (BROADCAST_npz.py:1011-1137, cylinder.py:841-985) written
once against the f2py-shaped module surface, so the same sequence runs on the product and on the
oracle."""
import numpy as np

from broadcast_b200 import cases


def rel_err(a, b):
    """max |a-b| per equation plane, relative to max |b| of that plane (0/0 -> 0)."""
    a, b = np.asarray(a), np.asarray(b)
    ax = tuple(range(a.ndim - 1)) if a.ndim > 1 else None
    scale = np.abs(b).max(axis=ax)
    diff = np.abs(a - b).max(axis=ax)
    return np.where(scale > 0, diff / np.where(scale > 0, scale, 1.0), diff)


def make_case(kind, im, jm, mods, **kw):
    if kind == "bl":
        return cases.make_bl_case(im, jm, f_geom=mods["f_geom"], **kw)
    return cases.make_cyl_case(im, jm, f_geom=mods["f_geom"], f_bnd=mods["f_bnd"], **kw)


def residual_sequence(mods, case, scheme="flux_num_dnc5_2d"):
    w = case.w.copy(order="F")
    cases.apply_bcs(case, w, mods["f_bnd"])
    res = case.zeros_state()
    getattr(mods["f_sch"], scheme)(res, w, *case.scheme_args())
    return w, res


def tangent_sequence(mods, case, w, wd, scheme="flux_num_dnc5_2d_d"):
    w = w.copy(order="F")
    wd = wd.copy(order="F")
    cases.apply_bcs_lin(case, w, wd, mods["f_bnd"], mods["f_lin"])
    res, resd = case.zeros_state(), case.zeros_state()
    getattr(mods["f_lin"], scheme)(res, resd, w, wd, *case.scheme_args())
    return wd, resd


def jacobian_sequence(mods, case, w, colours=None, coefdiag=None, scheme_d="flux_num_dnc5_2d_d"):
    """Colour loop of BROADCAST_npz.py:1068-1127 (or cylinder.py:941-978 for periodic-in-i cases).
    Returns the COO lists restricted to the visited colours (full lists if colours is None)."""
    im, jm, gh = case.im, case.jm, case.gh
    s = 2 * gh + 1
    n = 5 * im * jm
    nb = 25 * s * s * im * jm
    jac = np.zeros(nb)
    ia = np.zeros(nb, dtype=np.int32)
    ja = np.zeros(nb, dtype=np.int32)
    if coefdiag is None:
        coefdiag = np.zeros((im, jm), order="F")
    wd = case.zeros_state()
    res, resd = case.zeros_state(), case.zeros_state()
    if colours is None:
        colours = [(m, l, k) for m in range(5) for l in range(s) for k in range(s)]
    segs = []
    f_misc, f_lin = mods["f_misc"], mods["f_lin"]
    for (m, l, k) in colours:
        wd *= 0.0
        f_misc.testvector(wd, m, l, k, gh, im, jm)
        ww = w.copy(order="F")
        cases.apply_bcs_lin(case, ww, wd, mods["f_bnd"], f_lin)
        getattr(f_lin, scheme_d)(res, resd, ww, wd, *case.scheme_args())
        if case.periodic_i:
            f_misc.computejacobianfromjv_relaxed_withjn(jac, ia, ja, resd, m, l, k, gh, coefdiag)
        else:
            f_misc.computejacobianfromjv_relaxed(jac, ia, ja, resd, m, l, k, gh, coefdiag)
        base = k * n + l * n * s + m * n * s * s
        segs.append(slice(base, base + n))
    sel = np.concatenate([np.arange(sg.start, sg.stop) for sg in segs])
    return jac[sel], ia[sel], ja[sel]


def coo_to_dict(jac, ia, ja, thresh=2e-16):
    """remove_zero_jac (BROADCAST_npz.py:129-135) then duplicate summation (scipy csr semantics)."""
    import scipy.sparse as sp
    keep = np.abs(jac) > thresh
    n = int(max(ia.max(), ja.max())) + 1
    return sp.csr_matrix((jac[keep], (ia[keep], ja[keep])), shape=(n, n))


# ----------------------------------------------------------------------------------------------------------------------
# Error metrics for the residual (and anything else that is a DIFFERENCE OF FACE FLUXES).
#
# north_star: "residuals and Jacobian values within 1e-12 relative".  A residual is -(hn(i+1) - hn(i)) - (hn(j+1) - hn(j)):
# at a converged or smooth state it is orders of magnitude smaller than the fluxes it is made of, so any two correct
# evaluations (the reference built with and without FMA contraction, say) differ by a few ulp OF THE FLUXES, which is an
# arbitrarily large multiple of an ulp of the residual.  "Relative" is therefore read as a BACKWARD error: the difference
# divided by the magnitude of the face fluxes of the cell.  The forward, plane-maximum metric of round 1 (rel_err above) is
# still reported next to it, with the reference's own FMA / no-FMA spread as the noise floor of each metric.
# ----------------------------------------------------------------------------------------------------------------------
def flux_scale(case, w):
    """(im, jm, 5): per interior cell and equation, S * lam * W_e with S the summed length of the cell's four faces, lam = |V| + c
    the spectral radius of the cell state and W = (rho, rho lam, rho lam, rho lam, rho E + p) the characteristic size of the
    conservative variable (momenta by rho lam, so that a velocity component that happens to vanish does not make the scale vanish).
    Every term of a face flux -- convective (V.n) w_e, pressure p n <= rho lam^2 |n|, scalar dissipation rspec eps d(w_e), viscous --
    is bounded by |n| lam W_e: this is the magnitude of the fluxes whose difference the residual is."""
    gh, gam = case.gh, float(case.phys["gam"])
    q = np.asarray(w)[gh:-gh, gh:-gh]
    ro = q[..., 0]
    v2 = (q[..., 1] ** 2 + q[..., 2] ** 2 + q[..., 3] ** 2) / ro ** 2
    p = (gam - 1.0) * (q[..., 4] - 0.5 * ro * v2)
    lam = np.sqrt(v2) + np.sqrt(gam * p / ro)
    n = np.hypot(case.nx, case.ny)
    sl = slice(gh, -gh - 1)
    S = n[gh:-gh - 1, sl, 0] + n[gh + 1:-gh or None, sl, 0] + n[sl, gh:-gh - 1, 1] + n[sl, gh + 1:-gh or None, 1]
    W = np.stack([ro, ro * lam, ro * lam, ro * lam, q[..., 4] + p], axis=-1)
    return (S * lam)[..., None] * W


def backward_err(res, res_ref, case, w):
    """max over interior cells of |res - res_ref| / flux_scale, per equation"""
    gh = case.gh
    s = flux_scale(case, w)
    d = np.abs(np.asarray(res)[gh:-gh, gh:-gh] - np.asarray(res_ref)[gh:-gh, gh:-gh])
    return np.where(s > 0, d / np.where(s > 0, s, 1.0), d).max(axis=(0, 1))


def residual_errors(res, res_ref, case, w):
    """{'plane_max': forward error relative to the plane maximum (round-1 metric), 'backward': error / face-flux magnitude}"""
    gh = case.gh
    return {"plane_max": rel_err(np.asarray(res)[gh:-gh, gh:-gh], np.asarray(res_ref)[gh:-gh, gh:-gh]),
            "backward": backward_err(res, res_ref, case, w)}


BACKWARD_TOL = 1e-13   # ten times tighter than north_star's 1e-12, read as a backward error (measured: a few 1e-16)


def assert_residual_parity(res, res_ref, case, w, floor=None, what=""):
    """The parity assertion of every residual-like comparison: backward error below BACKWARD_TOL and, when the reference's own
    FMA / no-FMA spread `floor` (dict of residual_errors between the two oracle builds on the same input) is given, forward error
    within 1e-12 or four times that floor, whichever is larger (the excess a re-association may add to the reference's own noise)."""
    e = residual_errors(res, res_ref, case, w)
    assert np.all(np.isfinite(np.asarray(res))), what
    assert np.all(e["backward"] < BACKWARD_TOL), (what, e)
    if floor is not None:
        lim = np.maximum(1e-12, 4.0 * floor["plane_max"])
        assert np.all(e["plane_max"] <= lim), (what, e, floor)
    return e


_fast_mods = None


def fma_floor(case, scheme="flux_num_dnc5_2d"):
    """residual_errors between the oracle's two builds (-O2 -ffp-contract=off vs -O3 -march=x86-64-v3 with FMA) on the case's state:
    the noise floor of each metric on this grid"""
    global _fast_mods
    from oracle import refmods
    if _fast_mods is None:
        _fast_mods = (refmods.make(), refmods.make(fast=True))
    a, b = _fast_mods
    w, r0 = residual_sequence(a, case, scheme)
    _, r1 = residual_sequence(b, case, scheme)
    return residual_errors(r1, r0, case, w)
