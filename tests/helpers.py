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


def jacobian_sequence(mods, case, w, colours=None, coefdiag=None):
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
        f_lin.flux_num_dnc5_2d_d(res, resd, ww, wd, *case.scheme_args())
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
