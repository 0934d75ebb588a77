"""Synthetic inputs for the configurations named in BASELINE.json (host-side, numpy).

This is a python-3 restatement of the *set-up* part of the reference drivers: mesh recipes,
nondimensionalisation, interface (BC window) definitions and the boundary tables.  It is input
generation, not part of the timed hot path.  Geometry metrics and the periodic copies are
computed by whatever ``f_geom`` / ``f_bnd`` modules the caller injects (the CUDA drop-in in
the product, the oracle in CPU-only tests), so this file never touches ``oracle/``.

Reference anchors
-----------------
* boundary layer  : BROADCAST_npz.py:474-800 (``bl2d_prepro``), meshBL.py:57-76, card_bl2d_fv_npz.py:39-110
* cylinder O-mesh : cylinder.py:220-560 (``cyl2d``), meshCyl.py:61-114, card_cyl2d.py:41-128
* initial profile : initialisation.f90:2-58 (``initblfv``, polynomial Blasius), set_bnd.f90:2-24
* BC call order   : BROADCAST_npz.py:1018-1021 / 1079-1082, handleBC.py:32-127,129-242
"""
from __future__ import annotations

from dataclasses import dataclass, field as dc_field
from typing import Any, Dict, List, Tuple

import numpy as np


# --------------------------------------------------------------------------------------
# physics / nondimensionalisation   (BROADCAST_npz.py:121-127, 567-667; cylinder.py:310-385)
# --------------------------------------------------------------------------------------

def sutherland(propref, Ts, Cs, T):
    return propref * np.sqrt(T / Ts) * ((1.0 + Cs / Ts) / (1.0 + Cs / T))


def nondim_physics(mach, T0, runit, *, lref_unit_reynolds: bool, gam=1.4, Ts=273.15, Cs=110.4,
                   musuth=1.716e-5, rgaz=287.1, prandtl=0.72) -> Dict[str, float]:
    """StateRef == 'm0_runit_t0'.  ``lref_unit_reynolds`` selects Lref = mu/(rho U) (boundary layer,
    BROADCAST_npz.py:640-641) instead of Lref = 1 (cylinder.py:381-382)."""
    muinf = sutherland(musuth, Ts, Cs, T0)
    sound = np.sqrt(gam * rgaz * T0)
    uinf = mach * sound
    einf = rgaz / (gam - 1.0) * T0 + 0.5 * uinf * uinf
    rhoinf = runit * muinf / uinf
    pinf = rhoinf * rgaz * T0
    cp = gam * rgaz / (gam - 1.0)
    cv = rgaz / (gam - 1.0)
    Roref, Vref, Tref = rhoinf, uinf, T0
    if lref_unit_reynolds:
        Muref = muinf
        Lref = Muref / (Roref * Vref)
    else:
        Lref = 1.0
        Muref = Roref * Vref * Lref
    Cvref = Vref ** 2 / Tref
    out = dict(
        gam=gam, prandtl=prandtl, mach=mach, Lref=Lref,
        cp=cp / Cvref, cv=cv / Cvref, rgaz=rgaz / Cvref,
        tref=Ts / Tref, muref=musuth / Muref, cs=Cs / Tref, muinf=muinf / Muref,
        rhoinf=1.0, uinf=1.0, einf=einf / Vref ** 2, pinf=pinf / (Roref * Vref ** 2),
    )
    return out


# --------------------------------------------------------------------------------------
# meshes
# --------------------------------------------------------------------------------------

def bigeom_stretch_in(N_in, delta, percent):
    """meshBL.py:65-76"""
    x0 = np.linspace(0, N_in - 1, N_in)
    b = 1.0 + percent
    y_min = delta * (b - 1.0) / (b ** N_in - 1.0)
    mu = -y_min / (b - 1.0)
    xi = b / (b - 1.0) * y_min
    return np.concatenate((np.zeros(1), xi * b ** x0 + mu))


def exp_stretch_out(N_out, delta, percent, Nend):
    """meshBL.py:57-62"""
    x_0 = np.linspace(1.0, N_out, N_out)
    percent_max = np.log(Nend) / N_out
    return delta * np.exp((percent + (percent_max - percent) * x_0 / N_out) * x_0)


def stretch_tanh(x0, a, b, c):
    """meshBL.py:6-21"""
    x1 = (x0 - x0[0]) / (x0[-1] - x0[0])
    x2 = a / 2.0 * np.tanh(b * (x1 - c))
    x3 = (x2 - x2[0]) / (x2[-1] - x2[0])
    return x3 * (x0[-1] - x0[0]) + x0[0]


def cylinder2d_bisfull(r, c1, c2, imax, jmax):
    """meshCyl.py:61-114 ("MESH 5 to 8", join upstream); ``im/2`` is python-2 integer division."""
    t = np.flipud(np.linspace(0.0, 180.0, imax // 2 + 1))
    t1 = stretch_tanh(t, 1.0, 1.5, -1.0)
    t = np.linspace(180.0, 360.0, imax // 2 + 1)
    t2 = np.flipud(stretch_tanh(t, 1.0, 1.5, -1.0))
    t = np.concatenate((t1, t2[1:]))
    ang = t[:imax, None] * np.pi / 180.0
    x = c1 + r[None, :jmax] * np.cos(ang)
    y = c2 + r[None, :jmax] * np.sin(ang)
    return np.asfortranarray(x), np.asfortranarray(y)


# --------------------------------------------------------------------------------------
# case container
# --------------------------------------------------------------------------------------

@dataclass
class Case:
    name: str
    im: int
    jm: int
    gh: int
    phys: Dict[str, float]
    k2: float
    k4: float
    x0: np.ndarray
    y0: np.ndarray
    nx: np.ndarray
    ny: np.ndarray
    xc: np.ndarray
    yc: np.ndarray
    vol: np.ndarray
    volf: np.ndarray
    w: np.ndarray
    # ordered list of boundary operations, in the order the reference drivers call them:
    #   ('inflow', loc, interf, field) | ('noref', loc, interf, wbd) | ('outflow', loc, interf)
    #   ('wall', loc, interf) | ('jn', prr, prd, tr)
    bcs: List[Tuple[Any, ...]] = dc_field(default_factory=list)
    periodic_i: bool = False
    scheme: str = "flux_num_dnc5_2d"
    # i-slab of an i-periodic block (sharding.slab_of): the join across the cut has become a halo exchange between the first
    # and the last slab, which (like the join, cylinder.py:499-527: all rows, ghost rows included) must FOLLOW the fills
    slab_periodic: bool = False

    def scheme_args(self):
        """Trailing arguments of f_sch.flux_num_dnc5_2d after (res, w): BROADCAST_npz.py:1031.
        Note ``s_suth = cs`` (the drivers pass ``cs`` twice)."""
        p = self.phys
        return (self.x0, self.y0, self.nx, self.ny, self.xc, self.yc, self.vol, self.volf, self.gh,
                p["cp"], p["cv"], p["prandtl"], p["gam"], p["rgaz"], p["cs"], p["muref"], p["tref"],
                p["cs"], self.k2, self.k4, self.im, self.jm)

    def zeros_state(self):
        return np.zeros((self.im + 2 * self.gh, self.jm + 2 * self.gh, 5), order="F")


def _alloc(im, jm, gh):
    z = lambda *s: np.zeros(s, order="F")
    return dict(
        x0=z(im + 2 * gh + 1, jm + 2 * gh + 1), y0=z(im + 2 * gh + 1, jm + 2 * gh + 1),
        nx=z(im + 2 * gh + 1, jm + 2 * gh + 1, 2), ny=z(im + 2 * gh + 1, jm + 2 * gh + 1, 2),
        xc=z(im + 2 * gh, jm + 2 * gh), yc=z(im + 2 * gh, jm + 2 * gh),
        vol=z(im + 2 * gh, jm + 2 * gh), volf=z(im + 2 * gh, jm + 2 * gh, 2),
        w=z(im + 2 * gh, jm + 2 * gh, 5),
    )


def _interf(imin, jmin, imax, jmax):
    a = np.zeros((2, 2), order="F")
    a[0, 0], a[0, 1], a[1, 0], a[1, 1] = imin, jmin, imax, jmax
    return a


def _perturb(w, xc, yc, gh, amp, seed, with_w):
    """Deterministic smooth multiplicative perturbation: moves the state off
    the sensor ties (|.|, max) so branch choices are robust to 1e-16 differences."""
    if amp == 0.0:
        return
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0.0, 2.0 * np.pi, size=(5, 4))
    ni, nj = w.shape[0], w.shape[1]
    s = np.linspace(0.0, 1.0, ni)[:, None]
    t = np.linspace(0.0, 1.0, nj)[None, :]
    for e in range(5):
        mod = (np.sin(2 * np.pi * 3 * s + ph[e, 0]) * np.cos(2 * np.pi * 2 * t + ph[e, 1])
               + 0.5 * np.sin(2 * np.pi * 7 * s + 2 * np.pi * 5 * t + ph[e, 2])
               + 0.25 * np.cos(2 * np.pi * 11 * t + ph[e, 3]))
        if e == 3:
            if with_w:
                w[:, :, 3] = amp * w[:, :, 0] * mod
        elif e == 2:
            w[:, :, 2] = w[:, :, 2] + amp * w[:, :, 0] * mod
        else:
            w[:, :, e] = w[:, :, e] * (1.0 + amp * mod)


# --------------------------------------------------------------------------------------
# C1 / C4 / C5 : flat-plate boundary layer
# --------------------------------------------------------------------------------------

def make_bl_case(im=500, jm=150, *, f_geom, order=5, length=0.59, high=0.035, xini=0.006,
                 k2=1.01, k4=1.0, amp=1e-3, seed=0, with_w=False, name=None) -> Case:
    """2-D Blasius boundary layer on the ``meshBL`` stretched grid (card_bl2d_fv_npz.py, C1).

    For grids other than 500x150 the domain is scaled with (im/500, jm/150) and the wall-normal
    growth rate with 150/jm so that cell sizes and the total stretching stay those of C1
    (SURVEY.md section 8(d), C5)."""
    gh = (order + 1) // 2
    sx, sy = im / 500.0, jm / 150.0
    L, H = length * sx, high * sy
    phys = nondim_physics(4.5, 288.0, 3.4e6, lref_unit_reynolds=True)
    a = _alloc(im, jm, gh)

    x = np.linspace(xini, xini + L, im + 1)
    Ny_in = 80 * jm // 100
    deltaBL = H / 4.0
    percent = 0.02 / sy
    Ny_out = jm - Ny_in
    Nend = H / deltaBL
    y = np.concatenate((bigeom_stretch_in(Ny_in, deltaBL, percent), exp_stretch_out(Ny_out, deltaBL, percent, Nend)))
    a["x0"][gh:gh + im + 1, :] = x[:, None]
    a["y0"][:, gh:gh + jm + 1] = y[None, :]
    a["x0"] *= 1.0 / phys["Lref"]
    a["y0"] *= 1.0 / phys["Lref"]
    f_geom.computegeom_2d(a["x0"], a["y0"], a["nx"], a["ny"], a["xc"], a["yc"], a["vol"], a["volf"], im, jm, gh)

    # interfaces, BROADCAST_npz.py:706-731 (corner ownership: see docs/source/tutorialbl.rst:190-192)
    interf1 = _interf(1, 1, 1, jm)                 # Ilo  inlet
    interf2 = _interf(im, 1, im, jm + gh)          # Ihi  outflow
    interf3 = _interf(1 - gh, 1, im + gh, 1)       # Jlo  wall
    interf4 = _interf(1 - gh, jm, im, jm)          # Jhi  non-reflecting

    # state: free stream, then polynomial Blasius (initialisation.f90:20-40), then perturbation
    w = a["w"]
    state = np.array([phys["rhoinf"], phys["rhoinf"] * phys["uinf"], 0.0, 0.0, phys["rhoinf"] * phys["einf"]])
    w[:, :, :] = state[None, None, :]
    ro, rou, roE = state[0], state[1], state[4]
    ue = rou / ro
    rou2 = 0.5 * rou * ue
    xdeltam1 = np.sqrt(rou / phys["muinf"])
    xcp = np.maximum(a["xc"], 1e-300)
    eta = a["yc"] * xdeltam1 / np.sqrt(xcp)
    ubl = np.where(eta < 1.0, 2.0 * eta - 2.0 * eta ** 3 + eta ** 4, 1.0) * ue
    J = slice(gh, None)  # j = 1 .. jm+gh
    w[:, J, 1] = ro * ubl[:, J]
    w[:, J, 4] = roE - rou2 + 0.5 * ro * ubl[:, J] ** 2
    w[:, gh - 1, 1] = -w[:, gh, 1]
    _perturb(w, a["xc"], a["yc"], gh, amp, seed, with_w)

    # set_bnd.f90:13-23
    field = np.zeros((jm, gh, 5), order="F")
    wbd = np.zeros((im + gh, 5), order="F")
    for depth in range(1, gh + 1):
        field[:, depth - 1, :] = w[gh - depth, gh:gh + jm, :]
    wbd[:, :] = w[0:im + gh, gh + jm - 1, :]

    bcs = [("inflow", "Ilo", interf1, field), ("noref", "Jhi", interf4, wbd),
           ("outflow", "Ihi", interf2), ("wall", "Jlo", interf3)]
    return Case(name or f"bl2d_{im}x{jm}", im, jm, gh, phys, k2, k4, bcs=bcs, **a)


# --------------------------------------------------------------------------------------
# C2 / C3 : cylinder O-mesh, periodic in i
# --------------------------------------------------------------------------------------

def make_cyl_case(im=630, jm=300, *, f_geom, f_bnd, order=5, ri=0.5, rf=100.0, k2=0.0, k4=1.0,
                  amp=1e-3, seed=0, with_w=False, name=None) -> Case:
    """2-D cylinder at Re_D ~ 47, M = 0.3 (card_cyl2d.py, cylinder.py:271-283, 426-527)."""
    gh = (order + 1) // 2
    phys = nondim_physics(0.3, 288.0, 46.8, lref_unit_reynolds=False)
    a = _alloc(im, jm, gh)

    n_out = max(2, min(15, jm // 4))                # reference: jm-15 cells in the inner block
    Ny_in = jm - n_out
    deltaBL = 50.0 - ri
    percent = 0.016 * (285.0 / Ny_in)               # 0.016 at the reference size (jm = 300)
    high = rf - ri
    Nend = high / deltaBL
    r = np.concatenate((bigeom_stretch_in(Ny_in, deltaBL, percent), exp_stretch_out(jm - Ny_in, deltaBL, percent, Nend))) + ri
    x, y = cylinder2d_bisfull(r, 0.0, 0.0, im + 1, jm + 1)
    a["x0"][gh:gh + im + 1, gh:gh + jm + 1] = x
    a["y0"][gh:gh + im + 1, gh:gh + jm + 1] = y
    f_geom.computegeom_2d(a["x0"], a["y0"], a["nx"], a["ny"], a["xc"], a["yc"], a["vol"], a["volf"], im, jm, gh)

    jmin, jmax = 1 - gh, jm + gh
    I32 = lambda v: np.array(v, dtype=np.int32, order="F")
    prr1, prd1 = I32([[im + 1, jmin], [im + gh, jmax]]), I32([[im - gh + 1, jmin], [im, jmax]])
    prr2, prd2 = I32([[1 - gh, jmin], [0, jmax]]), I32([[1, jmin], [gh, jmax]])
    tr = I32([1, 2])
    # periodic copies of the metrics across the cut (cylinder.py:499-522)
    for arr in (a["xc"], a["yc"], a["vol"]):
        f_bnd.jn_match_geom_2d(arr, prr1, gh, gh, gh, gh, im, jm, arr, prd2, gh, gh, gh, gh, im, jm, tr)
        f_bnd.jn_match_geom_2d(arr, prr2, gh, gh, gh, gh, im, jm, arr, prd1, gh, gh, gh, gh, im, jm, tr)
    prr1g, prd1g = I32([[im + 2, jmin], [im + 1 + gh, jmax + 1]]), I32([[im + 1 - gh, jmin], [im, jmax + 1]])
    prr2g, prd2g = I32([[1 - gh, jmin], [0, jmax + 1]]), I32([[2, jmin], [1 + gh, jmax + 1]])
    for arr in (a["nx"], a["ny"]):
        f_bnd.jn_match_2d(arr, prr1g, gh, gh, gh, gh, im + 1, jm + 1, arr, prd2g, gh, gh, gh, gh, im + 1, jm + 1, tr)
        f_bnd.jn_match_2d(arr, prr2g, gh, gh, gh, gh, im + 1, jm + 1, arr, prd1g, gh, gh, gh, gh, im + 1, jm + 1, tr)
    vol, volf = a["vol"], a["volf"]
    I = slice(gh, im + 1 + gh)
    J = slice(gh, jm + 1 + gh)
    volf[I, J, 0] = 2.0 / (vol[I, J] + vol[gh - 1:im + gh, J])
    volf[I, J, 1] = 2.0 / (vol[I, J] + vol[I, gh - 1:jm + gh])
    x0, y0 = a["x0"], a["y0"]
    x0[:gh, :] = x0[im + 1:im + 1 + gh, :]
    y0[:gh, :] = y0[im + 1:im + 1 + gh, :]
    x0[im + 1 + gh:, :] = x0[gh:2 * gh, :]
    y0[im + 1 + gh:, :] = y0[gh:2 * gh, :]

    w = a["w"]
    state = np.array([phys["rhoinf"], phys["rhoinf"] * phys["uinf"], 0.0, 0.0, phys["rhoinf"] * phys["einf"]])
    w[:, :, :] = state[None, None, :]
    wbd = np.zeros((im, 5), order="F")
    wbd[:, :] = state[None, :]
    # potential-flow-like slow-down near the body so that the wall rows are not a trivial uniform state
    rr = np.sqrt(a["xc"] ** 2 + a["yc"] ** 2)
    damp = 1.0 - np.exp(-np.maximum(rr - ri, 0.0) / 2.0)
    w[:, :, 1] *= damp
    w[:, :, 4] = state[4] - 0.5 * state[1] ** 2 / state[0] + 0.5 * w[:, :, 1] ** 2 / w[:, :, 0]
    _perturb(w, a["xc"], a["yc"], gh, amp, seed, with_w)
    # make the perturbed interior periodic-consistent is the job of the BC fill (jn), nothing to do here

    interf1 = _interf(1, 1, im, 1)                 # Jlo wall
    interf2 = _interf(1, jm, im, jm)               # Jhi non-reflecting
    # order after handleBC.sortBC: noref, wall, then the join twice (Ilo and Ihi interfaces both copy)
    jn = ("jn", (prr1, prd2, tr), (prr2, prd1, tr))
    bcs = [("noref", "Jhi", interf2, wbd), ("wall", "Jlo", interf1), jn, jn]
    return Case(name or f"cyl2d_{im}x{jm}", im, jm, gh, phys, k2, k4, bcs=bcs, periodic_i=True, **a)


# --------------------------------------------------------------------------------------
# boundary fill in driver order
# --------------------------------------------------------------------------------------

def apply_bcs(case: Case, w, f_bnd):
    """Primal ghost fill, in the call order of BROADCAST_npz.py:1018-1021 / handleBC.applyBC(mode 0)."""
    im, jm, gh = case.im, case.jm, case.gh
    gam = case.phys["gam"]
    for bc in case.bcs:
        kind = bc[0]
        if kind == "inflow":
            f_bnd.bc_supandsubinlet_2d(w, bc[1], bc[2], bc[3], case.nx, case.ny, gam, im, jm)
        elif kind == "noref":
            f_bnd.bc_no_reflexion_2d(w, bc[3], bc[1], bc[2], case.nx, case.ny, gam, gh, im, jm)
        elif kind == "outflow":
            f_bnd.bc_extrapolate_o2_2d(w, bc[1], bc[2], im, jm, gh)
        elif kind == "wall":
            f_bnd.bc_wall_viscous_adia_2d(w, bc[1], gam, bc[2], gh, im, jm)
        elif kind in _EXTRA_BCS:
            _extra_bc(case, kind, bc, w, None, f_bnd, None)
        elif kind == "jn":
            for prr, prd, tr in bc[1:]:
                f_bnd.jn_match_2d(w, prr, gh, gh, gh, gh, im, jm, w, prd, gh, gh, gh, gh, im, jm, tr)
        else:
            raise ValueError(kind)


# boundary fills of SURVEY.md 8(f3): ("wall_iso", loc, interf, twall, rgaz), ("symmetry" | "antisymmetry", loc, interf),
# ("pressure", loc, interf, pext, noref)
# ("wall_blow_profile", loc, interf, velprof), ("wall_iso_profile", loc, interf, twallprof, rgaz): profile passive in the tangent
_EXTRA_BCS = ("wall_iso", "symmetry", "antisymmetry", "pressure", "wall_blow_profile", "wall_iso_profile")


def _extra_bc(case, kind, bc, w, wd, f_bnd, f_lin):
    im, jm, gh, gam = case.im, case.jm, case.gh, case.phys["gam"]
    if kind == "wall_iso":
        if wd is None:
            f_bnd.bc_wall_viscous_iso_2d(w, bc[3], bc[1], gam, bc[4], bc[2], gh, im, jm)
        else:
            f_lin.bc_wall_viscous_iso_2d_d(w, wd, bc[3], bc[1], gam, bc[4], bc[2], gh, im, jm)
    elif kind == "wall_blow_profile":
        if wd is None:
            f_bnd.bc_wall_blow_profile_2d(w, bc[3], bc[1], gam, bc[2], gh, im, jm)
        else:
            f_lin.bc_wall_blow_profile_2d_d(w, wd, bc[3], np.zeros_like(np.asarray(bc[3], dtype=float)), bc[1], gam, 0.0, bc[2], gh, im, jm)
    elif kind == "wall_iso_profile":
        if wd is None:
            f_bnd.bc_wall_viscous_iso_profile_2d(w, bc[3], bc[1], gam, bc[4], bc[2], gh, im, jm)
        else:
            f_lin.bc_wall_viscous_iso_profile_2d_d(w, wd, bc[3], np.zeros_like(np.asarray(bc[3], dtype=float)), bc[1], gam, 0.0, bc[4], 0.0,
                                                   bc[2], gh, im, jm)
    elif kind in ("symmetry", "antisymmetry"):
        if wd is None:
            getattr(f_bnd, "bc_%s_2d" % kind)(w, bc[1], bc[2], case.nx, case.ny, gh, im, jm)
        else:
            getattr(f_lin, "bc_%s_2d_d" % kind)(w, wd, bc[1], bc[2], case.nx, case.ny, gh, im, jm)
    else:
        if wd is None:
            f_bnd.bc_pressure_2d(w, bc[1], bc[2], bc[3], bool(bc[4]), gam, case.nx, case.ny, im, jm, gh)
        else:
            f_lin.bc_pressure_2d_d(w, wd, bc[1], bc[2], bc[3], bool(bc[4]), gam, case.nx, case.ny, im, jm, gh)


def apply_bcs_lin(case: Case, w, wd, f_bnd, f_lin, *, zero_frame: bool = None):
    """Linearised ghost fill on (w, wd).

    Boundary-layer driver (BROADCAST_npz.py:1079-1082): only the ``_d`` routines are called.
    Cylinder driver (handleBC.applyBC mode 1, handleBC.py:149-240): the ghost frame of ``w`` is
    zeroed, then for every interface the ``_d`` routine is followed by the primal one; joins copy
    ``wd`` with the *primal* jn_match_2d (jn_match_2d_d zeroes its output: tangent/jn_match_2d_d.f90:71)."""
    im, jm, gh = case.im, case.jm, case.gh
    gam = case.phys["gam"]
    handle_bc_style = case.periodic_i if zero_frame is None else zero_frame
    if handle_bc_style:
        w[:gh, :, :] = 0.0
        w[:, :gh, :] = 0.0
        w[-gh:, :, :] = 0.0
        w[:, -gh:, :] = 0.0
    for bc in case.bcs:
        kind = bc[0]
        if kind == "inflow":
            f_lin.bc_supandsubinlet_2d_d(w, wd, bc[1], bc[2], bc[3], case.nx, case.ny, gam, im, jm)
            if handle_bc_style:
                f_bnd.bc_supandsubinlet_2d(w, bc[1], bc[2], bc[3], case.nx, case.ny, gam, im, jm)
        elif kind == "noref":
            f_lin.bc_no_reflexion_2d_d(w, wd, bc[3], bc[1], bc[2], case.nx, case.ny, gam, gh, im, jm)
            if handle_bc_style:
                f_bnd.bc_no_reflexion_2d(w, bc[3], bc[1], bc[2], case.nx, case.ny, gam, gh, im, jm)
        elif kind == "outflow":
            f_lin.bc_extrapolate_o2_2d_d(w, wd, bc[1], bc[2], im, jm, gh)
            if handle_bc_style:
                f_bnd.bc_extrapolate_o2_2d(w, bc[1], bc[2], im, jm, gh)
        elif kind == "wall":
            f_lin.bc_wall_viscous_adia_2d_d(w, wd, bc[1], gam, bc[2], gh, im, jm)
            if handle_bc_style:
                f_bnd.bc_wall_viscous_adia_2d(w, bc[1], gam, bc[2], gh, im, jm)
        elif kind in _EXTRA_BCS:
            _extra_bc(case, kind, bc, w, wd, f_bnd, f_lin)
            if handle_bc_style:
                _extra_bc(case, kind, bc, w, None, f_bnd, None)
        elif kind == "jn":
            for prr, prd, tr in bc[1:]:
                f_bnd.jn_match_2d(wd, prr, gh, gh, gh, gh, im, jm, wd, prd, gh, gh, gh, gh, im, jm, tr)
            for prr, prd, tr in bc[1:]:
                f_bnd.jn_match_2d(w, prr, gh, gh, gh, gh, im, jm, w, prd, gh, gh, gh, gh, im, jm, tr)
        else:
            raise ValueError(kind)
