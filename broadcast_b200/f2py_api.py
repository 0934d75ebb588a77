"""Call-site signatures of the reference's f2py modules, independent of what executes them.

The reference drivers resolve Fortran routines by attribute name on f2py modules and call them
positionally (BROADCAST_npz.py:1018-1031,1072-1127,1242-1246; cylinder.py:857-978;
handleBC.py:129-242).  f2py turns dimension arguments that can be derived from array shapes into
optional trailing arguments; the wrappers below reproduce exactly that surface (SURVEY.md section
8(b)) and forward the FULL Fortran argument list, in Fortran order, to ``backend(name, *args)``.

``build(backend)`` returns a dict of module-like namespaces
``{'f_sch', 'f_lin', 'f_bnd', 'f_geom', 'f_norm', 'f_misc', 'f_init', 'f_dz', 'f_lindz'}``.
"""
from __future__ import annotations

import types

import numpy as np


def _interf(a):
    """integer(2,2) interface/window -> int32[4] = imin, jmin, imax, jmax (the drivers pass float arrays)."""
    a = np.asarray(a)
    if a.shape != (2, 2):
        raise ValueError("interface arrays must have shape (2,2): [[imin,jmin],[imax,jmax]]")
    return np.array([a[0, 0], a[0, 1], a[1, 0], a[1, 1]], dtype=np.int32)


def _loc(loc):
    if isinstance(loc, bytes):
        loc = loc.decode()
    loc = str(loc)[:3]
    if loc not in ("Ilo", "Ihi", "Jlo", "Jhi"):
        raise ValueError("loc must be one of 'Ilo','Ihi','Jlo','Jhi', got %r" % (loc,))
    return loc


def _state(a, name, ndim=3):
    """inout float64 Fortran-ordered array, as f2py demands for intent(inout)."""
    if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags.f_contiguous or a.ndim != ndim:
        raise ValueError(f"{name} must be a Fortran-contiguous float64 array of rank {ndim} (f2py intent(inout))")
    return a


def _in(a, name=None):
    """intent(in) float64 array: f2py copies/casts silently when needed."""
    if isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.f_contiguous:
        return a
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _check_cells(a, im, jm, gh, name, planes=5):
    want = (im + 2 * gh, jm + 2 * gh) + ((planes,) if planes else ())
    if tuple(a.shape) != want:
        raise ValueError(f"{name} has shape {tuple(a.shape)}, expected {want}")


def build(backend):
    B = backend

    # ------------------------------------------------------------------ f_sch / f_lin (schemes)
    def _scheme(name):
        def f(res, w, x0, y0, nx, ny, xc, yc, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4,
              im=None, jm=None):
            gh = int(gh)
            im = int(im) if im is not None else w.shape[0] - 2 * gh
            jm = int(jm) if jm is not None else w.shape[1] - 2 * gh
            _state(res, "residu")
            _check_cells(res, im, jm, gh, "residu")
            w = _in(w)
            _check_cells(w, im, jm, gh, "w")
            B(name, res, w, _in(x0), _in(y0), _in(nx), _in(ny), _in(xc), _in(yc), _in(vol), _in(volf), gh, cp, cv, prandtl,
              gam, rgaz, cs, muref, tref, s_suth, k2, k4, im, jm)
        f.__name__ = name
        return f

    def _scheme_d(name):
        def f(res, resd, w, wd, x0, y0, nx, ny, xc, yc, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth,
              k2, k4, im=None, jm=None):
            gh = int(gh)
            im = int(im) if im is not None else w.shape[0] - 2 * gh
            jm = int(jm) if jm is not None else w.shape[1] - 2 * gh
            _state(res, "residu")
            _state(resd, "residud")
            _check_cells(resd, im, jm, gh, "residud")
            w, wd = _in(w), _in(wd)
            _check_cells(w, im, jm, gh, "w")
            _check_cells(wd, im, jm, gh, "wd")
            B(name, res, resd, w, wd, _in(x0), _in(y0), _in(nx), _in(ny), _in(xc), _in(yc), _in(vol), _in(volf), gh, cp, cv,
              prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4, im, jm)
        f.__name__ = name
        return f

    def flux_num_dnc5_iso_2d(res, w, twall, x0, y0, nx, ny, xc, yc, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref,
                             s_suth, k2, k4, im=None, jm=None):
        """srcfv/rhs/flux_num_dnc5_iso.F90: the order-5 scheme with the isothermal wall flux (twall follows w)"""
        gh = int(gh)
        im = int(im) if im is not None else w.shape[0] - 2 * gh
        jm = int(jm) if jm is not None else w.shape[1] - 2 * gh
        _state(res, "residu")
        _check_cells(res, im, jm, gh, "residu")
        w = _in(w)
        _check_cells(w, im, jm, gh, "w")
        B("flux_num_dnc5_iso_2d", res, w, twall, _in(x0), _in(y0), _in(nx), _in(ny), _in(xc), _in(yc), _in(vol), _in(volf), gh, cp, cv,
          prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4, im, jm)

    def flux_num_dnc5_iso_2d_d(res, resd, w, wd, twall, x0, y0, nx, ny, xc, yc, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref,
                               tref, s_suth, k2, k4, im=None, jm=None):
        """srcfv/tangent/flux_num_dnc5_iso_d.f90 (twall passive)"""
        gh = int(gh)
        im = int(im) if im is not None else w.shape[0] - 2 * gh
        jm = int(jm) if jm is not None else w.shape[1] - 2 * gh
        _state(res, "residu")
        _state(resd, "residud")
        _check_cells(resd, im, jm, gh, "residud")
        w, wd = _in(w), _in(wd)
        _check_cells(w, im, jm, gh, "w")
        _check_cells(wd, im, jm, gh, "wd")
        B("flux_num_dnc5_iso_2d_d", res, resd, w, wd, twall, _in(x0), _in(y0), _in(nx), _in(ny), _in(xc), _in(yc), _in(vol), _in(volf), gh,
          cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4, im, jm)

    f_sch = types.SimpleNamespace(
        flux_num_dnc5_2d=_scheme("flux_num_dnc5_2d"),
        flux_num_dnc5_nowall_2d=_scheme("flux_num_dnc5_nowall_2d"),
        flux_num_dnc5_iso_2d=flux_num_dnc5_iso_2d,
        # the other orders of the scheme family (srcfv/rhs/flux_num_dnc{3,7,9}.F90 and their _nowall variants); gh = (order + 1) / 2
        **{f"flux_num_dnc{o}{v}_2d": _scheme(f"flux_num_dnc{o}{v}_2d") for o in (3, 7, 9) for v in ("", "_nowall")},
    )

    # ------------------------------------------------------------------ f_bnd (primal boundary fills)
    def bc_wall_viscous_adia_2d(w, loc, gam, interf, gh, im, jm):
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        B("bc_wall_viscous_adia_2d", w, _loc(loc), gam, _interf(interf), int(gh), int(im), int(jm))

    def bc_wall_viscous_iso_2d(w, twall, loc, gam, rgaz, interf, gh, im, jm):
        """srcfv/borders/bc_wall_viscous_iso.F90:1 (w, twall, loc, gam, rgaz, interf, gh, im, jm)"""
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        B("bc_wall_viscous_iso_2d", w, float(twall), _loc(loc), gam, rgaz, _interf(interf), int(gh), int(im), int(jm))

    def bc_symmetry_2d(w, loc, interf, nx, ny, gh, im, jm):
        """srcfv/borders/bc_symmetry.F90:1; call site handleBC.py:231 fsym(w, loc, interf, nx, ny, gh, im, jm)"""
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        B("bc_symmetry_2d", w, _loc(loc), _interf(interf), _in(nx), _in(ny), int(gh), int(im), int(jm))

    def bc_antisymmetry_2d(w, loc, interf, nx, ny, gh, im, jm):
        """srcfv/borders/bc_antisymmetry.F90:1"""
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        B("bc_antisymmetry_2d", w, _loc(loc), _interf(interf), _in(nx), _in(ny), int(gh), int(im), int(jm))

    def bc_pressure_2d(w, loc, interf, pext, noref, gam, nx, ny, im, jm, gh, em=None):
        """srcfv/borders/bc_pressure.F90:1 (w, loc, interf, pext, noref, gam, nx, ny, im, jm, gh[, em])"""
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        B("bc_pressure_2d", w, _loc(loc), _interf(interf), float(pext), int(bool(noref)), gam, _in(nx), _in(ny), int(im), int(jm), int(gh),
          int(em) if em is not None else w.shape[2])

    def _prof(a, name):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
        if a.size < 1:
            raise ValueError(f"{name}: empty profile")
        return a

    def bc_wall_blow_profile_2d(w, velprof, loc, gam, interf, gh, im, jm, lm=None):
        """srcfv/borders/bc_wall_blow_profile.F90:1 (w, velprof, loc, gam, interf, gh, im, jm[, lm])"""
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        vp = _prof(velprof, "velprof")
        B("bc_wall_blow_profile_2d", w, vp, _loc(loc), gam, _interf(interf), int(gh), int(im), int(jm), int(lm) if lm is not None else vp.size)

    def bc_wall_viscous_iso_profile_2d(w, twallprof, loc, gam, rgaz, interf, gh, im, jm, lm=None):
        """srcfv/borders/bc_wall_viscous_iso_profile.F90:1 (w, twallprof, loc, gam, rgaz, interf, gh, im, jm[, lm])"""
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        tp = _prof(twallprof, "twallprof")
        B("bc_wall_viscous_iso_profile_2d", w, tp, _loc(loc), gam, rgaz, _interf(interf), int(gh), int(im), int(jm),
          int(lm) if lm is not None else tp.size)

    def bc_no_reflexion_2d(w, wbd, loc, interf, nx, ny, gam, gh, im, jm, lm=None):
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        wbd = _in(wbd)
        lm = int(lm) if lm is not None else wbd.shape[0]
        B("bc_no_reflexion_2d", w, wbd, _loc(loc), _interf(interf), _in(nx), _in(ny), gam, int(gh), int(im), int(jm), lm)

    def bc_supandsubinlet_2d(w, loc, interf, field, nx, ny, gam, im, jm, *extra):
        # handleBC.py:189-194 passes one extra trailing argument (gh); BROADCAST_npz.py:1018 does not
        field = _in(field)
        lm, gh = field.shape[0], field.shape[1]
        if len(extra) == 2:
            lm, gh = int(extra[0]), int(extra[1])
        elif len(extra) > 2:
            raise TypeError("bc_supandsubinlet_2d: too many arguments")
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        B("bc_supandsubinlet_2d", w, _loc(loc), _interf(interf), field, _in(nx), _in(ny), gam, int(im), int(jm), lm, gh)

    def bc_general_2d(w, loc, interf, field, gh, im, jm, lm=None, em=None, gh1=None):
        """srcfv/borders/bc_general.F90: ghost layers from the table field(lm, gh1, em)"""
        field = _in(field)
        if field.ndim != 3:
            raise ValueError("field must have shape (lm, gh1, em)")
        lm = int(lm) if lm is not None else field.shape[0]
        em = int(em) if em is not None else field.shape[2]
        gh1 = int(gh1) if gh1 is not None else field.shape[1]
        _check_cells(_state(w, "w"), im, jm, gh, "w", planes=em)
        B("bc_general_2d", w, _loc(loc), _interf(interf), field, int(gh), int(im), int(jm), lm, em, gh1)

    def bc_extrapolate_o2_2d(w, loc, interf, im, jm, gh, em=None):
        _state(w, "w")
        em = int(em) if em is not None else w.shape[2]
        _check_cells(w, im, jm, gh, "w", planes=em)
        B("bc_extrapolate_o2_2d", w, _loc(loc), _interf(interf), int(im), int(jm), int(gh), em)

    def jn_match_2d(wr, prr, gh1r, gh2r, gh3r, gh4r, imr, jmr, wd, prd, gh1d, gh2d, gh3d, gh4d, imd, jmd, tr, em=None):
        _state(wr, "wr")
        em = int(em) if em is not None else wr.shape[2]
        wd_ = wd if wd is wr else _in(wd)
        B("jn_match_2d", wr, _interf(prr), int(gh1r), int(gh2r), int(gh3r), int(gh4r), int(imr), int(jmr), wd_, _interf(prd),
          int(gh1d), int(gh2d), int(gh3d), int(gh4d), int(imd), int(jmd), np.asarray(tr, dtype=np.int32), em)

    def jn_match_geom_2d(wr, prr, gh1r, gh2r, gh3r, gh4r, imr, jmr, wd, prd, gh1d, gh2d, gh3d, gh4d, imd, jmd, tr):
        _state(wr, "wr", ndim=2)
        wd_ = wd if wd is wr else _in(wd)
        B("jn_match_geom_2d", wr, _interf(prr), int(gh1r), int(gh2r), int(gh3r), int(gh4r), int(imr), int(jmr), wd_,
          _interf(prd), int(gh1d), int(gh2d), int(gh3d), int(gh4d), int(imd), int(jmd), np.asarray(tr, dtype=np.int32))

    f_bnd = types.SimpleNamespace(
        bc_wall_viscous_adia_2d=bc_wall_viscous_adia_2d, bc_no_reflexion_2d=bc_no_reflexion_2d,
        bc_wall_viscous_iso_2d=bc_wall_viscous_iso_2d, bc_symmetry_2d=bc_symmetry_2d, bc_antisymmetry_2d=bc_antisymmetry_2d,
        bc_pressure_2d=bc_pressure_2d, bc_wall_blow_profile_2d=bc_wall_blow_profile_2d,
        bc_wall_viscous_iso_profile_2d=bc_wall_viscous_iso_profile_2d,
        bc_supandsubinlet_2d=bc_supandsubinlet_2d, bc_extrapolate_o2_2d=bc_extrapolate_o2_2d, bc_general_2d=bc_general_2d,
        jn_match_2d=jn_match_2d, jn_match_geom_2d=jn_match_geom_2d)

    # ------------------------------------------------------------------ f_lin (tangents)
    def bc_wall_viscous_adia_2d_d(w, wd, loc, gam, interf, gh, im, jm):
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        B("bc_wall_viscous_adia_2d_d", w, wd, _loc(loc), gam, _interf(interf), int(gh), int(im), int(jm))

    def bc_wall_viscous_iso_2d_d(w, wd, twall, loc, gam, rgaz, interf, gh, im, jm):
        """srcfv/tangent/bc_wall_viscous_iso_d.f90 (w, wd, twall, loc, gam, rgaz, interf, gh, im, jm)"""
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        B("bc_wall_viscous_iso_2d_d", w, wd, float(twall), _loc(loc), gam, rgaz, _interf(interf), int(gh), int(im), int(jm))

    def bc_symmetry_2d_d(w, wd, loc, interf, nx, ny, gh, im, jm):
        """srcfv/tangent/bc_symmetry_d.f90; call site handleBC.py:230 flinsym(w, wd, loc, interf, nx, ny, gh, im, jm)"""
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        B("bc_symmetry_2d_d", w, wd, _loc(loc), _interf(interf), _in(nx), _in(ny), int(gh), int(im), int(jm))

    def bc_antisymmetry_2d_d(w, wd, loc, interf, nx, ny, gh, im, jm):
        """srcfv/tangent/bc_antisymmetry_d.f90"""
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        B("bc_antisymmetry_2d_d", w, wd, _loc(loc), _interf(interf), _in(nx), _in(ny), int(gh), int(im), int(jm))

    def bc_pressure_2d_d(w, wd, loc, interf, pext, noref, gam, nx, ny, im, jm, gh, em=None):
        """srcfv/tangent/bc_pressure_d.f90 (w, wd0, loc, interf, pext, noref, gam, nx, ny, im, jm, gh[, em])"""
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        B("bc_pressure_2d_d", w, wd, _loc(loc), _interf(interf), float(pext), int(bool(noref)), gam, _in(nx), _in(ny), int(im), int(jm),
          int(gh), int(em) if em is not None else w.shape[2])

    def bc_wall_blow_profile_2d_d(w, wd, velprof, velprofd, loc, gam, *rest):
        """srcfv/tangent/bc_wall_blow_profile_d.f90 (w, wd, velprof, velprofd, loc, gam, gamd, interf, gh, im, jm[, lm]); the sensitivity
        driver calls it WITHOUT gamd (BROADCAST_npz_sens.py:1763: ..., 'Jlo', gam, interf3, gh, im, jm): both arities are accepted,
        a missing gamd is 0"""
        gamd, rest = (0.0, rest) if np.ndim(rest[0]) > 0 else (float(rest[0]), rest[1:])
        interf, gh, im, jm = rest[:4]
        lm = rest[4] if len(rest) > 4 else None
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        vp, vpd = _prof(velprof, "velprof"), _prof(velprofd, "velprofd")
        B("bc_wall_blow_profile_2d_d", w, wd, vp, vpd, _loc(loc), gam, gamd, _interf(interf), int(gh), int(im), int(jm),
          int(lm) if lm is not None else vp.size)

    def bc_wall_viscous_iso_profile_2d_d(w, wd, twallprof, twallprofd, loc, gam, *rest):
        """srcfv/tangent/bc_wall_viscous_iso_profile_d.f90 (w, wd, twallprof, twallprofd, loc, gam, gamd, rgaz, rgazd, interf, gh, im,
        jm[, lm]); the driver-style call without gamd / rgazd (..., loc, gam, rgaz, interf, gh, im, jm) is accepted too"""
        if np.ndim(rest[1]) > 0:
            gamd, rgaz, rgazd, rest = 0.0, float(rest[0]), 0.0, rest[1:]
        else:
            gamd, rgaz, rgazd, rest = float(rest[0]), float(rest[1]), float(rest[2]), rest[3:]
        interf, gh, im, jm = rest[:4]
        lm = rest[4] if len(rest) > 4 else None
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        tp, tpd = _prof(twallprof, "twallprof"), _prof(twallprofd, "twallprofd")
        B("bc_wall_viscous_iso_profile_2d_d", w, wd, tp, tpd, _loc(loc), gam, gamd, rgaz, rgazd, _interf(interf), int(gh), int(im), int(jm),
          int(lm) if lm is not None else tp.size)

    def bc_no_reflexion_2d_d(w, wd, wbd, loc, interf, nx, ny, gam, gh, im, jm, lm=None):
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        wbd = _in(wbd)
        lm = int(lm) if lm is not None else wbd.shape[0]
        B("bc_no_reflexion_2d_d", w, wd, wbd, _loc(loc), _interf(interf), _in(nx), _in(ny), gam, int(gh), int(im), int(jm), lm)

    def bc_supandsubinlet_2d_d(w, wd, loc, interf, field, nx, ny, gam, im, jm, *extra):
        field = _in(field)
        lm, gh = field.shape[0], field.shape[1]
        if len(extra) == 2:
            lm, gh = int(extra[0]), int(extra[1])
        elif len(extra) > 2:
            raise TypeError("bc_supandsubinlet_2d_d: too many arguments")
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        B("bc_supandsubinlet_2d_d", w, wd, _loc(loc), _interf(interf), field, _in(nx), _in(ny), gam, int(im), int(jm), lm, gh)

    def bc_general_2d_d(w, wd, loc, interf, field, gh, im, jm, lm=None):
        """srcfv/tangent/bc_general_d.f90: the ghost tangents of the Dirichlet fill are zero"""
        field = _in(field)
        lm = int(lm) if lm is not None else field.shape[0]
        _check_cells(_state(w, "w"), im, jm, gh, "w")
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        B("bc_general_2d_d", w, wd, _loc(loc), _interf(interf), field, int(gh), int(im), int(jm), lm)

    def bc_extrapolate_o2_2d_d(w, wd, loc, interf, im, jm, gh, em=None):
        _state(w, "w")
        _state(wd, "wd")
        em = int(em) if em is not None else w.shape[2]
        B("bc_extrapolate_o2_2d_d", w, wd, _loc(loc), _interf(interf), int(im), int(jm), int(gh), em)

    f_lin = types.SimpleNamespace(
        flux_num_dnc5_2d_d=_scheme_d("flux_num_dnc5_2d_d"),
        flux_num_dnc5_nowall_2d_d=_scheme_d("flux_num_dnc5_nowall_2d_d"),
        flux_num_dnc5_iso_2d_d=flux_num_dnc5_iso_2d_d,
        # srcfv/tangent/flux_num_dnc{3,7,9}_d.f90, flux_num_dnc{3,7,9}_nowall_d.f90
        **{f"flux_num_dnc{o}{v}_2d_d": _scheme_d(f"flux_num_dnc{o}{v}_2d_d") for o in (3, 7, 9) for v in ("", "_nowall")},
        bc_wall_viscous_adia_2d_d=bc_wall_viscous_adia_2d_d, bc_no_reflexion_2d_d=bc_no_reflexion_2d_d,
        bc_wall_viscous_iso_2d_d=bc_wall_viscous_iso_2d_d, bc_symmetry_2d_d=bc_symmetry_2d_d,
        bc_antisymmetry_2d_d=bc_antisymmetry_2d_d, bc_pressure_2d_d=bc_pressure_2d_d,
        bc_wall_blow_profile_2d_d=bc_wall_blow_profile_2d_d, bc_wall_viscous_iso_profile_2d_d=bc_wall_viscous_iso_profile_2d_d,
        bc_supandsubinlet_2d_d=bc_supandsubinlet_2d_d, bc_extrapolate_o2_2d_d=bc_extrapolate_o2_2d_d,
        bc_general_2d_d=bc_general_2d_d)

    # ------------------------------------------------------------------ f_geom
    def computegeom_2d(x0, y0, nx, ny, xc, yc, vol, volf, im, jm, gh):
        for a, n, d in ((x0, "x0", 2), (y0, "y0", 2), (nx, "nx", 3), (ny, "ny", 3), (xc, "xc", 2), (yc, "yc", 2),
                        (vol, "vol", 2), (volf, "volf", 3)):
            _state(a, n, ndim=d)
        B("computegeom_2d", x0, y0, nx, ny, xc, yc, vol, volf, int(im), int(jm), int(gh))

    f_geom = types.SimpleNamespace(computegeom_2d=computegeom_2d)

    # ------------------------------------------------------------------ f_norm
    def compute_norml2inf(rhs, im, jm, gh):
        norm, ninf = np.zeros(5), np.zeros(5)
        B("compute_norml2inf", norm, ninf, _in(rhs), int(im), int(jm), int(gh))
        return norm, ninf

    def compute_norml2(rhs, im, jm, gh):
        norm, nmoy = np.zeros(5), np.zeros(5)
        B("compute_norml2", norm, nmoy, _in(rhs), int(im), int(jm), int(gh))
        return norm, nmoy

    f_norm = types.SimpleNamespace(compute_norml2inf=compute_norml2inf, compute_norml2=compute_norml2)

    # ------------------------------------------------------------------ f_misc (colouring / scatter)
    def _coo(jac, ia, ja):
        if not (isinstance(jac, np.ndarray) and jac.dtype == np.float64 and jac.ndim == 1 and jac.flags.c_contiguous):
            raise ValueError("jac must be a contiguous 1-D float64 array")
        for a, n in ((ia, "ia"), (ja, "ja")):
            if not (isinstance(a, np.ndarray) and a.dtype == np.int32 and a.ndim == 1 and a.flags.c_contiguous):
                raise ValueError(f"{n} must be a contiguous 1-D int32 array")
        if not (len(jac) == len(ia) == len(ja)):
            raise ValueError("jac, ia, ja must have the same length")
        return len(jac)

    def testvector(wd, m, l, k, gh, im, jm):
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        B("testvector", wd, int(m), int(l), int(k), int(gh), int(im), int(jm))

    def testvector_partial(wd, m, l, k, gh, im, jm, istart, iend, jstart, jend):
        _check_cells(_state(wd, "wd"), im, jm, gh, "wd")
        B("testvector_partial", wd, int(m), int(l), int(k), int(gh), int(im), int(jm), int(istart), int(iend), int(jstart),
          int(jend))

    def _dims_from(resd, gh, im, jm):
        gh = int(gh)
        im = int(im) if im is not None else resd.shape[0] - 2 * gh
        jm = int(jm) if jm is not None else resd.shape[1] - 2 * gh
        return gh, im, jm

    def computejacobianfromjv(jac, ia, ja, resd, m, l, k, gh, im, jm, nbentry=None):
        n = _coo(jac, ia, ja)
        gh, im, jm = _dims_from(resd, gh, im, jm)
        B("computejacobianfromjv", jac, ia, ja, _in(resd), int(m), int(l), int(k), gh, im, jm, n if nbentry is None else int(nbentry))

    def _relaxed(name):
        def f(jac, ia, ja, resd, m, l, k, gh, coefdiag, im=None, jm=None, nbentry=None):
            n = _coo(jac, ia, ja)
            coefdiag = _in(coefdiag)
            gh = int(gh)
            im = int(im) if im is not None else coefdiag.shape[0]
            jm = int(jm) if jm is not None else coefdiag.shape[1]
            B(name, jac, ia, ja, _in(resd), int(m), int(l), int(k), gh, im, jm, n if nbentry is None else int(nbentry), coefdiag)
        f.__name__ = name
        return f

    def computejacobianfromjv_relaxed_withjnandcheck(jac, ia, ja, resd, m, l, k, gh, coefdiag, mini, n, im=None, jm=None, nbentry=None):
        nb = _coo(jac, ia, ja)
        coefdiag = _in(coefdiag)
        gh = int(gh)
        im = int(im) if im is not None else coefdiag.shape[0]
        jm = int(jm) if jm is not None else coefdiag.shape[1]
        B("computejacobianfromjv_relaxed_withjnandcheck", jac, ia, ja, _in(resd), int(m), int(l), int(k), gh, im, jm,
          nb if nbentry is None else int(nbentry), coefdiag, float(mini), int(n))

    def computejacobianfromjv_withjn(jac, ia, ja, resd, m, l, k, gh, im, jm, nbentry=None):
        n = _coo(jac, ia, ja)
        gh, im, jm = _dims_from(resd, gh, im, jm)
        B("computejacobianfromjv_withjn", jac, ia, ja, _in(resd), int(m), int(l), int(k), gh, im, jm,
          n if nbentry is None else int(nbentry))

    def computejacobianfromdz(jac, ia, ja, dz, m, l, k, gh, im, jm, nbentry=None):
        n = _coo(jac, ia, ja)
        gh, im, jm = _dims_from(dz, gh, im, jm)
        B("computejacobianfromdz", jac, ia, ja, _in(dz), int(m), int(l), int(k), gh, im, jm, n if nbentry is None else int(nbentry))

    f_misc = types.SimpleNamespace(
        testvector=testvector, testvector_partial=testvector_partial, computejacobianfromjv=computejacobianfromjv,
        computejacobianfromjv_relaxed=_relaxed("computejacobianfromjv_relaxed"),
        computejacobianfromjv_relaxed_withjn=_relaxed("computejacobianfromjv_relaxed_withjn"),
        computejacobianfromjv_relaxed_withjnandcheck=computejacobianfromjv_relaxed_withjnandcheck,
        computejacobianfromjv_withjn=computejacobianfromjv_withjn, computejacobianfromdz=computejacobianfromdz)

    # ------------------------------------------------------------------ f_dz (spanwise operator rows)
    def _dz(name):
        def f(dz, w, wd, x0, y0, nx, ny, xc, yc, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth,
              im=None, jm=None):
            gh = int(gh)
            im = int(im) if im is not None else w.shape[0] - 2 * gh
            jm = int(jm) if jm is not None else w.shape[1] - 2 * gh
            _check_cells(_state(dz, "dz_out"), im, jm, gh, "dz_out")
            w, wd = _in(w), _in(wd)
            _check_cells(w, im, jm, gh, "w")
            _check_cells(wd, im, jm, gh, "wd")
            B(name, dz, w, wd, _in(x0), _in(y0), _in(nx), _in(ny), _in(xc), _in(yc), _in(vol), _in(volf), gh, cp, cv, prandtl,
              gam, rgaz, cs, muref, tref, s_suth, im, jm)
        f.__name__ = name
        return f

    f_dz = types.SimpleNamespace(coeffs_5p_dz=_dz("coeffs_5p_dz"), coeffs_5p_dz2=_dz("coeffs_5p_dz2"))

    # ------------------------------------------------------------------ f_lindz (tangent of the operator rows w.r.t. the base flow)
    # srcfv/tangentdz/coeffs_5p_dz_d.f90, coeffs_5p_dz2_d.f90; call sites BROADCAST_npz_sens.py:1768-1797, 2157-2185:
    #   f_lindz.coeffs_5p_dz_d(resd, dzr, w, wd, wmoder, x0, ..., im, jm)  ->  Fortran (dz_out, dz_outd, w, wd0, wd, ...)
    # dz_out is left untouched (Tapenade sliced its assignments), the WHOLE of dz_outd is written (ghost frame = 0).
    def _dz_d(name):
        def f(dz, dzd, w, wd0, wd, x0, y0, nx, ny, xc, yc, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth,
              im=None, jm=None):
            gh = int(gh)
            im = int(im) if im is not None else w.shape[0] - 2 * gh
            jm = int(jm) if jm is not None else w.shape[1] - 2 * gh
            _check_cells(_state(dz, "dz_out"), im, jm, gh, "dz_out")
            _check_cells(_state(dzd, "dz_outd"), im, jm, gh, "dz_outd")
            w, wd0, wd = _in(w), _in(wd0), _in(wd)
            _check_cells(w, im, jm, gh, "w")
            _check_cells(wd0, im, jm, gh, "wd0")
            _check_cells(wd, im, jm, gh, "wd")
            B(name, dz, dzd, w, wd0, wd, _in(x0), _in(y0), _in(nx), _in(ny), _in(xc), _in(yc), _in(vol), _in(volf), gh, cp, cv,
              prandtl, gam, rgaz, cs, muref, tref, s_suth, im, jm)
        f.__name__ = name
        return f

    f_lindz = types.SimpleNamespace(coeffs_5p_dz_d=_dz_d("coeffs_5p_dz_d"), coeffs_5p_dz2_d=_dz_d("coeffs_5p_dz2_d"))

    # ------------------------------------------------------------------ f_init (boundary tables from the initial field)
    def set_bndbl_2d(w, field, wbd, im, jm=None, gh=None):
        w = _in(w)
        _state(field, "field")
        _state(wbd, "wbd", ndim=2)
        gh = int(gh) if gh is not None else field.shape[1]
        jm = int(jm) if jm is not None else field.shape[0]
        B("set_bndbl_2d", w, field, wbd, int(im), jm, gh)

    f_init = types.SimpleNamespace(set_bndbl_2d=set_bndbl_2d)

    return dict(f_sch=f_sch, f_lin=f_lin, f_bnd=f_bnd, f_geom=f_geom, f_norm=f_norm, f_misc=f_misc, f_dz=f_dz, f_lindz=f_lindz, f_init=f_init)
