"""CPU oracle for the BROADCAST hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference algorithm (onera/Broadcast, Fortran behind f2py) for the
one path this repository accelerates: geometry, boundary fill, the order-5 FE-MUSCL ("dnc5")
residual, its forward-mode tangent, the colouring seeds and the COO scatter of the Jacobian,
and the residual norms.  Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module.  Nothing under ``broadcast_b200/`` does.

Tangents
--------
The reference tangents are Tapenade 3.16 output (``srcfv/tangent/*_d.f90``).  Here the same
primal code is evaluated on ``Dual`` arrays (value + one derivative) whose non-smooth
intrinsics follow the Tapenade conventions visible in the generated code
(``srcfv/tangent/flux_num_dnc5_d.f90``):

* ``abs``  : branch on ``x >= 0``                                  (:945-951)
* ``max``  : ``max(a,b)`` takes ``b`` iff ``a < b``                (:1054-1074), ``max(0,x)`` (:1117-1123)
* ``sqrt`` : derivative forced to 0 where the argument is 0        (:923-927)
* ``tanh`` : ``1 - tanh**2``                                       (:1025)
* ``sign`` : not differentiated (piecewise constant)               (tangent/bc_no_reflexion_d.f90)

This differs from the generated code only in floating-point association.

Pinning
-------
The reference ships no tests or golden vectors, and it cannot be built
as shipped (no Fortran compiler).  The oracle is pinned against ``oracle/_ref`` -- the
reference's own Fortran files machine-translated to C by ``oracle/f90_to_c.py`` and run in
this container -- by ``tests/test_oracle_vs_ref.py`` and by the golden fixtures under
``tests/golden/`` that ``oracle/make_golden.py`` generated from ``oracle/_ref``.

Array conventions are the reference's: float64, Fortran order, padded cell arrays
``(im+2gh, jm+2gh[,5])``, node/face arrays ``(im+2gh+1, jm+2gh+1[,2])``, Fortran index
``(i,j,e)`` <-> python ``[i+gh-1, j+gh-1, e-1]``.
"""
from __future__ import annotations

import numpy as np

# ======================================================================================
# dual numbers on arrays
# ======================================================================================


class Dual:
    """value + one tangent direction, elementwise on numpy arrays."""

    __array_ufunc__ = None
    __slots__ = ("v", "d")

    def __init__(self, v, d):
        self.v = v
        self.d = d

    @property
    def shape(self):
        return np.shape(self.v)

    def __getitem__(self, k):
        return Dual(self.v[k], self.d[k])

    def __setitem__(self, k, val):
        if isinstance(val, Dual):
            self.v[k] = val.v
            self.d[k] = val.d
        else:
            self.v[k] = val
            self.d[k] = 0.0

    def copy(self):
        return Dual(np.array(self.v, copy=True), np.array(self.d, copy=True))

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __add__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v + o.v, self.d + o.d)
        return Dual(self.v + o, self.d + 0.0 * o if np.ndim(o) > np.ndim(self.d) else self.d)

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v - o.v, self.d - o.d)
        return Dual(self.v - o, self.d)

    def __rsub__(self, o):
        return Dual(o - self.v, -self.d)

    def __mul__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v * o.v, self.d * o.v + self.v * o.d)
        return Dual(self.v * o, self.d * o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Dual):
            q = self.v / o.v
            return Dual(q, (self.d - q * o.d) / o.v)
        return Dual(self.v / o, self.d / o)

    def __rtruediv__(self, o):
        q = o / self.v
        return Dual(q, -(q * self.d) / self.v)

    # comparisons act on the values (Tapenade branches on the primal)
    def __lt__(self, o):
        return self.v < val(o)

    def __le__(self, o):
        return self.v <= val(o)

    def __gt__(self, o):
        return self.v > val(o)

    def __ge__(self, o):
        return self.v >= val(o)


def val(x):
    return x.v if isinstance(x, Dual) else x


def dot(x):
    return x.d if isinstance(x, Dual) else np.zeros_like(np.asarray(x, dtype=float))


def f_sqrt(x):
    if isinstance(x, Dual):
        r = np.sqrt(x.v)
        with np.errstate(divide="ignore", invalid="ignore"):
            d = np.where(x.v == 0.0, 0.0, x.d / (2.0 * r))
        return Dual(r, d)
    return np.sqrt(x)


def f_abs(x):
    if isinstance(x, Dual):
        pos = x.v >= 0.0
        return Dual(np.where(pos, x.v, -x.v), np.where(pos, x.d, -x.d))
    return np.abs(x)


def f_tanh(x):
    if isinstance(x, Dual):
        t = np.tanh(x.v)
        return Dual(t, (1.0 - t * t) * x.d)
    return np.tanh(x)


def f_max(a, b):
    """Fortran MAX(a,b) with the Tapenade branch: result is b iff a < b."""
    if isinstance(a, Dual) or isinstance(b, Dual):
        lt = val(a) < val(b)
        da = a.d if isinstance(a, Dual) else 0.0
        db = b.d if isinstance(b, Dual) else 0.0
        return Dual(np.where(lt, val(b), val(a)), np.where(lt, db, da))
    return np.where(a < b, b, a)


def f_pow(x, y):
    """x**y with a passive real exponent y (borders/bc_wall_viscous.F90:62)."""
    if isinstance(x, Dual):
        r = np.power(x.v, y)
        return Dual(r, y * np.power(x.v, y - 1.0) * x.d)
    return np.power(x, y)


def f_where(c, a, b):
    if isinstance(a, Dual) or isinstance(b, Dual):
        return Dual(np.where(c, val(a), val(b)), np.where(c, dot(a) if isinstance(a, Dual) else 0.0,
                                                          dot(b) if isinstance(b, Dual) else 0.0))
    return np.where(c, a, b)


def f_sign(a, b):
    """Fortran SIGN(a,b) for a passive magnitude a; never differentiated."""
    return np.copysign(np.abs(a), val(b))


# ======================================================================================
# geometry : srcfv/geom/computegeom.F90:3-104 (== srcfv/prepro/computegeom.f90)
# ======================================================================================


def computegeom_2d(x0, y0, nx, ny, xc, yc, vol, volf, im, jm, gh):
    """All arrays modified in place, like the f2py routine (BROADCAST_npz.py:702)."""
    o = gh - 1  # python index = fortran index + o

    # ghost extension of the nodes, two passes per layer (computegeom.F90:29-46)
    for g in range(1, gh + 1):
        for _ in range(2):
            x0[1 - g + o, :] = 2.0 * x0[2 - g + o, :] - x0[3 - g + o, :]
            x0[im + 1 + g + o, :] = 2.0 * x0[im + g + o, :] - x0[im - 1 + g + o, :]
            x0[:, 1 - g + o] = 2.0 * x0[:, 2 - g + o] - x0[:, 3 - g + o]
            x0[:, jm + 1 + g + o] = 2.0 * x0[:, jm + g + o] - x0[:, jm - 1 + g + o]
            y0[:, 1 - g + o] = 2.0 * y0[:, 2 - g + o] - y0[:, 3 - g + o]
            y0[:, jm + 1 + g + o] = 2.0 * y0[:, jm + g + o] - y0[:, jm - 1 + g + o]
            y0[1 - g + o, :] = 2.0 * y0[2 - g + o, :] - y0[3 - g + o, :]
            y0[im + 1 + g + o, :] = 2.0 * y0[im + g + o, :] - y0[im - 1 + g + o, :]

    # cells 1..im+1 x 1..jm+1 (computegeom.F90:49-59, geom/centers.F, volumes.F, normals_*.F)
    I = slice(1 + o, im + 1 + o + 1)
    Ip = slice(2 + o, im + 2 + o + 1)
    J = slice(1 + o, jm + 1 + o + 1)
    Jp = slice(2 + o, jm + 2 + o + 1)
    xa, ya = x0[I, J], y0[I, J]
    xb, yb = x0[Ip, J], y0[Ip, J]
    xc1, yc1 = x0[I, Jp], y0[I, Jp]
    xd, yd = x0[Ip, Jp], y0[Ip, Jp]
    xc[I, J] = 0.25 * (xa + xb + xc1 + xd)
    yc[I, J] = 0.25 * (ya + yb + yc1 + yd)
    abx, aby = xb - xa, yb - ya
    acx, acy = xc1 - xa, yc1 - ya
    q1 = 0.5 * np.abs(abx * acy - acx * aby)
    dcx, dcy = xc1 - xd, yc1 - yd
    dbx, dby = xb - xd, yb - yd
    q2 = 0.5 * np.abs(dcx * dby - dbx * dcy)
    vol[I, J] = q1 + q2
    nx[I, J, 0] = yc1 - ya
    ny[I, J, 0] = xa - xc1
    nx[I, J, 1] = ya - yb
    ny[I, J, 1] = xb - xa
    _computegeom_tail(x0, y0, nx, ny, xc, yc, vol, volf, im, jm, gh)


def _computegeom_tail(x0, y0, nx, ny, xc, yc, vol, volf, im, jm, gh):
    """Ghost extension of the metrics and volf (computegeom.F90:61-103); see the source for the
    exact statement order, reproduced here."""
    o = gh - 1
    for g in range(1, gh + 1):
        for _ in range(2):
            for a in (xc, yc, vol):
                # vol is extended in i over j=1:jm only (computegeom.F90:80-81)
                jr = slice(1 + o, jm + o + 1) if a is vol else slice(None)
                a[1 - g + o, jr] = 2.0 * a[2 - g + o, jr] - a[3 - g + o, jr]
                a[im + g + o, jr] = 2.0 * a[im - 1 + g + o, jr] - a[im - 2 + g + o, jr]
                a[:, 1 - g + o] = 2.0 * a[:, 2 - g + o] - a[:, 3 - g + o]
                a[:, jm + g + o] = 2.0 * a[:, jm - 1 + g + o] - a[:, jm - 2 + g + o]
            for a in (nx, ny):
                a[1 - g + o, :, :] = 2.0 * a[2 - g + o, :, :] - a[3 - g + o, :, :]
                a[im + 1 + g + o, :, :] = 2.0 * a[im + g + o, :, :] - a[im - 1 + g + o, :, :]
                a[:, 1 - g + o, :] = 2.0 * a[:, 2 - g + o, :] - a[:, 3 - g + o, :]
                a[:, jm + 1 + g + o, :] = 2.0 * a[:, jm + g + o, :] - a[:, jm - 1 + g + o, :]
    I = slice(1 + o, im + 1 + o + 1)
    Im = slice(o, im + o + 1)
    J = slice(1 + o, jm + 1 + o + 1)
    Jm = slice(o, jm + o + 1)
    volf[I, J, 0] = 2.0 / (vol[I, J] + vol[Im, J])
    volf[I, J, 1] = 2.0 / (vol[I, J] + vol[I, Jm])


# ======================================================================================
# residual : srcfv/rhs/flux_num_dnc5.F90:7-226  (== srcfv/prepro/flux_num_dnc5.f90:7-2590)
# tangent  : srcfv/tangent/flux_num_dnc5_d.f90:15-3870   (same code on Dual arrays)
# ======================================================================================


class _Grid:
    """index helper: Fortran inclusive ranges -> python slices on padded arrays."""

    def __init__(self, im, jm, gh):
        self.im, self.jm, self.gh = im, jm, gh
        self.o = gh - 1

    def sl(self, i0, i1, j0, j1):
        o = self.o
        return (slice(i0 + o, i1 + o + 1), slice(j0 + o, j1 + o + 1))


def _zeros_like_state(x, shape):
    if isinstance(x, Dual):
        return Dual(np.zeros(shape), np.zeros(shape))
    return np.zeros(shape)


def _primitives(w, cv, gam, betas, s_suth):
    """rhs/primvisc.F:2-9, phys/Primitives.F:2-34, phys/viscosity.F:1 -- all cells incl. ghosts.
    ``w`` is a list of the five conservative planes (ndarray or Dual)."""
    ro = w[0]
    rom1 = 1.0 / ro
    velx = w[1] * rom1
    vely = w[2] * rom1
    velz = w[3] * rom1
    ec = 0.5 * (velx * velx + vely * vely + velz * velz)
    eloc = (w[4] - ec * ro) * rom1
    tloc = eloc * (1.0 / cv)
    p = (gam - 1.0) * ro * eloc
    htot = (w[4] + p) * rom1
    f = [w[1], w[1] * velx + p, w[1] * vely, w[1] * velz, w[1] * htot]
    g = [w[2], w[2] * velx, w[2] * vely + p, w[2] * velz, w[2] * htot]
    mu = betas / (tloc + s_suth) * f_sqrt(tloc) * tloc
    return velx, vely, velz, tloc, p, mu, f, g


def _residual_core(w, nx, ny, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4,
                   im, jm, wall=True):
    """Returns the list of the five residual planes on the interior (shape (im, jm) each).
    ``w``: list of five padded planes (ndarray -> residual, Dual -> residual + tangent)."""
    G = _Grid(im, jm, gh)
    o = G.o
    HALF, ONE, TWO = 0.5, 1.0, 2.0
    TWOTHIRD = 2.0 / 3.0
    FOURTH = 0.25
    TWELFTH = 0.25 / 3.0
    cpprandtl = cp / prandtl
    TWENTYFOURTH = ONE / 24.0
    ccross = TWELFTH * 0.0625
    denom = 1.0 / 12.0
    b1 = 8.0 * denom
    b2 = -denom
    denom = 1.0 / 60.0
    c1, c2, c3 = 37.0 * denom, -8.0 * denom, denom
    d1, d2, d3 = 10.0 * denom, 5.0 * denom, denom
    betas = muref * (tref + cs) / (np.sqrt(tref) * tref)          # flux_num_dnc5.F90:120

    velx, vely, velz, tloc, p, mu, f, g = _primitives(w, cv, gam, betas, s_suth)
    shape = (im + 2 * gh, jm + 2 * gh)

    # ---- gradients on interior cells: gradop_5pi.F, gradop_5pj.F, gradient.F, geom/dxdy.F ----
    def sh(a, di, dj, i0=1, i1=im, j0=1, j1=jm):
        return a[G.sl(i0 + di, i1 + di, j0 + dj, j1 + dj)]

    def nsh(a, k, di, dj, i0=1, i1=im, j0=1, j1=jm):
        s = G.sl(i0 + di, i1 + di, j0 + dj, j1 + dj)
        return a[s[0], s[1], k]

    volm1 = ONE / sh(vol, 0, 0)
    dxm1 = HALF * (nsh(nx, 0, 0, 0) + nsh(nx, 0, 1, 0)) * volm1
    dxm2 = HALF * (nsh(nx, 1, 0, 0) + nsh(nx, 1, 0, 1)) * volm1
    dym1 = HALF * (nsh(ny, 0, 0, 0) + nsh(ny, 0, 1, 0)) * volm1
    dym2 = HALF * (nsh(ny, 1, 0, 0) + nsh(ny, 1, 0, 1)) * volm1
    grads = {}
    for name, q in (("u", velx), ("v", vely)):
        gi = b1 * (sh(q, 1, 0) - sh(q, -1, 0)) + b2 * (sh(q, 2, 0) - sh(q, -2, 0))
        gj = b1 * (sh(q, 0, 1) - sh(q, 0, -1)) + b2 * (sh(q, 0, 2) - sh(q, 0, -2))
        gx = _zeros_like_state(q, shape)
        gy = _zeros_like_state(q, shape)
        gx[G.sl(1, im, 1, jm)] = dxm1 * gi + dxm2 * gj
        gy[G.sl(1, im, 1, jm)] = dym1 * gi + dym2 * gj
        grads[name] = (gx, gy)

    # ---- rhs/gradveloingh.F:1-19 : linear extrapolation into the ghosts, two passes ----
    for _ in range(2):
        for h in range(1, gh + 1):
            for name in ("u", "v"):
                for a in grads[name]:
                    a[:, 1 - h + o] = TWO * a[:, 2 - h + o] - a[:, 3 - h + o]
                for a in grads[name]:
                    a[:, jm + h + o] = TWO * a[:, jm - 1 + h + o] - a[:, jm - 2 + h + o]
                for a in grads[name]:
                    a[1 - h + o, :] = TWO * a[2 - h + o, :] - a[3 - h + o, :]
                for a in grads[name]:
                    a[im + h + o, :] = TWO * a[im - 1 + h + o, :] - a[im - 2 + h + o, :]
    gradu, gradv = grads["u"], grads["v"]

    # ---- face-flux building blocks; every block works on a rectangle of faces (i0..i1, j0..j1) ----
    def face_blocks(i0, i1, j0, j1):
        def S(a, di=0, dj=0):
            return a[G.sl(i0 + di, i1 + di, j0 + dj, j1 + dj)]

        def N(a, k, di=0, dj=0):
            s = G.sl(i0 + di, i1 + di, j0 + dj, j1 + dj)
            return a[s[0], s[1], k]

        return S, N

    def euler_i(S, N):  # rhs/euler_o6_i.F:1-34
        return [(c1 * (S(f[e]) + S(f[e], -1)) + c2 * (S(f[e], 1) + S(f[e], -2)) + c3 * (S(f[e], 2) + S(f[e], -3))) * N(nx, 0)
                + (c1 * (S(g[e]) + S(g[e], -1)) + c2 * (S(g[e], 1) + S(g[e], -2)) + c3 * (S(g[e], 2) + S(g[e], -3))) * N(ny, 0)
                for e in range(5)]

    def euler_j(S, N):  # rhs/euler_o6_j.F
        return [(c1 * (S(f[e]) + S(f[e], 0, -1)) + c2 * (S(f[e], 0, 1) + S(f[e], 0, -2)) + c3 * (S(f[e], 0, 2) + S(f[e], 0, -3))) * N(nx, 1)
                + (c1 * (S(g[e]) + S(g[e], 0, -1)) + c2 * (S(g[e], 0, 1) + S(g[e], 0, -2)) + c3 * (S(g[e], 0, 2) + S(g[e], 0, -3))) * N(ny, 1)
                for e in range(5)]

    def euler_wall_j(S, N, cc, i0, i1):  # rhs/nearbndfluxes5demi_7p.F / nearbndfluxes3demi_7p.F (rows j=1..5 absolute)
        def row(a, jabs):
            return a[G.sl(i0, i1, jabs, jabs)]
        return [(cc[0] * row(f[e], 1) + cc[1] * row(f[e], 2) + cc[2] * row(f[e], 3) + cc[3] * row(f[e], 4) + cc[4] * row(f[e], 5)) * N(nx, 1)
                + (cc[0] * row(g[e], 1) + cc[1] * row(g[e], 2) + cc[2] * row(g[e], 3) + cc[3] * row(g[e], 4) + cc[4] * row(g[e], 5)) * N(ny, 1)
                for e in range(5)]

    def pred_i(S):  # rhs/predictor_7p_i.F
        return [-d3 * S(w[e], -3) + d2 * S(w[e], -2) - d1 * S(w[e], -1) + d1 * S(w[e]) - d2 * S(w[e], 1) + d3 * S(w[e], 2)
                for e in range(5)]

    def pred_j(S):  # rhs/predictor_7p_j.F
        return [-d3 * S(w[e], 0, -3) + d2 * S(w[e], 0, -2) - d1 * S(w[e], 0, -1) + d1 * S(w[e]) - d2 * S(w[e], 0, 1) + d3 * S(w[e], 0, 2)
                for e in range(5)]

    def visc_normals_i(N):
        nx_N = HALF * (N(nx, 0, 1, 0) + N(nx, 0))
        nx_S = -HALF * (N(nx, 0, -1, 0) + N(nx, 0))
        nx_O = HALF * (N(nx, 1, -1, 1) + N(nx, 1, 0, 1))
        nx_E = -HALF * (N(nx, 1, -1, 0) + N(nx, 1))
        ny_N = HALF * (N(ny, 0, 1, 0) + N(ny, 0))
        ny_S = -HALF * (N(ny, 0, -1, 0) + N(ny, 0))
        ny_O = HALF * (N(ny, 1, -1, 1) + N(ny, 1, 0, 1))
        ny_E = -HALF * (N(ny, 1, -1, 0) + N(ny, 1))
        return nx_N, nx_S, nx_O, nx_E, ny_N, ny_S, ny_O, ny_E

    def visc_normals_j(N):
        nx_N = HALF * (N(nx, 0, 1, -1) + N(nx, 0, 1, 0))
        nx_S = -HALF * (N(nx, 0, 0, -1) + N(nx, 0))
        nx_O = HALF * (N(nx, 1, 0, 1) + N(nx, 1))
        nx_E = -HALF * (N(nx, 1, 0, -1) + N(nx, 1))
        ny_N = HALF * (N(ny, 0, 1, -1) + N(ny, 0, 1, 0))
        ny_S = -HALF * (N(ny, 0, 0, -1) + N(ny, 0))
        ny_O = HALF * (N(ny, 1, 0, 1) + N(ny, 1))
        ny_E = -HALF * (N(ny, 1, 0, -1) + N(ny, 1))
        return nx_N, nx_S, nx_O, nx_E, ny_N, ny_S, ny_O, ny_E

    def visc_fluxes(grad, uu, vv, ww, mmu, lam):
        (ux, uy), (vx, vy), (wx, wy), (tx, ty) = grad
        fvrou = TWOTHIRD * mmu * (TWO * ux - vy)
        fvrov = mmu * (uy + vx)
        fvrow = mmu * wx
        fvroe = lam * tx + uu * fvrou + vv * fvrov + ww * fvrow
        gvrou = mmu * (uy + vx)
        gvrov = TWOTHIRD * mmu * (-ux + TWO * vy)
        gvrow = mmu * wy
        gvroe = lam * ty + uu * gvrou + vv * gvrov + ww * gvrow
        return (fvrou, fvrov, fvrow, fvroe), (gvrou, gvrov, gvrow, gvroe)

    def green(vals, nrm, volm1_):
        val_N, val_S, val_E, val_O = vals
        nx_N, nx_S, nx_O, nx_E, ny_N, ny_S, ny_O, ny_E = nrm
        qx = (val_N * nx_N + val_S * nx_S + val_O * nx_O + val_E * nx_E) * volm1_
        qy = (val_N * ny_N + val_S * ny_S + val_O * ny_O + val_E * ny_E) * volm1_
        return qx, qy

    def visc_o4_i(S, N):  # rhs/flux_visqueux_o4_i.F:8-103
        nrm = visc_normals_i(N)
        vm1 = N(volf, 0)
        grad = []
        for q in (velx, vely, velz, tloc):
            val_N = TWENTYFOURTH * (-S(q, 1, 0) + 26.0 * S(q, 0, 0) - S(q, -1, 0))
            val_S = TWENTYFOURTH * (-S(q, 0, 0) + 26.0 * S(q, -1, 0) - S(q, -2, 0))

            def psi(dj):
                return -S(q, -2, dj) + 9.0 * S(q, -1, dj) + 9.0 * S(q, 0, dj) - S(q, 1, dj)
            val_E = ccross * (-psi(-2) + 7.0 * psi(-1) + 7.0 * psi(0) - psi(1))
            val_O = ccross * (-psi(-1) + 7.0 * psi(0) + 7.0 * psi(1) - psi(2))
            grad.append(green((val_N, val_S, val_E, val_O), nrm, vm1))

        def face(q):
            return 0.0625 * (-S(q, -2, 0) + 9.0 * S(q, -1, 0) + 9.0 * S(q, 0, 0) - S(q, 1, 0))
        mmu = face(mu)
        return visc_fluxes(grad, face(velx), face(vely), face(velz), mmu, mmu * cpprandtl)

    def visc_o4_j(S, N):  # rhs/flux_visqueux_o4_j.F:9-106
        nrm = visc_normals_j(N)
        vm1 = N(volf, 1)
        grad = []
        for q in (velx, vely, velz, tloc):
            def psi(di):
                return -S(q, di, -2) + 9.0 * S(q, di, -1) + 9.0 * S(q, di, 0) - S(q, di, 1)
            val_N = ccross * (-psi(-1) + 7.0 * psi(0) + 7.0 * psi(1) - psi(2))
            val_S = ccross * (-psi(-2) + 7.0 * psi(-1) + 7.0 * psi(0) - psi(1))
            val_E = TWENTYFOURTH * (-S(q, 0, 0) + 26.0 * S(q, 0, -1) - S(q, 0, -2))
            val_O = TWENTYFOURTH * (-S(q, 0, 1) + 26.0 * S(q, 0, 0) - S(q, 0, -1))
            grad.append(green((val_N, val_S, val_E, val_O), nrm, vm1))

        def face(q):
            return 0.0625 * (-S(q, 0, -2) + 9.0 * S(q, 0, -1) + 9.0 * S(q, 0, 0) - S(q, 0, 1))
        mmu = face(mu)
        return visc_fluxes(grad, face(velx), face(vely), face(velz), mmu, mmu * cpprandtl)

    def visc_o2_i(S, N):  # rhs/flux_visqueux_o2_i.F:5-71
        nrm = visc_normals_i(N)
        vm1 = N(volf, 0)
        grad = []
        for q in (velx, vely, velz, tloc):
            val_N = S(q, 0, 0)
            val_S = S(q, -1, 0)
            val_E = FOURTH * (S(q, 0, 0) + S(q, 0, -1) + S(q, -1, 0) + S(q, -1, -1))
            val_O = FOURTH * (S(q, 0, 0) + S(q, 0, 1) + S(q, -1, 0) + S(q, -1, 1))
            grad.append(green((val_N, val_S, val_E, val_O), nrm, vm1))
        uu = HALF * (S(velx) + S(velx, -1, 0))
        vv = HALF * (S(vely) + S(vely, -1, 0))
        ww = HALF * (S(velz) + S(velz, -1, 0))
        mmu = HALF * (S(mu) + S(mu, -1, 0))
        lam = HALF * (S(mu) + S(mu, -1, 0)) * cpprandtl
        return visc_fluxes(grad, uu, vv, ww, mmu, lam)

    def visc_o2_j(S, N):  # rhs/flux_visqueux_o2_j.F:4-64
        nrm = visc_normals_j(N)
        vm1 = N(volf, 1)
        grad = []
        for q in (velx, vely, velz, tloc):
            val_N = FOURTH * (S(q, 0, 0) + S(q, 1, 0) + S(q, 0, -1) + S(q, 1, -1))
            val_S = FOURTH * (S(q, 0, 0) + S(q, -1, 0) + S(q, 0, -1) + S(q, -1, -1))
            val_O = S(q, 0, 0)
            val_E = S(q, 0, -1)
            grad.append(green((val_N, val_S, val_E, val_O), nrm, vm1))
        uu = HALF * (S(velx) + S(velx, 0, -1))
        vv = HALF * (S(vely) + S(vely, 0, -1))
        ww = HALF * (S(velz) + S(velz, 0, -1))
        mmu = HALF * (S(mu) + S(mu, 0, -1))
        lam = HALF * (S(mu) + S(mu, 0, -1)) * cpprandtl
        return visc_fluxes(grad, uu, vv, ww, mmu, lam)

    def dissipation(S, N, k, di, dj, pred):
        """rhs/dissipation_ducros_i.F / _j.F with spectralradius_{i,j}.F and ducrosfordnc_{i,j}.F.
        k = 0 (i-faces, neighbour (i-1,j)) or 1 (j-faces, neighbour (i,j-1)); (di,dj) = that offset."""
        rhomr = S(w[0])
        ur = S(w[1]) / rhomr
        vr = S(w[2]) / rhomr
        c2r = gam * rgaz * S(tloc)
        rhoml = S(w[0], di, dj)
        ul = S(w[1], di, dj) / rhoml
        vl = S(w[2], di, dj) / rhoml
        c2l = gam * rgaz * S(tloc, di, dj)
        r = f_sqrt(rhomr / rhoml)
        rr = ONE / (ONE + r)
        omrr = ONE - rr
        u = ul * rr + ur * omrr
        v = vl * rr + vr * omrr
        c2x = c2l * rr + c2r * omrr
        nx2 = N(nx, k) * N(nx, k) + N(ny, k) * N(ny, k)
        ab = f_abs(N(nx, k) * u + N(ny, k) * v)
        sq = f_sqrt(c2x * nx2)
        rspec = ab + sq
        # Jameson pressure sensor and Ducros sensor (ducrosfordnc_i.F:3-78)
        k_sensor1 = f_abs(S(p, di, dj) - TWO * S(p) + S(p, -di, -dj)) / f_abs(S(p, di, dj) + TWO * S(p) + S(p, -di, -dj))
        k_sensor2 = f_abs(S(p, 2 * di, 2 * dj) - TWO * S(p, di, dj) + S(p)) / f_abs(S(p, 2 * di, 2 * dj) + TWO * S(p, di, dj) + S(p))
        divu = S(gradu[0]) + S(gradv[1])
        divu2 = divu * divu
        vort2 = (S(gradv[0]) - S(gradu[1])) * (S(gradv[0]) - S(gradu[1]))
        ducros1 = divu2 / (divu2 + vort2 + 1e-15)
        dxm1_ = HALF * (ONE - f_tanh(2.5 + 10.0 * S(vol) / (f_sqrt(c2r * nx2) + 1e-15) * divu))
        divu = S(gradu[0], di, dj) + S(gradv[1], di, dj)
        divu2 = divu * divu
        vort2 = (S(gradv[0], di, dj) - S(gradu[1], di, dj)) * (S(gradv[0], di, dj) - S(gradu[1], di, dj))
        ducros2 = divu2 / (divu2 + vort2 + 1e-15)
        dxm2_ = HALF * (ONE - f_tanh(2.5 + 10.0 * S(vol, di, dj) / (f_sqrt(c2l * nx2) + 1e-15) * divu))
        coef = f_max(k_sensor1, k_sensor2) * f_max(ducros1, ducros2) * f_max(dxm1_, dxm2_)
        eps2 = k2 * coef
        eps4 = f_max(0.0, k4 - eps2 * 12.0)
        out = []
        for e in range(5):
            diff = HALF * (S(w[e]) - S(w[e], di, dj))
            out.append(rspec * (eps2 * diff + eps4 * pred[e]))
        return out

    def assemble(S, N, k, fx, diss, fv, gv):  # rhs/fluxnumassembly_i.F:1-16 / _j.F
        sc1 = N(nx, k)
        sc2 = N(ny, k)
        sn = np.sqrt(sc1 * sc1 + sc2 * sc2)
        invsn = ONE / sn
        nxloc = sc1 * invsn
        nyloc = sc2 * invsn
        hn = [fx[0] - diss[0]]
        for e in range(1, 5):
            hn.append(fx[e] - diss[e] - (fv[e - 1] * nxloc + gv[e - 1] * nyloc) * sn)
        return hn

    def iface(i0, i1, j0, j1, visc):
        S, N = face_blocks(i0, i1, j0, j1)
        fx = euler_i(S, N)
        pred = pred_i(S)
        fv, gv = (visc_o4_i if visc == 4 else visc_o2_i)(S, N)
        diss = dissipation(S, N, 0, -1, 0, pred)
        return assemble(S, N, 0, fx, diss, fv, gv)

    def jface(i0, i1, j0, j1, visc, wallcoef=None):
        S, N = face_blocks(i0, i1, j0, j1)
        fx = euler_j(S, N) if wallcoef is None else euler_wall_j(S, N, wallcoef, i0, i1)
        pred = pred_j(S)
        fv, gv = (visc_o4_j if visc == 4 else visc_o2_j)(S, N)
        diss = dissipation(S, N, 1, 0, -1, pred)
        return assemble(S, N, 1, fx, diss, fv, gv)

    # hn over faces i = 1..im+1, j = 1..jm+1 (only the faces the balance uses are filled)
    hshape = (im + 1, jm + 1)
    hn_i = [_zeros_like_state(w[0], hshape) for _ in range(5)]
    hn_j = [_zeros_like_state(w[0], hshape) for _ in range(5)]

    def put(dst, src, i0, j0):
        ni, nj = src[0].shape
        for e in range(5):
            dst[e][i0 - 1:i0 - 1 + ni, j0 - 1:j0 - 1 + nj] = src[e]

    jstart = 4 if wall else 1
    # main loop  (flux_num_dnc5.F90:139-157): j = jstart..jm+1, i = 1..im+1, both directions
    put(hn_i, iface(1, im + 1, jstart, jm, 4), 1, jstart)
    put(hn_j, jface(1, im, jstart, jm + 1, 4), 1, jstart)
    if wall:
        denom = 1.0 / 60.0
        c5 = (-3.0 * denom, 27.0 * denom, 47.0 * denom, -13.0 * denom, 2.0 * denom)   # coefnearbnd_7p.F:2-6
        c3_ = (12.0 * denom, 77.0 * denom, -43.0 * denom, 17.0 * denom, -3.0 * denom)  # coefnearbnd_7p.F:9-13
        # j = 3 (flux_num_dnc5.F90:165-177)
        put(hn_i, iface(1, im + 1, 3, 3, 4), 1, 3)
        put(hn_j, jface(1, im, 3, 3, 4, wallcoef=c5), 1, 3)
        # j = 2 (flux_num_dnc5.F90:178-192)
        put(hn_i, iface(1, im + 1, 2, 2, 2), 1, 2)
        put(hn_j, jface(1, im, 2, 2, 2, wallcoef=c3_), 1, 2)
        # j = 1 (flux_num_dnc5.F90:203-216): i-faces with o2 viscous flux, wall flux for the j-face
        put(hn_i, iface(1, im + 1, 1, 1, 2), 1, 1)
        S, N = face_blocks(1, im, 1, 1)
        ct0, ct1 = 1.125, -0.125
        pw = ct0 * S(p) + ct1 * S(p, 0, 1)                          # rhs/fluxwall.F:3-5
        mmu = S(mu)
        nxw, nyw, vfw = N(nx, 1), N(ny, 1), N(volf, 1)
        ux = TWO * S(velx) * nxw * vfw
        vx = TWO * S(vely) * nxw * vfw
        wx = TWO * S(velz) * nxw * vfw
        uy = TWO * S(velx) * nyw * vfw
        vy = TWO * S(vely) * nyw * vfw
        wy = TWO * S(velz) * nyw * vfw
        fvrou = TWOTHIRD * mmu * (TWO * ux - vy)
        fvrov = mmu * (uy + vx)
        fvrow = mmu * wx
        gvrou = mmu * (uy + vx)
        gvrov = TWOTHIRD * mmu * (-ux + TWO * vy)
        gvrow = mmu * wy
        zero = pw * 0.0
        put(hn_j, [zero,
                   pw * nxw - (fvrou * nxw + gvrou * nyw),
                   pw * nyw - (fvrov * nxw + gvrov * nyw),
                   zero - (fvrow * nxw + gvrow * nyw),
                   zero], 1, 1)

    # balance (rhs/balance.F:2-15)
    res = []
    for e in range(5):
        hi, hj = hn_i[e], hn_j[e]
        res.append(-(hi[1:im + 1, 0:jm] - hi[0:im, 0:jm]) - (hj[0:im, 1:jm + 1] - hj[0:im, 0:jm]))
    return res


def _planes(w):
    return [w[:, :, e] for e in range(5)]


def flux_num_dnc5_2d(residu, w, x0, y0, nx, ny, xc, yc, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref,
                     tref, s_suth, k2, k4, im, jm, wall=True):
    """rhs/flux_num_dnc5.F90:7-226.  Writes residu(1:im,1:jm,1:5) only (ghosts untouched)."""
    res = _residual_core(_planes(w), nx, ny, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth,
                         k2, k4, im, jm, wall=wall)
    for e in range(5):
        residu[gh:gh + im, gh:gh + jm, e] = res[e]


def flux_num_dnc5_nowall_2d(residu, w, *args):
    """rhs/flux_num_dnc5_nowall.F90:120-157 : main loop from j = 1, no wall rows."""
    flux_num_dnc5_2d(residu, w, *args, wall=False)


def flux_num_dnc5_2d_d(residu, residud, w, wd, x0, y0, nx, ny, xc, yc, vol, volf, gh, cp, cv, prandtl, gam, rgaz,
                       cs, muref, tref, s_suth, k2, k4, im, jm, wall=True):
    """tangent/flux_num_dnc5_d.f90:15-3870.  ``residud`` is zeroed entirely, then its interior is
    written (:3853-3868); ``residu`` is NOT written (dead code sliced by Tapenade)."""
    wl = [Dual(w[:, :, e], wd[:, :, e]) for e in range(5)]
    res = _residual_core(wl, nx, ny, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth,
                         k2, k4, im, jm, wall=wall)
    residud[...] = 0.0
    for e in range(5):
        residud[gh:gh + im, gh:gh + jm, e] = res[e].d


def flux_num_dnc5_nowall_2d_d(residu, residud, w, wd, *args):
    flux_num_dnc5_2d_d(residu, residud, w, wd, *args, wall=False)
