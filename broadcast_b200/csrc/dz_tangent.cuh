// Tangent of the spanwise operator rows with respect to the base flow (the sensitivity driver's f_lindz):
//   dz_outd (i,j,e) = d/dw [ sum_n coeffs_n(w)(e) * d func_n(w)[wd](e) ] . wd0
//                   = sum_n  coeffsd_n(e) * funcd_n(e)  +  coeffs_n(e) * funcdd_n(e)
// Reference: srcfv/tangentdz/coeffs_5p_dz_d.f90 (COEFFS_5P_DZ_D, FUNCTION_5P_DZ_D_D), coeffs_5p_dz2_d.f90 (Tapenade 3.16, tangent of
// the tangent), called from BROADCAST_npz_sens.py:1768-1797, 2157-2185 as f_lindz.coeffs_5p_dz_d(dz, dzd, w, wd0, wd, ...).
//
// Not a translation: the reference carries ~50 full-grid work arrays (func0..15, funcNd, funcNdd, coeffs(.,16), coeffsd) through seven
// sweeps.  Here ONE fused tile kernel evaluates the operator row in HYPER-DUAL arithmetic (value, d/d wd, d/d wd0, mixed second
// derivative): with w -> w + e1 wd + e2 wd0 every function of the table carries funcd in its e1 part and funcdd in its e1 e2 part, every
// coefficient carries coeffsd in its e2 part, so a term contributes  k.b * f.a + k.v * f.ab.  Unused components are dead code after
// inlining.  The same pass emits the d/dz rows (5-point cross stencil) and the d2/dz2 rows (cell-local): the reference calls them back to
// back on the same arguments, so the 15 input planes are read once.
//
// Tile: 32 x 8 cells, 256 threads.  Phase A: every thread turns its own cell into primitives (u, v, w, mu as hyper-duals -> shared
// memory, the rest stays in registers) and finishes the d2/dz2 row.  Phase B: 160 threads do the same for the cross-shaped halo (2 cells
// each side; corners are never read).  Phase C: 5-point gradients from shared memory (bank-conflict free: i is the fastest index),
// coefficient table, coalesced stores.  Algorithmic traffic per cell: 15 + 2 + 2 + 1 doubles read, 5 (+5) written = 200 (240) B.
// Host-compilable (tests/host/dz_tangent_host.cpp emulates the CTA phase by phase against oracle/_ref).
#pragma once
#include "grid.cuh"
#include <cmath>

namespace bcast {
namespace dzt {

struct HD {
  double v, a, b, ab;
};
BC_HD HD hd(double v) { return HD{v, 0.0, 0.0, 0.0}; }
BC_HD HD operator+(HD x, HD y) { return HD{x.v + y.v, x.a + y.a, x.b + y.b, x.ab + y.ab}; }
BC_HD HD operator-(HD x, HD y) { return HD{x.v - y.v, x.a - y.a, x.b - y.b, x.ab - y.ab}; }
BC_HD HD operator-(HD x) { return HD{-x.v, -x.a, -x.b, -x.ab}; }
BC_HD HD operator*(double s, HD x) { return HD{s * x.v, s * x.a, s * x.b, s * x.ab}; }
BC_HD HD operator+(HD x, double s) { return HD{x.v + s, x.a, x.b, x.ab}; }
BC_HD HD operator*(HD x, HD y) {
  return HD{x.v * y.v, x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b, x.ab * y.v + x.a * y.b + x.b * y.a + x.v * y.ab};
}
BC_HD HD recip(HD x) {
  const double r = 1.0 / x.v, r2 = r * r;
  return HD{r, -x.a * r2, -x.b * r2, (2.0 * x.a * x.b * r - x.ab) * r2};
}
// Tapenade convention (tangentdz/coeffs_5p_dz_d.f90: "IF (tloc .EQ. 0.0) result1d = 0.0"): sqrt'(0) = 0
BC_HD HD sqrt_hd(HD x) {
  const double s = ::sqrt(x.v);
  if (x.v == 0.0) return HD{s, 0.0, 0.0, 0.0};
  const double h = 0.5 / s;
  return HD{s, x.a * h, x.b * h, (x.ab - 0.5 * x.a * x.b / x.v) * h};
}

// one term of the row:  coefficient k (varies with the base flow), function f (tangent along wd, then along wd0)
BC_HD double term(HD k, HD f) { return k.b * f.a + k.v * f.ab; }

constexpr int TI = 32, TJ = 8, HALO = 2;
constexpr int SI = TI + 2 * HALO, SJ = TJ + 2 * HALO, NCELL = SI * SJ, NT = TI * TJ;
constexpr int NHALO = 2 * HALO * TI + 2 * HALO * TJ;  // cross-shaped halo, no corners: 160
enum { Q_U = 0, Q_V = 1, Q_W = 2, Q_MU = 3, NQ = 4 };
constexpr int NSM = NQ * 4 * NCELL;  // doubles of shared memory (55 296 B)

struct Consts {
  double cvm1, gm1, betas, s_suth, cpprandtl;
};
inline Consts make_dz_consts(double cp, double cv, double prandtl, double gam, double cs, double muref, double tref, double s_suth) {
  return Consts{1.0 / cv, gam - 1.0, muref * (tref + cs) / (std::sqrt(tref) * tref), s_suth, cp / prandtl};
}

struct Tile {
  GridDesc g;
  Consts c;
  const double* w;    // base flow, 5 planes
  const double* wa;   // wd  (the mode: first tangent direction)
  const double* wb;   // wd0 (the base-flow variation: second tangent direction)
  const double* nx;
  const double* ny;
  const double* vol;
  double* out1;       // dz_outd  (may be null)
  double* out2;       // dz2_outd (may be null)
  double* sm;         // NSM doubles
  int i0, j0;         // first interior cell of the tile (Fortran indices)
  int i1, j1;         // last cell to write
};

struct Cell {
  HD q[5], u, v, wz, t, p, mu;
};

// the 15 doubles of one cell (w, wd, wd0): loaded first, for the own cell AND the halo cell of a thread, so that both sets of loads
// are in flight before any arithmetic starts (the pass is latency bound on them at 16 warps per SM)
struct Raw {
  double w[5], a[5], b[5];
  bool ok;
};
BC_HD Raw load_raw(const Tile& t, bool ok, long long k) {
  Raw r;
  r.ok = ok;
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    r.w[e] = ok ? BC_LDG(t.w + e * t.g.sc + k) : 1.0;
    r.a[e] = ok ? BC_LDG(t.wa + e * t.g.sc + k) : 0.0;
    r.b[e] = ok ? BC_LDG(t.wb + e * t.g.sc + k) : 0.0;
  }
  return r;
}

BC_HD Cell cell_prims(const Tile& t, const Raw& in) {
  Cell r;
#pragma unroll
  for (int e = 0; e < 5; ++e) r.q[e] = HD{in.w[e], in.a[e], in.b[e], 0.0};
  const HD rom1 = recip(r.q[0]);
  r.u = r.q[1] * rom1;
  r.v = r.q[2] * rom1;
  r.wz = r.q[3] * rom1;
  const HD ec = 0.5 * (r.u * r.u + r.v * r.v + r.wz * r.wz);
  const HD eloc = (r.q[4] - ec * r.q[0]) * rom1;
  r.t = t.c.cvm1 * eloc;
  r.p = t.c.gm1 * (r.q[0] * eloc);
  r.mu = t.c.betas * (sqrt_hd(r.t) * r.t * recip(r.t + t.c.s_suth));
  return r;
}

BC_HD void sm_put(const Tile& t, int q, int s, HD x) {
  double* p = t.sm + (q * 4) * NCELL + s;
  p[0] = x.v;
  p[NCELL] = x.a;
  p[2 * NCELL] = x.b;
  p[3 * NCELL] = x.ab;
}
BC_HD HD sm_get(const Tile& t, int q, int s) {
  const double* p = t.sm + (q * 4) * NCELL + s;
  return HD{p[0], p[NCELL], p[2 * NCELL], p[3 * NCELL]};
}
BC_HD void sm_put_cell(const Tile& t, int s, const Cell& c) {
  sm_put(t, Q_U, s, c.u);
  sm_put(t, Q_V, s, c.v);
  sm_put(t, Q_W, s, c.wz);
  sm_put(t, Q_MU, s, c.mu);
}
// is staged cell (sx, sy) inside the padded array?  (partial tiles at the high ends run past it)
BC_HD bool staged_in_array(const Tile& t, int sx, int sy, long long* k) {
  const int ii = t.i0 - HALO + sx - 1 + t.g.gh, jj = t.j0 - HALO + sy - 1 + t.g.gh;
  *k = ii + (long long)jj * t.g.ldc;
  return ii >= 0 && ii < t.g.ni() && jj >= 0 && jj < t.g.nj();
}

// d2/dz2 row (srcfv/tangentdz/coeffs_5p_dz2_d.f90: coefficients 201-262, sum 270-306; functions matrix_dz2/function_dz2.F)
BC_HD void row_dz2(const Tile& t, const Cell& c, double (&r)[5]) {
  constexpr double FT = 2.0 * (2.0 / 3.0);
  const HD mmu = -c.mu;
  r[0] = 0.0;
  r[1] = term(mmu, c.u);
  r[2] = term(mmu, c.v);
  r[3] = term(FT * mmu, c.wz);
  r[4] = term(t.c.cpprandtl * mmu, c.t) + term(mmu * c.u, c.u) + term(mmu * c.v, c.v) + term(FT * (mmu * c.wz), c.wz);
}

// Phase A: own cell -> primitives; u, v, w, mu to shared memory; d2/dz2 row finished here.
BC_HD int own_cell(const Tile& t, int tid, long long* k) {   // staged index of the thread's own cell, -1 if outside the padded array
  const int tx = tid % TI, ty = tid / TI;
  return staged_in_array(t, tx + HALO, ty + HALO, k) ? (ty + HALO) * SI + tx + HALO : -1;
}
BC_HD Cell phase_a(const Tile& t, int tid, const Raw& in, long long k) {
  const int tx = tid % TI, ty = tid / TI;
  Cell c{};
  if (!in.ok) return c;
  c = cell_prims(t, in);
  if (t.out1) sm_put_cell(t, (ty + HALO) * SI + tx + HALO, c);   // the d2/dz2 rows alone are cell-local
  if (t.out2 && t.i0 + tx <= t.i1 && t.j0 + ty <= t.j1) {
    double r[5];
    row_dz2(t, c, r);
#pragma unroll
    for (int e = 0; e < 5; ++e) t.out2[e * t.g.sc + k] = r[e];
  }
  return c;
}

// Phase B: the cross-shaped halo (threads 0..159)
BC_HD int halo_cell(const Tile& t, int tid, long long* k) {   // staged index of the thread's halo cell, -1 if none
  if (tid >= NHALO) return -1;
  int sx, sy;
  if (tid < 2 * HALO * TI) {
    const int row = tid / TI;
    sy = row < HALO ? row : TJ + row;
    sx = HALO + tid % TI;
  } else {
    const int r = tid - 2 * HALO * TI, col = r % (2 * HALO);
    sy = HALO + r / (2 * HALO);
    sx = col < HALO ? col : TI + col;
  }
  return staged_in_array(t, sx, sy, k) ? sy * SI + sx : -1;
}
BC_HD void phase_b(const Tile& t, int s, const Raw& in) {
  if (in.ok) sm_put_cell(t, s, cell_prims(t, in));
}

// Phase C: d/dz row (srcfv/tangentdz/coeffs_5p_dz_d.f90; tables dz/coeffs_dz.F, matrix_dz/function_dz.F, gradients rhs/gradop_5pi.F,
// gradop_5pj.F, gradient.F with geom/dxdy.F)
struct Carry {   // what a thread keeps in registers across the barrier
  HD q1, q2, q3, q4, p;
};
BC_HD Carry carry_of(const Cell& c) { return Carry{c.q[1], c.q[2], c.q[3], c.q[4], c.p}; }
// metric factors of the 5-point gradient (geom/dxdy.F); loaded with the state, before the barrier
struct Met {
  double dxm1, dxm2, dym1, dym2;
};
BC_HD Met load_metrics(const Tile& t, int tid) {
  const int i = t.i0 + tid % TI, j = t.j0 + tid / TI;
  if (!t.out1 || i > t.i1 || j > t.j1) return Met{0.0, 0.0, 0.0, 0.0};
  const long long k = t.g.cidx(i, j), n = t.g.nidx(i, j);
  const double volm1 = 1.0 / BC_LDG(t.vol + k);
  return Met{0.5 * (BC_LDG(t.nx + n) + BC_LDG(t.nx + n + 1)) * volm1, 0.5 * (BC_LDG(t.nx + t.g.sn + n) + BC_LDG(t.nx + t.g.sn + n + t.g.ldn)) * volm1,
             0.5 * (BC_LDG(t.ny + n) + BC_LDG(t.ny + n + 1)) * volm1, 0.5 * (BC_LDG(t.ny + t.g.sn + n) + BC_LDG(t.ny + t.g.sn + n + t.g.ldn)) * volm1};
}
BC_HD void phase_c(const Tile& t, int tid, const Carry& c, const Met& m) {
  const int tx = tid % TI, ty = tid / TI;
  const int i = t.i0 + tx, j = t.j0 + ty;
  if (!t.out1 || i > t.i1 || j > t.j1) return;
  const long long k = t.g.cidx(i, j);
  const int s = (ty + HALO) * SI + tx + HALO;
  constexpr double b1 = 8.0 * (1.0 / 12.0), b2 = -(1.0 / 12.0), TT = 2.0 / 3.0;
  const double dxm1 = m.dxm1, dxm2 = m.dxm2, dym1 = m.dym1, dym2 = m.dym2;
  HD gx[NQ], gy[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const HD gi = b1 * (sm_get(t, q, s + 1) - sm_get(t, q, s - 1)) + b2 * (sm_get(t, q, s + 2) - sm_get(t, q, s - 2));
    const HD gj = b1 * (sm_get(t, q, s + SI) - sm_get(t, q, s - SI)) + b2 * (sm_get(t, q, s + 2 * SI) - sm_get(t, q, s - 2 * SI));
    gx[q] = dxm1 * gi + dxm2 * gj;
    gy[q] = dym1 * gi + dym2 * gj;
  }
  const HD gu0 = gx[Q_U], gv1 = gy[Q_V], gw0 = gx[Q_W], gw1 = gy[Q_W], gm0 = gx[Q_MU], gm1 = gy[Q_MU];
  const HD divu = gu0 + gv1;
  const HD u = sm_get(t, Q_U, s), v = sm_get(t, Q_V, s), wz = sm_get(t, Q_W, s), mu = sm_get(t, Q_MU, s);
  const HD mgw0 = mu * gw0, mgw1 = mu * gw1, muwz = mu * wz;
  double r[5];
  // coefficient ONE: the function's own mixed derivative
  r[0] = c.q3.ab;
  r[1] = (c.q1 * wz - mgw0).ab + term(TT * gm0, wz) + term(TT * mu, gw0);
  r[2] = (c.q2 * wz - mgw1).ab + term(TT * gm1, wz) + term(TT * mu, gw1);
  r[3] = (c.q3 * wz + c.p + TT * (mu * divu)).ab + term(-gm0, u) + term(-gm1, v) + term(-mu, divu);
  r[4] = ((c.q4 + c.p) * wz).ab                           //
         + term(TT * (gm0 * u + mu * gu0), wz)              // coeffs(5,2)  func1
         + term(TT * (mu * u), gw0)                         // coeffs(5,3)  func2
         + term(-(gm0 * wz + mgw0), u)                      // coeffs(5,4)  func3
         + term(-muwz, gu0)                                 // coeffs(5,5)  func4
         + term(TT * (gm1 * v + mu * gv1), wz)              // coeffs(5,6)  func5
         + term(TT * (mu * v), gw1)                         // coeffs(5,7)  func6
         + term(-(gm1 * wz + mgw1), v)                      // coeffs(5,8)  func7
         + term(-muwz, gv1)                                 // coeffs(5,9)  func8
         + term(-mgw0, u)                                   // coeffs(5,10) func9
         + term(-u, mgw0)                                   // coeffs(5,11) func10
         + term(-mgw1, v)                                   // coeffs(5,12) func11
         + term(-v, mgw1)                                   // coeffs(5,13) func12
         + term(TT * (mu * divu), wz)                       // coeffs(5,14) func13
         + term(TT * (wz * divu), mu)                       // coeffs(5,15) func14
         + term(TT * muwz, divu);                           // coeffs(5,16) func15
#pragma unroll
  for (int e = 0; e < 5; ++e) t.out1[e * t.g.sc + k] = r[e];
}

}  // namespace dzt
}  // namespace bcast
