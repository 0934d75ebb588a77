// Semi-analytic block Jacobian of the order-5 scheme on regular interior rows (FACE_MAIN faces, compact o4
// viscous fluxes), organised per FACE instead of per (row, column) pair:
//
//   producer  face_package<DIR>()   one face -> 44 doubles: everything NON-LINEAR about that face flux,
//             differentiated once (spectral radius and sensor by forward AD on a handful of scalars with the
//             Tapenade conventions of dual.cuh; viscous stresses by their closed form);
//   consumer  face_contrib<DIR,S,T>()  adds d hn / d (primitives, conservatives of stencil cell (S,T)) of one face
//             into the accumulators of a (row cell, column cell) block using only compile-time stencil weights,
//             because every face scalar is LINEAR in the cell primitives with constant coefficients;
//   chain     block_finish()        accumulators x d primitives / d conservatives of the column cell -> 5x5 block.
//
// Entry (e,m) of the block of row cell (i,j) and column cell (i+DI, j+DJ) is
//     -d residu(i,j,e) / d w(i+DI,j+DJ,m)  =  d[hn_i(i+1,j) - hn_i(i,j) + hn_j(i,j+1) - hn_j(i,j)]_e / d w_m
// i.e. what the reference's colour loop + computejacobianfromjv attribute to that pair
// (srcfv/tangent/flux_num_dnc5_d.f90, misc/ComputeJacobian.f90:518-569), at ~1/100 of the arithmetic: the
// reference re-evaluates the whole tangent for 245 seed vectors, the direct AD kernels (jac_blocks.cuh)
// re-evaluate the passive part of four faces for each of 29 column offsets.
//
// hn_e = fx_e - rspec (eps2 diff_e + eps4 pred_e) - visc_e       (scheme.cuh, face_flux<DIR,false,FACE_MAIN>)
#pragma once
#include "grid.cuh"
#include "scheme.cuh"

// the 29 structural offsets in column order (di major, dj minor): slot index = position in this list
#define BCAST_JAC_OFFSETS(X)                                                                                      \
  X(-3, 0) X(-2, -2) X(-2, -1) X(-2, 0) X(-2, 1) X(-2, 2) X(-1, -2) X(-1, -1) X(-1, 0) X(-1, 1) X(-1, 2) X(0, -3) \
  X(0, -2) X(0, -1) X(0, 0) X(0, 1) X(0, 2) X(0, 3) X(1, -2) X(1, -1) X(1, 0) X(1, 1) X(1, 2) X(2, -2) X(2, -1)   \
  X(2, 0) X(2, 1) X(2, 2) X(3, 0)

namespace bcast {

constexpr int JAC_NSLOT = 29;
constexpr int FPK_N = 54;
enum {
  // fields only the two or four column cells next to the face read (rare: they stay in global memory)
  FPK_SD = 0,     // [5] eps2 diff_e + eps4 pred_e              (x d rspec)
  FPK_DRS = 5,    // [2][5] d rspec / d w_m of along-cells -1, 0
  FPK_DEP = 15,   // [4] d eps2 / d p(s), s = -2..1
  FPK_DET = 19,   // [2] d eps2 / d T(s), s = -1, 0
  // fields the six cells of the face's along-line / the four cells of its face-value row read: staged too since r2_c (FPK_LS:
  // the compiler hoists their loads above the slot's branches, ~130 of the ~460 package loads per cell were these six fields)
  FPK_LS = 21,
  FPK_RSE2 = 21,  // rspec * eps2
  FPK_RSE4 = 22,  // rspec * eps4
  FPK_V1M = 23, FPK_V2M = 24, FPK_V3M = 25, FPK_V4M = 26,  // visc_e / mmu
  // "staged subset": the 27 fields that many column cells of a face's stencil read (14 of its 22 cells read SE and DEG, all
  // 20 cells of the viscous box read the last 14); the assembly kernel stages them in shared memory
  FPK_VS = 27,
  FPK_DEG = 27,   // [2][4] sensor cell s = -1, 0: coefficients of (wI, wJ) in d eps2/dU_c and in d eps2/dV_c
  FPK_SE = 35,    // [5] rspec (diff_e - 12 chi pred_e)          (x d eps2),  chi = [eps4 > 0]
  FPK_NXF = 40, FPK_NYF = 41,                              // face normal (length-scaled)
  FPK_MMU = 42, FPK_UU = 43, FPK_VV = 44, FPK_WW = 45,
  FPK_DNX = 46,   // [4] dual-cell normals x volm1: A+, A-, C+, C-   (x components)
  FPK_DNY = 50,   // [4]                                               (y components)
};
constexpr int FPK_NVS = 27;

// a single cell seen through the accessor interface (for flux_f / flux_g)
struct OneCell {
  const Var<Tan<5>>* q;
  const CellPrims<Tan<5>>* pp;
  template <int OI, int OJ> BC_HD Var<Tan<5>> W(int e) const { return q[e]; }
  template <int OI, int OJ> BC_HD Var<Tan<5>> U() const { return pp->u; }
  template <int OI, int OJ> BC_HD Var<Tan<5>> V() const { return pp->v; }
  template <int OI, int OJ> BC_HD Var<Tan<5>> Wz() const { return pp->w; }
  template <int OI, int OJ> BC_HD Var<Tan<5>> P() const { return pp->p; }
  template <int OI, int OJ> BC_HD Var<Tan<5>> H() const { return pp->h; }
};

BC_HD void seed_cell(const double (&w5)[5], Var<Tan<5>> (&q)[5]) {
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    q[e].v = w5[e];
#pragma unroll
    for (int m = 0; m < 5; ++m) q[e].d.d[m] = (e == m) ? 1.0 : 0.0;
  }
}

// 5-point gradient metric of a cell (geom/dxdy.F): gradient = (dxm1 d/di + dxm2 d/dj, dym1 d/di + dym2 d/dj)
template <int CI, int CJ, class A>
BC_HD void cell_metric(const A& a, double& dxm1, double& dxm2, double& dym1, double& dym2) {
  const double volm1 = 1.0 / a.template VOL<CI, CJ>();
  dxm1 = 0.5 * (a.template NX<CI, CJ>(0) + a.template NX<CI + 1, CJ>(0)) * volm1;
  dxm2 = 0.5 * (a.template NX<CI, CJ>(1) + a.template NX<CI, CJ + 1>(1)) * volm1;
  dym1 = 0.5 * (a.template NY<CI, CJ>(0) + a.template NY<CI + 1, CJ>(0)) * volm1;
  dym2 = 0.5 * (a.template NY<CI, CJ>(1) + a.template NY<CI, CJ + 1>(1)) * volm1;
}

// ---------------------------------------------------------------------------------------------
// producer: `a` is a PASSIVE accessor based at the face cell; out(field) = value
// ---------------------------------------------------------------------------------------------
template <int DIR, class A, class OUT>
BC_HD void face_package(const A& a, const SchemeConsts& c, OUT&& out) {
  constexpr double denom = 1.0 / 60.0;
  constexpr double d1 = 10.0 * denom, d2 = 5.0 * denom, d3 = denom;
  const double nxf = a.template NX<0, 0>(DIR);
  const double nyf = a.template NY<0, 0>(DIR);
  const double nx2 = nxf * nxf + nyf * nyf;

  double wr[5], wl[5];
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    wr[e] = a.template W<0, 0>(e).v;
    wl[e] = a.template W<AT(-1, 0)>(e).v;
  }
  const PVar tr = a.template T<0, 0>(), tl = a.template T<AT(-1, 0)>();

  // ---- spectral radius: value and gradient w.r.t. the conservative variables of the two face cells
  const PVar ur = a.template U<0, 0>(), vr = a.template V<0, 0>(), ul = a.template U<AT(-1, 0)>(), vl = a.template V<AT(-1, 0)>();
  const PVar rspec = spectral_radius(cst(wr[0]), ur, vr, tr, cst(wl[0]), ul, vl, tl, nxf, nyf, c);
  {
    Var<Tan<5>> q[5];
    seed_cell(wr, q);
    const CellPrims<Tan<5>> pp = cell_prims(q, c);
    const auto rs = spectral_radius(q[0], pp.u, pp.v, pp.t, cst(wl[0]), ul, vl, tl, nxf, nyf, c);
#pragma unroll
    for (int m = 0; m < 5; ++m) out(FPK_DRS + 5 + m, rs.d.d[m]);
  }
  {
    Var<Tan<5>> q[5];
    seed_cell(wl, q);
    const CellPrims<Tan<5>> pp = cell_prims(q, c);
    const auto rs = spectral_radius(cst(wr[0]), ur, vr, tr, q[0], pp.u, pp.v, pp.t, nxf, nyf, c);
#pragma unroll
    for (int m = 0; m < 5; ++m) out(FPK_DRS + m, rs.d.d[m]);
  }

  // ---- sensor: eps2 as a function of 10 face-level scalars, differentiated in vector forward mode
  using T10 = Tan<10>;
  auto seed = [](double v, int k) {
    Var<T10> r;
    r.v = v;
#pragma unroll
    for (int n = 0; n < 10; ++n) r.d.d[n] = (n == k) ? 1.0 : 0.0;
    return r;
  };
  const auto g0p = a.template GR<0, 0>();
  const auto g1p = a.template GR<AT(-1, 0)>();
  // directions: 0..3 p(-2..1), 4 T(-1), 5 T(0), 6 divu(-1), 7 vort(-1), 8 divu(0), 9 vort(0)   (vort = gv0 - gu1)
  const auto sn1 = sens_from_divu_vort(seed((g1p.u0 + g1p.v1).v, 6), seed((g1p.v0 - g1p.u1).v, 7));
  const auto sn0 = sens_from_divu_vort(seed((g0p.u0 + g0p.v1).v, 8), seed((g0p.v0 - g0p.u1).v, 9));
  const auto c2l = c.gam * c.rgaz * seed(tl.v, 4);
  const auto c2r = c.gam * c.rgaz * seed(tr.v, 5);
  const auto coef = sensor_coef(seed(a.template P<AT(-2, 0)>().v, 0), seed(a.template P<AT(-1, 0)>().v, 1), seed(a.template P<0, 0>().v, 2),
                                seed(a.template P<AT(1, 0)>().v, 3), sn0, sn1, a.template VOL<0, 0>(), a.template VOL<AT(-1, 0)>(), c2r, c2l,
                                nx2);
  const auto eps2 = c.k2 * coef;
  const auto eps4 = fmax(0.0, c.k4 - eps2 * 12.0);
  const bool chi = eps4.v > 0.0;

  out(FPK_RSE2, rspec.v * eps2.v);
  out(FPK_RSE4, rspec.v * eps4.v);
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    const double pred = -d3 * a.template W<AT(-3, 0)>(e).v + d2 * a.template W<AT(-2, 0)>(e).v - d1 * wl[e] + d1 * wr[e] -
                        d2 * a.template W<AT(1, 0)>(e).v + d3 * a.template W<AT(2, 0)>(e).v;
    const double diff = 0.5 * (wr[e] - wl[e]);
    out(FPK_SD + e, eps2.v * diff + eps4.v * pred);
    out(FPK_SE + e, rspec.v * (chi ? diff - 12.0 * pred : diff));
  }
#pragma unroll
  for (int n = 0; n < 4; ++n) out(FPK_DEP + n, eps2.d.d[n]);
  out(FPK_DET + 0, eps2.d.d[4]);
  out(FPK_DET + 1, eps2.d.d[5]);
  {
    double dxm1, dxm2, dym1, dym2;
    cell_metric<AT(-1, 0)>(a, dxm1, dxm2, dym1, dym2);
    const double dv = eps2.d.d[6], vo = eps2.d.d[7];
    out(FPK_DEG + 0, dv * dxm1 - vo * dym1);  // wI in d eps2 / dU_c
    out(FPK_DEG + 1, dv * dxm2 - vo * dym2);  // wJ in d eps2 / dU_c
    out(FPK_DEG + 2, dv * dym1 + vo * dxm1);  // wI in d eps2 / dV_c
    out(FPK_DEG + 3, dv * dym2 + vo * dxm2);  // wJ in d eps2 / dV_c
  }
  {
    double dxm1, dxm2, dym1, dym2;
    cell_metric<0, 0>(a, dxm1, dxm2, dym1, dym2);
    const double dv = eps2.d.d[8], vo = eps2.d.d[9];
    out(FPK_DEG + 4, dv * dxm1 - vo * dym1);
    out(FPK_DEG + 5, dv * dxm2 - vo * dym2);
    out(FPK_DEG + 6, dv * dym1 + vo * dxm1);
    out(FPK_DEG + 7, dv * dym2 + vo * dxm2);
  }

  // ---- viscous part
  const DualNormals dn = dual_normals<DIR>(a);
  const auto s = visc_scalars<DIR, false>(a, dn);
  constexpr double TWOTHIRD = 2.0 / 3.0;
  const double v1m = TWOTHIRD * (2.0 * s.ux.v - s.vy.v) * nxf + (s.uy.v + s.vx.v) * nyf;
  const double v2m = (s.uy.v + s.vx.v) * nxf + TWOTHIRD * (-s.ux.v + 2.0 * s.vy.v) * nyf;
  const double v3m = s.wx.v * nxf + s.wy.v * nyf;
  const double v4m = c.cpprandtl * (s.tx.v * nxf + s.ty.v * nyf) + s.uu.v * v1m + s.vv.v * v2m + s.ww.v * v3m;
  out(FPK_NXF, nxf);
  out(FPK_NYF, nyf);
  out(FPK_DNX + 0, dn.nAp_x * dn.volm1);
  out(FPK_DNX + 1, dn.nAm_x * dn.volm1);
  out(FPK_DNX + 2, dn.nCp_x * dn.volm1);
  out(FPK_DNX + 3, dn.nCm_x * dn.volm1);
  out(FPK_DNY + 0, dn.nAp_y * dn.volm1);
  out(FPK_DNY + 1, dn.nAm_y * dn.volm1);
  out(FPK_DNY + 2, dn.nCp_y * dn.volm1);
  out(FPK_DNY + 3, dn.nCm_y * dn.volm1);
  out(FPK_MMU, s.mmu.v);
  out(FPK_UU, s.uu.v);
  out(FPK_VV, s.vv.v);
  out(FPK_WW, s.ww.v);
  out(FPK_V1M, v1m);
  out(FPK_V2M, v2m);
  out(FPK_V3M, v3m);
  out(FPK_V4M, v4m);
}

// ---------------------------------------------------------------------------------------------
// consumer
// ---------------------------------------------------------------------------------------------
struct ColAcc {
  double gU[5], gV[5], gW[5], gT[5], gMu[5], gP[5];  // sum over faces of sgn * d hn_e / d primitive of the column cell
  double diag;                                       // coefficient of delta_em
  double Nx, Ny;                                     // Euler: sum of sgn * c(s) * face normal
  double dir[5][5];                                  // conservative-level rank-1 terms (SD x d rspec)
  BC_HD void clear() {
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      gU[e] = gV[e] = gW[e] = gT[e] = gMu[e] = gP[e] = 0.0;
#pragma unroll
      for (int m = 0; m < 5; ++m) dir[e][m] = 0.0;
    }
    diag = Nx = Ny = 0.0;
  }
};

// handle of one face's package: field f at pk[f * stride]; the staged subset (fields FPK_VS ..) may live in a second,
// faster array (shared memory in the assembly kernel): field FPK_VS + k at vs[k * vstride]
// ST = first staged field: fields >= ST are read from vs (field ST + k at vs[k * vstride]), the others from global memory
template <int ST>
struct FaceCtxT {
  const double* pk;
  long long stride;
  const double* vs;
  int vstride;
  BC_HD double operator()(int f) const { return BC_LDG(pk + f * stride); }
  BC_HD double v(int f) const { return f >= ST ? vs[(f - ST) * vstride] : BC_LDG(pk + f * stride); }
};
using FaceCtx = FaceCtxT<FPK_VS>;

namespace fj {
constexpr double kDen = 1.0 / 60.0;
constexpr double euler_c(int s) { return (s == 0 || s == -1) ? 37.0 * kDen : ((s == 1 || s == -2) ? -8.0 * kDen : ((s == 2 || s == -3) ? kDen : 0.0)); }
constexpr double pred_c(int s) {
  return s == -3 ? -kDen : s == -2 ? 5.0 * kDen : s == -1 ? -10.0 * kDen : s == 0 ? 10.0 * kDen : s == 1 ? -5.0 * kDen : s == 2 ? kDen : 0.0;
}
constexpr double diff_c(int s) { return s == 0 ? 0.5 : (s == -1 ? -0.5 : 0.0); }
constexpr double grad_w(int d) { return d == 1 ? 8.0 / 12.0 : d == -1 ? -8.0 / 12.0 : d == 2 ? -1.0 / 12.0 : d == -2 ? 1.0 / 12.0 : 0.0; }
// compact o4 interpolation weights (flux_visqueux_o4_{i,j}.F): along-rows r(s), cross m(t)
constexpr double o4_r(int s) { return (s == -2 || s == 1) ? -1.0 : ((s == -1 || s == 0) ? 9.0 : 0.0); }
constexpr double o4_mCm(int t) { return (t == -2 || t == 1) ? -1.0 : ((t == -1 || t == 0) ? 7.0 : 0.0); }
constexpr double o4_mCp(int t) { return (t == -1 || t == 2) ? -1.0 : ((t == 0 || t == 1) ? 7.0 : 0.0); }
constexpr double o4_Ap(int s, int t) { return t != 0 ? 0.0 : (s == 0 ? 26.0 / 24.0 : ((s == 1 || s == -1) ? -1.0 / 24.0 : 0.0)); }
constexpr double o4_Am(int s, int t) { return t != 0 ? 0.0 : (s == -1 ? 26.0 / 24.0 : ((s == 0 || s == -2) ? -1.0 / 24.0 : 0.0)); }
constexpr double kCross = (0.25 / 3.0) * 0.0625;
constexpr double o4_Cm(int s, int t) { return kCross * o4_r(s) * o4_mCm(t); }
constexpr double o4_Cp(int s, int t) { return kCross * o4_r(s) * o4_mCp(t); }
constexpr bool in_visc(int s, int t) { return s >= -2 && s <= 1 && t >= -2 && t <= 2; }
constexpr bool in_face(int s, int t) { return (t == 0 && s >= -3 && s <= 2) || in_visc(s, t); }
}  // namespace fj

// contribution of one face (direction DIR, sign sgn in the row balance) to the column cell at
// (along, cross) = (S, T) from the face cell
template <int DIR, int S, int T, class FC>
BC_HD void face_contrib(const FC& f, const SchemeConsts& c, double sgn, ColAcc& acc) {
  using namespace fj;
  if constexpr (!in_face(S, T)) {
    return;
  } else {
    if constexpr (T == 0 && S >= -3 && S <= 2) {
      constexpr double cE = euler_c(S), dd = diff_c(S), dp = pred_c(S);
      acc.Nx += sgn * cE * f.v(FPK_NXF);
      acc.Ny += sgn * cE * f.v(FPK_NYF);
      if constexpr (dd != 0.0)
        acc.diag -= sgn * (f(FPK_RSE2) * dd + f(FPK_RSE4) * dp);
      else
        acc.diag -= sgn * (f(FPK_RSE4) * dp);
    }
    if constexpr (T == 0 && (S == 0 || S == -1)) {
      double drs[5];
#pragma unroll
      for (int m = 0; m < 5; ++m) drs[m] = f(FPK_DRS + (S + 1) * 5 + m);
      const double det = f(FPK_DET + (S + 1));
#pragma unroll
      for (int e = 0; e < 5; ++e) {
        const double sd = sgn * f(FPK_SD + e);
#pragma unroll
        for (int m = 0; m < 5; ++m) acc.dir[e][m] -= sd * drs[m];
        acc.gT[e] -= sgn * f.v(FPK_SE + e) * det;
      }
    }
    if constexpr (T == 0 && S >= -2 && S <= 1) {
      const double dep = f(FPK_DEP + (S + 2));
#pragma unroll
      for (int e = 0; e < 5; ++e) acc.gP[e] -= sgn * f.v(FPK_SE + e) * dep;
    }
    // sensor: velocity gradients of the along-cells -1 and 0 (5-point crosses in GRID directions)
    {
      constexpr int daL = S + 1, daR = S;                 // along offset from sensor cell -1 / 0
      constexpr int diL = DIR == 0 ? daL : T, djL = DIR == 0 ? T : daL;
      constexpr int diR = DIR == 0 ? daR : T, djR = DIR == 0 ? T : daR;
      constexpr double wIL = djL == 0 ? grad_w(diL) : 0.0, wJL = diL == 0 ? grad_w(djL) : 0.0;
      constexpr double wIR = djR == 0 ? grad_w(diR) : 0.0, wJR = diR == 0 ? grad_w(djR) : 0.0;
      constexpr bool anyL = wIL != 0.0 || wJL != 0.0, anyR = wIR != 0.0 || wJR != 0.0;
      if constexpr (anyL || anyR) {
        double eU = 0.0, eV = 0.0;
        if constexpr (anyL) {
          eU += f.v(FPK_DEG + 0) * wIL + f.v(FPK_DEG + 1) * wJL;
          eV += f.v(FPK_DEG + 2) * wIL + f.v(FPK_DEG + 3) * wJL;
        }
        if constexpr (anyR) {
          eU += f.v(FPK_DEG + 4) * wIR + f.v(FPK_DEG + 5) * wJR;
          eV += f.v(FPK_DEG + 6) * wIR + f.v(FPK_DEG + 7) * wJR;
        }
#pragma unroll
        for (int e = 0; e < 5; ++e) {
          const double se = sgn * f.v(FPK_SE + e);
          acc.gU[e] -= se * eU;
          acc.gV[e] -= se * eV;
        }
      }
    }
    if constexpr (in_visc(S, T)) {
      constexpr double aAp = o4_Ap(S, T), aAm = o4_Am(S, T), aCp = o4_Cp(S, T), aCm = o4_Cm(S, T);
      constexpr double c0 = T == 0 ? o4_r(S) * 0.0625 : 0.0;
      double cx = aCp * f.v(FPK_DNX + 2) + aCm * f.v(FPK_DNX + 3);
      double cy = aCp * f.v(FPK_DNY + 2) + aCm * f.v(FPK_DNY + 3);
      if constexpr (aAp != 0.0) {
        cx += aAp * f.v(FPK_DNX + 0);
        cy += aAp * f.v(FPK_DNY + 0);
      }
      if constexpr (aAm != 0.0) {
        cx += aAm * f.v(FPK_DNX + 1);
        cy += aAm * f.v(FPK_DNY + 1);
      }
      const double mmu = f.v(FPK_MMU), uu = f.v(FPK_UU), vv = f.v(FPK_VV), ww = f.v(FPK_WW);
      const double nxf = f.v(FPK_NXF), nyf = f.v(FPK_NYF);
      const double a_ = nxf * cx, b_ = nyf * cy, c_ = nyf * cx, d_ = nxf * cy;
      constexpr double FT = 4.0 / 3.0, TT = 2.0 / 3.0;
      const double al = sgn * mmu * (FT * a_ + b_);   // d visc_1 / dU
      const double be = sgn * mmu * (c_ - TT * d_);   // d visc_1 / dV
      const double ga = sgn * mmu * (d_ - TT * c_);   // d visc_2 / dU
      const double de = sgn * mmu * (a_ + FT * b_);   // d visc_2 / dV
      const double ep = sgn * mmu * (a_ + b_);        // d visc_3 / dWz
      acc.gU[1] -= al;
      acc.gV[1] -= be;
      acc.gU[2] -= ga;
      acc.gV[2] -= de;
      acc.gW[3] -= ep;
      double g4U = uu * al + vv * ga, g4V = uu * be + vv * de, g4W = ww * ep;
      if constexpr (c0 != 0.0) {
        const double v1m = f(FPK_V1M), v2m = f(FPK_V2M), v3m = f(FPK_V3M), v4m = f(FPK_V4M);
        const double k = sgn * c0;
        g4U += k * mmu * v1m;
        g4V += k * mmu * v2m;
        g4W += k * mmu * v3m;
        acc.gMu[1] -= k * v1m;
        acc.gMu[2] -= k * v2m;
        acc.gMu[3] -= k * v3m;
        acc.gMu[4] -= k * v4m;
      }
      acc.gU[4] -= g4U;
      acc.gV[4] -= g4V;
      acc.gW[4] -= g4W;
      acc.gT[4] -= c.cpprandtl * ep;
    }
  }
}

// which of the four faces of a row cell see the column offset (DI, DJ), and where
template <int DI, int DJ>
struct BlockFaces {
  static constexpr bool iL = fj::in_face(DI, DJ);        // i-face (i,j):   (S,T) = (DI, DJ),   sign -
  static constexpr bool iR = fj::in_face(DI - 1, DJ);    // i-face (i+1,j): (S,T) = (DI-1, DJ), sign +
  static constexpr bool jL = fj::in_face(DJ, DI);        // j-face (i,j):   (S,T) = (DJ, DI),   sign -
  static constexpr bool jR = fj::in_face(DJ - 1, DI);    // j-face (i,j+1): (S,T) = (DJ-1, DI), sign +
};

// chain rule with the column cell: B[e][m] (row-major 25) = -d residu_e / d w_m
BC_HD void block_finish(const ColAcc& acc, const double (&wc)[5], const SchemeConsts& c, bool euler, double (&B)[25]) {
  Var<Tan<5>> q[5];
  seed_cell(wc, q);
  const CellPrims<Tan<5>> pp = cell_prims(q, c);
#pragma unroll
  for (int e = 0; e < 5; ++e)
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      double v = acc.dir[e][m] + acc.gU[e] * pp.u.d.d[m] + acc.gV[e] * pp.v.d.d[m] + acc.gW[e] * pp.w.d.d[m] + acc.gT[e] * pp.t.d.d[m] +
                 acc.gMu[e] * pp.mu.d.d[m] + acc.gP[e] * pp.p.d.d[m];
      if (e == m) v += acc.diag;
      B[e * 5 + m] = v;
    }
  if (euler) {
    const OneCell oc{q, &pp};
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      const auto F = flux_f<0, 0>(oc, e) * acc.Nx + flux_g<0, 0>(oc, e) * acc.Ny;
#pragma unroll
      for (int m = 0; m < 5; ++m) B[e * 5 + m] += F.d.d[m];
    }
  }
}

// whole block of row cell (i,j), column offset (DI,DJ).  fi0/fi1/fj0/fj1: contexts of faces i, i+1, j, j+1.
template <int DI, int DJ, class FC>
BC_HD void block_of(const FC& fi0, const FC& fi1, const FC& fj0, const FC& fj1, const double (&wc)[5],
                    const SchemeConsts& c, double (&B)[25]) {
  ColAcc acc;
  acc.clear();
  face_contrib<0, DI, DJ>(fi0, c, -1.0, acc);
  face_contrib<0, DI - 1, DJ>(fi1, c, 1.0, acc);
  face_contrib<1, DJ, DI>(fj0, c, -1.0, acc);
  face_contrib<1, DJ - 1, DI>(fj1, c, 1.0, acc);
  block_finish(acc, wc, c, DI == 0 || DJ == 0, B);
}

// ---------------------------------------------------------------------------------------------
// Table-driven consumer: the same arithmetic as face_contrib<DIR,S,T> / block_finish with the compile-time
// stencil weights moved into a constant table indexed by (face, column slot), so that the assembly kernel is ONE
// short runtime loop over the 29 column offsets (the fully unrolled template version is ~25 000 SASS instructions
// and stalls on instruction fetch).  All branches depend on (face, slot) only: they are warp-uniform.
// ---------------------------------------------------------------------------------------------
struct FaceTab {
  double cE, dd, dp;              // Euler / diff / pred weights of this cell in the face's along-line
  double wIL, wJL, wIR, wJR;      // 5-point gradient weights of this cell in the sensor cells -1 (L) and 0 (R)
  double aAp, aAm, aCp, aCm, c0;  // compact o4 interpolation weights (dual-cell sides A+, A-, C+, C-; face value)
  int side;                       // -1, or 0 / 1: this cell is the face's along-cell -1 / 0 (spectral radius, sensor T)
  int pk;                         // -1, or index 0..3 of this cell in the pressure sensor line
  int flags;                      // FT_* bits
  int pad;
};
enum { FT_ANY = 1, FT_LINE = 2, FT_SENS = 4, FT_VISC = 8, FT_C0 = 16 };
struct JacTab {
  FaceTab t[JAC_NSLOT][4];
  int di[JAC_NSLOT], dj[JAC_NSLOT];
  int euler[JAC_NSLOT];           // column cell on a grid axis through the row cell: Euler part present
  int needmu[JAC_NSLOT];          // some face sees the column cell in its face-value row (mu interpolation)
};

namespace fj {
constexpr FaceTab make_face_tab(int dir, int S, int T) {
  FaceTab t{};
  t.side = -1;
  t.pk = -1;
  if (!in_face(S, T)) return t;
  t.flags = FT_ANY;
  if (T == 0 && S >= -3 && S <= 2) {
    t.flags |= FT_LINE;
    t.cE = euler_c(S);
    t.dd = diff_c(S);
    t.dp = pred_c(S);
  }
  if (T == 0 && (S == 0 || S == -1)) t.side = S + 1;
  if (T == 0 && S >= -2 && S <= 1) t.pk = S + 2;
  const int daL = S + 1, daR = S;
  const int diL = dir == 0 ? daL : T, djL = dir == 0 ? T : daL;
  const int diR = dir == 0 ? daR : T, djR = dir == 0 ? T : daR;
  t.wIL = djL == 0 ? grad_w(diL) : 0.0;
  t.wJL = diL == 0 ? grad_w(djL) : 0.0;
  t.wIR = djR == 0 ? grad_w(diR) : 0.0;
  t.wJR = diR == 0 ? grad_w(djR) : 0.0;
  if (t.wIL != 0.0 || t.wJL != 0.0 || t.wIR != 0.0 || t.wJR != 0.0) t.flags |= FT_SENS;
  if (in_visc(S, T)) {
    t.flags |= FT_VISC;
    t.aAp = o4_Ap(S, T);
    t.aAm = o4_Am(S, T);
    t.aCp = o4_Cp(S, T);
    t.aCm = o4_Cm(S, T);
    if (T == 0) {
      t.c0 = o4_r(S) * 0.0625;
      t.flags |= FT_C0;
    }
  }
  return t;
}
constexpr JacTab make_jac_tab() {
  JacTab J{};
  int s = 0;
#define X(DI, DJ)                                   \
  J.di[s] = (DI);                                   \
  J.dj[s] = (DJ);                                   \
  J.euler[s] = ((DI) == 0 || (DJ) == 0) ? 1 : 0;    \
  J.t[s][0] = make_face_tab(0, (DI), (DJ));         \
  J.t[s][1] = make_face_tab(0, (DI)-1, (DJ));       \
  J.t[s][2] = make_face_tab(1, (DJ), (DI));         \
  J.t[s][3] = make_face_tab(1, (DJ)-1, (DI));       \
  J.needmu[s] = ((J.t[s][0].flags | J.t[s][1].flags | J.t[s][2].flags | J.t[s][3].flags) & FT_C0) ? 1 : 0; \
  ++s;
  BCAST_JAC_OFFSETS(X)
#undef X
  return J;
}
}  // namespace fj

// runtime-table version of face_contrib.  The viscous part is branch-free (table weights are zero for cells outside
// the 4x5 box) so that the loads of the four faces of a slot batch up; the rarer along-line / sensor terms branch.
template <class FC>
BC_HD void face_contrib_rt(const FC& f, const FaceTab& t, const SchemeConsts& c, double sgn, ColAcc& acc, double (&B)[25]) {
  if (!(t.flags & FT_ANY)) return;
  const double nxf = f.v(FPK_NXF), nyf = f.v(FPK_NYF);
  {
    const double cx = t.aCp * f.v(FPK_DNX + 2) + t.aCm * f.v(FPK_DNX + 3) + t.aAp * f.v(FPK_DNX + 0) + t.aAm * f.v(FPK_DNX + 1);
    const double cy = t.aCp * f.v(FPK_DNY + 2) + t.aCm * f.v(FPK_DNY + 3) + t.aAp * f.v(FPK_DNY + 0) + t.aAm * f.v(FPK_DNY + 1);
    const double mmu = sgn * f.v(FPK_MMU), uu = f.v(FPK_UU), vv = f.v(FPK_VV), ww = f.v(FPK_WW);
    const double a_ = nxf * cx, b_ = nyf * cy, c_ = nyf * cx, d_ = nxf * cy;
    constexpr double FT = 4.0 / 3.0, TT = 2.0 / 3.0;
    const double al = mmu * (FT * a_ + b_);
    const double be = mmu * (c_ - TT * d_);
    const double ga = mmu * (d_ - TT * c_);
    const double de = mmu * (a_ + FT * b_);
    const double ep = mmu * (a_ + b_);
    acc.gU[1] -= al;
    acc.gV[1] -= be;
    acc.gU[2] -= ga;
    acc.gV[2] -= de;
    acc.gW[3] -= ep;
    acc.gU[4] -= uu * al + vv * ga;
    acc.gV[4] -= uu * be + vv * de;
    acc.gW[4] -= ww * ep;
    acc.gT[4] -= c.cpprandtl * ep;
  }
  if (t.flags & FT_C0) {
    const double v1m = f.v(FPK_V1M), v2m = f.v(FPK_V2M), v3m = f.v(FPK_V3M), v4m = f.v(FPK_V4M);
    const double k = sgn * t.c0;
    const double km = k * f.v(FPK_MMU);
    acc.gU[4] -= km * v1m;
    acc.gV[4] -= km * v2m;
    acc.gW[4] -= km * v3m;
    acc.gMu[1] -= k * v1m;
    acc.gMu[2] -= k * v2m;
    acc.gMu[3] -= k * v3m;
    acc.gMu[4] -= k * v4m;
  }
  if (t.flags & FT_LINE) {
    acc.Nx += sgn * t.cE * nxf;
    acc.Ny += sgn * t.cE * nyf;
    acc.diag -= sgn * (f.v(FPK_RSE2) * t.dd + f.v(FPK_RSE4) * t.dp);
  }
  if (t.side >= 0 || t.pk >= 0 || (t.flags & FT_SENS)) {
    double se[5];
#pragma unroll
    for (int e = 0; e < 5; ++e) se[e] = sgn * f.v(FPK_SE + e);
    double eP = 0.0, eU = 0.0, eV = 0.0, eT = 0.0;
    if (t.pk >= 0) eP = f(FPK_DEP + t.pk);
    if (t.flags & FT_SENS) {
      eU = f.v(FPK_DEG + 0) * t.wIL + f.v(FPK_DEG + 1) * t.wJL + f.v(FPK_DEG + 4) * t.wIR + f.v(FPK_DEG + 5) * t.wJR;
      eV = f.v(FPK_DEG + 2) * t.wIL + f.v(FPK_DEG + 3) * t.wJL + f.v(FPK_DEG + 6) * t.wIR + f.v(FPK_DEG + 7) * t.wJR;
    }
    if (t.side >= 0) {
      eT = f(FPK_DET + t.side);
      double drs[5];
#pragma unroll
      for (int m = 0; m < 5; ++m) drs[m] = f(FPK_DRS + t.side * 5 + m);
#pragma unroll
      for (int e = 0; e < 5; ++e) {
        const double sd = sgn * f(FPK_SD + e);
#pragma unroll
        for (int m = 0; m < 5; ++m) B[e * 5 + m] -= sd * drs[m];
      }
    }
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      acc.gP[e] -= se[e] * eP;
      acc.gU[e] -= se[e] * eU;
      acc.gV[e] -= se[e] * eV;
      acc.gT[e] -= se[e] * eT;
    }
  }
}

// closed-form chain rule with the column cell (same result as block_finish, which differentiates cell_prims and
// the flux formulas by forward AD): d(u,v,w,T,p,mu)/d(conservatives) are sparse and cheap by hand.
// `B` enters holding the conservative-level rank-1 terms (SD x d rspec) and leaves holding the block.
BC_HD void block_finish_fast(const ColAcc& acc, const double (&wc)[5], const SchemeConsts& c, bool euler, bool needmu, double (&B)[25]) {
  const double ro = wc[0];
  const double r = 1.0 / ro;
  const double U = wc[1] * r, V = wc[2] * r, Wz = wc[3] * r;
  const double ec = 0.5 * (U * U + V * V + Wz * Wz);
  const double eloc = (wc[4] - ec * ro) * r;
  const double T = eloc * c.cvm1;
  double dmu = 0.0;   // d mu / d T  (Sutherland, phys/viscosity.F; sqrt'(0) = 0 as in the reference tangent)
  if (needmu) {
    const double s = ::sqrt(T);
    const double q = 1.0 / (T + c.s_suth);
    dmu = (T == 0.0) ? 0.0 : c.betas * s * q * (1.5 - T * q);
  }
  const double g1 = c.gam1;
  const double kTr = c.cvm1 * r;
  // d p / d w = g1 (ec, -U, -V, -Wz, 1);  d T / d w = cvm1 r (ec - eloc, -U, -V, -Wz, 1)
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    const double gT = acc.gT[e] + dmu * acc.gMu[e];
    const double kT = gT * kTr + acc.gP[e] * g1;
    const double gU = acc.gU[e], gV = acc.gV[e], gW = acc.gW[e];
    B[e * 5 + 0] += -r * (gU * U + gV * V + gW * Wz) + kT * ec - gT * kTr * eloc;
    B[e * 5 + 1] += r * gU - kT * U;
    B[e * 5 + 2] += r * gV - kT * V;
    B[e * 5 + 3] += r * gW - kT * Wz;
    B[e * 5 + 4] += kT;
    B[e * 6] += acc.diag;
  }
  if (euler) {
    // F_e = mn phi_e + p n_e,  mn = w1 Nx + w2 Ny,  phi = (1, U, V, Wz, H),  n = (0, Nx, Ny, 0, 0)
    const double Nx = acc.Nx, Ny = acc.Ny;
    const double P = g1 * ro * eloc;
    const double H = (wc[4] + P) * r;
    const double Vn = U * Nx + V * Ny;
    const double dP[5] = {g1 * ec, -g1 * U, -g1 * V, -g1 * Wz, g1};
    B[1] += Nx;
    B[2] += Ny;
    B[5] += -Vn * U + Nx * dP[0];
    B[6] += U * Nx + Vn + Nx * dP[1];
    B[7] += U * Ny + Nx * dP[2];
    B[8] += Nx * dP[3];
    B[9] += Nx * dP[4];
    B[10] += -Vn * V + Ny * dP[0];
    B[11] += V * Nx + Ny * dP[1];
    B[12] += V * Ny + Vn + Ny * dP[2];
    B[13] += Ny * dP[3];
    B[14] += Ny * dP[4];
    B[15] += -Vn * Wz;
    B[16] += Wz * Nx;
    B[17] += Wz * Ny;
    B[18] += Vn;
    B[20] += Vn * (dP[0] - H);
    B[21] += H * Nx + Vn * dP[1];
    B[22] += H * Ny + Vn * dP[2];
    B[23] += Vn * dP[3];
    B[24] += Vn * (1.0 + dP[4]);
  }
}

// block of column slot `s` through the table
template <class FC>
BC_HD void block_of_rt(const JacTab& J, int s, const FC& fi0, const FC& fi1, const FC& fj0, const FC& fj1,
                       const double (&wc)[5], const SchemeConsts& c, double (&B)[25]) {
  ColAcc acc;
  acc.clear();
#pragma unroll
  for (int q = 0; q < 25; ++q) B[q] = 0.0;
  face_contrib_rt(fi0, J.t[s][0], c, -1.0, acc, B);
  face_contrib_rt(fi1, J.t[s][1], c, 1.0, acc, B);
  face_contrib_rt(fj0, J.t[s][2], c, -1.0, acc, B);
  face_contrib_rt(fj1, J.t[s][3], c, 1.0, acc, B);
  block_finish_fast(acc, wc, c, J.euler[s] != 0, J.needmu[s] != 0, B);
}

}  // namespace bcast
