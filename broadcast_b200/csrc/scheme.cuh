// Order-5 FE-MUSCL ("dnc5") face fluxes of the BROADCAST finite-volume residual, written once for
// every arithmetic mode (passive, dense tangent, single-perturbed-cell tangent) and both grid
// directions.
//
// Reference (restated, not copied): srcfv/rhs/flux_num_dnc5.F90:7-226 and the fragments it includes:
//   euler_o6_{i,j}.F, predictor_7p_{i,j}.F, flux_visqueux_o4_{i,j}.F, flux_visqueux_o2_{i,j}.F,
//   spectralradius_{i,j}.F, ducrosfordnc_{i,j}.F, dissipation_ducros_{i,j}.F,
//   fluxnumassembly_{i,j}.F, nearbndfluxes{5,3}demi_7p.F, coefnearbnd_7p.F, fluxwall.F,
//   phys/Primitives.F, phys/viscosity.F, gradop_5p{i,j}.F, gradient.F, geom/dxdy.F.
//
// A face is addressed through an accessor `A` whose methods take COMPILE-TIME cell offsets relative
// to the face cell (i,j) (i-face (i,j) lies between cells (i-1,j) and (i,j); j-face between (i,j-1)
// and (i,j)).  `DIR` = 0 for i-faces, 1 for j-faces; AT(s,t) maps (along, cross) offsets to (di,dj).
// Operation order inside every formula follows the reference so that passive results agree with the
// Fortran to rounding of FMA contraction only.
#pragma once
#include "dual.cuh"
#include "grid.cuh"
#include <utility>

namespace bcast {

struct SchemeConsts {
  double gam, rgaz, cpprandtl, cvm1, betas, s_suth, k2, k4;
  double gam1;  // gam - 1
  // isothermal-wall variant of the scheme (flux_num_dnc5_iso.F90: rhs/fluxwall_iso.F instead of rhs/fluxwall.F): the wall face
  // carries a heat flux towards the wall temperature twall
  double twall;
  int wall_iso;
};

// Which wall flux the calling HOST thread's next launches use: set by bc_flux_num_dnc5_iso_2d(_d) / bcd_wall_iso around their
// calls (the scheme arguments of every other entry point are the reference's and have no place for twall).
struct WallIso {
  int on;
  double twall;
};
inline WallIso& current_wall_iso() {   // host function (declared in both compilation passes, called from host code only)
  static thread_local WallIso w{0, 0.0};
  return w;
}

BC_HD SchemeConsts make_consts(double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref,
                               double tref, double s_suth, double k2, double k4) {
  SchemeConsts c;
  c.gam = gam;
  c.rgaz = rgaz;
  c.cpprandtl = cp / prandtl;
  c.cvm1 = 1.0 / cv;
  c.betas = muref * (tref + cs) / (::sqrt(tref) * tref);  // flux_num_dnc5.F90:120
  c.s_suth = s_suth;
  c.k2 = k2;
  c.k4 = k4;
  c.gam1 = gam - 1.0;
  c.twall = 0.0;
  c.wall_iso = 0;
#if !defined(__CUDA_ARCH__)
  c.twall = current_wall_iso().twall;
  c.wall_iso = current_wall_iso().on;
#endif
  return c;
}

template <int DIR, int S, int T>
struct Off {
  static constexpr int i = DIR == 0 ? S : T;
  static constexpr int j = DIR == 0 ? T : S;
};
#define AT(S, T) Off<DIR, (S), (T)>::i, Off<DIR, (S), (T)>::j

// ---------------------------------------------------------------------------------------------
// The scheme family (row f2 of SURVEY.md section 8): orders 3 / 5 / 7 / 9 share every fragment except the stencil tables.
//   c(k)   centred Euler flux, pair k = cells (k, -k-1)            euler_o4 / o6 / o8 / o10_{i,j}.F
//   d(k)   predictor (the (ORD+2)-point difference), alternating   predictor_5p / 7p / 9p / 11p_{i,j}.F
//   b(k)   cell-centred gradient of the sensor velocities          gradop_3p / 5p / 7p / 9p{i,j}.F
//   near(row, r)  off-centred Euler flux of the j-face `row` (2 .. GH) over the cell rows r = 1 .. NP next to a wall
//                 nearbndfluxes{3,5,7,9}demi_{5,7,9,11}p.F with coefnearbnd_{7,9,11}p.F (order 3: flux_num_dnc3.F90:171-172)
// Every value is formed by the same floating-point operations as the reference's coefficient statements
// (flux_num_dnc3.F90:105-124, flux_num_dnc5.F90:90-106, flux_num_dnc7.F90:94-126, flux_num_dnc9.F90:100-129).
// ---------------------------------------------------------------------------------------------
template <int ORD>
struct SchemeOrd;

template <>
struct SchemeOrd<3> {
  static constexpr int GH = 2, NB = 1, NP = 2;
  static constexpr bool VISC_O4 = false;   // flux_visqueux_o2 on every face (flux_num_dnc3.F90:160)
  static BC_HD constexpr double c(int k) { return k == 0 ? 7.0 * (1.0 / 12.0) : -(1.0 / 12.0); }
  static BC_HD constexpr double d(int k) { return k == 0 ? 3.0 * (1.0 / 12.0) : (1.0 / 12.0); }
  static BC_HD constexpr double b(int) { return 0.5; }
  static BC_HD constexpr double near(int, int) { return 0.5; }
};

template <>
struct SchemeOrd<5> {
  static constexpr int GH = 3, NB = 2, NP = 5;
  static constexpr bool VISC_O4 = true;
  static BC_HD constexpr double c(int k) { return k == 0 ? 37.0 * (1.0 / 60.0) : k == 1 ? -8.0 * (1.0 / 60.0) : (1.0 / 60.0); }
  static BC_HD constexpr double d(int k) { return k == 0 ? 10.0 * (1.0 / 60.0) : k == 1 ? 5.0 * (1.0 / 60.0) : (1.0 / 60.0); }
  static BC_HD constexpr double b(int k) { return k == 0 ? 8.0 * (1.0 / 12.0) : -(1.0 / 12.0); }
  static BC_HD constexpr double near(int row, int r) {
    constexpr double denom = 1.0 / 60.0;
    constexpr double t3[5] = {-3.0, 27.0, 47.0, -13.0, 2.0}, t2[5] = {12.0, 77.0, -43.0, 17.0, -3.0};
    return (row == 3 ? t3[r - 1] : t2[r - 1]) * denom;
  }
};

template <>
struct SchemeOrd<7> {
  static constexpr int GH = 4, NB = 3, NP = 9;
  static constexpr bool VISC_O4 = true;
  static BC_HD constexpr double c(int k) {
    constexpr double denom = 1.0 / 840.0;
    return k == 0 ? 533.0 * denom : k == 1 ? -139.0 * denom : k == 2 ? 29.0 * denom : -3.0 * denom;
  }
  static BC_HD constexpr double d(int k) {
    constexpr double denom = 1.0 / 280.0;
    return k == 0 ? 35.0 * denom : k == 1 ? 21.0 * denom : k == 2 ? 7.0 * denom : 1.0 * denom;
  }
  static BC_HD constexpr double b(int k) {
    constexpr double denom = 1.0 / 60.0;
    return k == 0 ? 45.0 * denom : k == 1 ? -9.0 * denom : denom;
  }
  static BC_HD constexpr double near(int row, int r) {   // coefnearbnd_9p.F
    constexpr double denom = 1.0 / 840.0;
    constexpr double t4[9] = {2.0, -31.0, 281.0, 911.0, -517.0, 281.0, -111.0, 27.0, -3.0};
    constexpr double t3[9] = {-13.0, 209.0, 1079.0, -769.0, 533.0, -279.0, 99.0, -21.0, 2.0};
    constexpr double t2[9] = {92.0, 1547.0, -1861.0, 2171.0, -1917.0, 1191.0, -489.0, 119.0, -13.0};
    return (row == 4 ? t4[r - 1] : row == 3 ? t3[r - 1] : t2[r - 1]) * denom;
  }
};

template <>
struct SchemeOrd<9> {
  static constexpr int GH = 5, NB = 4, NP = 11;
  static constexpr bool VISC_O4 = true;
  static BC_HD constexpr double c(int k) {
    constexpr double denom = 1.0 / 2520.0;
    return k == 0 ? 1627.0 * denom : k == 1 ? -473.0 * denom : k == 2 ? 127.0 * denom : k == 3 ? -23.0 * denom : 2.0 * denom;
  }
  static BC_HD constexpr double d(int k) {
    constexpr double denom = 1.0 / 1260.0;
    return k == 0 ? 126.0 * denom : k == 1 ? 84.0 * denom : k == 2 ? 36.0 * denom : k == 3 ? 9.0 * denom : denom;
  }
  static BC_HD constexpr double b(int k) {
    constexpr double denom = 1.0 / 840.0;
    return k == 0 ? 672.0 * denom : k == 1 ? -168.0 * denom : k == 2 ? 32.0 * denom : -3.0 * denom;
  }
  // coefnearbnd_11p.F: the face coefficients are built at run time by successive subtraction of the difference weights a0 .. a10
  // from the centred flux of face 11/2; the same chain of subtractions here (compile time, IEEE double)
  static BC_HD constexpr double a9(int r) {
    constexpr double a[11] = {1.0 / 840.0, -1.0 / 63.0, 3.0 / 28.0, -4.0 / 7.0, -11.0 / 30.0, 6.0 / 5.0, -0.5, 4.0 / 21.0, -3.0 / 56.0,
                              1.0 / 105.0, -1.0 / 1260.0};
    return a[r];
  }
  static BC_HD constexpr double a7(int r) {
    constexpr double a[11] = {-1.0 / 360.0, 1.0 / 24.0, -3.0 / 8.0, -319.0 / 420.0, 1.75, -21.0 / 20.0, 7.0 / 12.0, -0.25, 0.075,
                              -1.0 / 72.0, 1.0 / 840.0};
    return a[r];
  }
  static BC_HD constexpr double a5(int r) {
    constexpr double a[11] = {1.0 / 90.0, -2.0 / 9.0, -341.0 / 280.0, 8.0 / 3.0, -7.0 / 3.0, 28.0 / 15.0, -7.0 / 6.0, 8.0 / 15.0, -1.0 / 6.0,
                              2.0 / 63.0, -1.0 / 360.0};
    return a[r];
  }
  static BC_HD constexpr double a3(int r) {
    constexpr double a[11] = {-0.1, -4609.0 / 2520.0, 4.5, -6.0, 7.0, -6.3, 21.0 * 0.2, -2.0, 9.0 / 14.0, -0.125, 1.0 / 90.0};
    return a[r];
  }
  static BC_HD constexpr double c9(int r) { return r == 10 ? -a9(10) : c(r <= 4 ? 4 - r : r - 5) - a9(r); }
  static BC_HD constexpr double near(int row, int r) {
    const int q = r - 1;
    double v = c9(q);
    if (row <= 4) v = v - a7(q);
    if (row <= 3) v = v - a5(q);
    if (row <= 2) v = v - a3(q);
    return v;
  }
};

// left-to-right sums over compile-time index packs (the reference's statements add their terms in source order)
template <int... K>
using ISeq = std::integer_sequence<int, K...>;
template <int N>
using MakeISeq = std::make_integer_sequence<int, N>;
template <int ORD, int K> struct EulerC { static constexpr double v = SchemeOrd<ORD>::c(K); };
template <int ORD, int K> struct GradC { static constexpr double v = SchemeOrd<ORD>::b(K); };
template <int ORD, int ROW, int R> struct NearC { static constexpr double v = SchemeOrd<ORD>::near(ROW, R); };
// predictor weight of the cell at offset O from the face cell: ... + d2 w(-2) - d1 w(-1) + d1 w(0) - d2 w(1) + ...
template <int ORD, int O> struct PredC {
  static constexpr int k = O < 0 ? -O : O + 1;
  static constexpr double sgn = ((k & 1) != 0) == (O >= 0) ? 1.0 : -1.0;
  static constexpr double v = sgn * SchemeOrd<ORD>::d(k - 1);
};

// ---------------------------------------------------------------------------------------------
// cell-local primitives (phys/Primitives.F:2-34, phys/viscosity.F:1)
// ---------------------------------------------------------------------------------------------
template <class D>
struct CellPrims {
  Var<D> u, v, w, t, p, mu, h;
};

template <class D>
BC_HD CellPrims<D> cell_prims(const Var<D> (&q)[5], const SchemeConsts& c) {
  CellPrims<D> r;
  const Var<D> ro = q[0];
  const Var<D> rom1 = 1.0 / ro;
  r.u = q[1] * rom1;
  r.v = q[2] * rom1;
  r.w = q[3] * rom1;
  const Var<D> ec = 0.5 * (r.u * r.u + r.v * r.v + r.w * r.w);
  const Var<D> eloc = (q[4] - ec * ro) * rom1;
  r.t = eloc * c.cvm1;
  r.p = c.gam1 * ro * eloc;
  r.h = (q[4] + r.p) * rom1;
  r.mu = c.betas / (r.t + c.s_suth) * sqrt(r.t) * r.t;
  return r;
}

// ---------------------------------------------------------------------------------------------
// 5-point gradients of (velx, vely) at a cell (gradop_5pi.F, gradop_5pj.F, gradient.F, dxdy.F).
// Offsets are relative to the accessor's base cell plus (CI,CJ).
// ---------------------------------------------------------------------------------------------
template <class D>
struct Grad4 {
  Var<D> u0, u1, v0, v1;  // gradu(.,1), gradu(.,2), gradv(.,1), gradv(.,2)
};

// centred differences of the sensor velocities over 2 NB + 1 points (gradop_{3,5,7,9}p{i,j}.F), terms in source order
template <int ORD, int CI, int CJ, class A, int... K>
BC_HD auto grad_diff_ui(const A& a, ISeq<K...>) { return (... + (GradC<ORD, K>::v * (a.template U<CI + K + 1, CJ>() - a.template U<CI - K - 1, CJ>()))); }
template <int ORD, int CI, int CJ, class A, int... K>
BC_HD auto grad_diff_vi(const A& a, ISeq<K...>) { return (... + (GradC<ORD, K>::v * (a.template V<CI + K + 1, CJ>() - a.template V<CI - K - 1, CJ>()))); }
template <int ORD, int CI, int CJ, class A, int... K>
BC_HD auto grad_diff_uj(const A& a, ISeq<K...>) { return (... + (GradC<ORD, K>::v * (a.template U<CI, CJ + K + 1>() - a.template U<CI, CJ - K - 1>()))); }
template <int ORD, int CI, int CJ, class A, int... K>
BC_HD auto grad_diff_vj(const A& a, ISeq<K...>) { return (... + (GradC<ORD, K>::v * (a.template V<CI, CJ + K + 1>() - a.template V<CI, CJ - K - 1>()))); }

template <int CI, int CJ, int ORD = 5, class A>
BC_HD auto cell_gradients(const A& a) {
  using NBS = MakeISeq<SchemeOrd<ORD>::NB>;
  auto gui = grad_diff_ui<ORD, CI, CJ>(a, NBS{});
  auto gvi = grad_diff_vi<ORD, CI, CJ>(a, NBS{});
  auto guj = grad_diff_uj<ORD, CI, CJ>(a, NBS{});
  auto gvj = grad_diff_vj<ORD, CI, CJ>(a, NBS{});
  const double volm1 = 1.0 / a.template VOL<CI, CJ>();
  const double dxm1 = 0.5 * (a.template NX<CI, CJ>(0) + a.template NX<CI + 1, CJ>(0)) * volm1;
  const double dxm2 = 0.5 * (a.template NX<CI, CJ>(1) + a.template NX<CI, CJ + 1>(1)) * volm1;
  const double dym1 = 0.5 * (a.template NY<CI, CJ>(0) + a.template NY<CI + 1, CJ>(0)) * volm1;
  const double dym2 = 0.5 * (a.template NY<CI, CJ>(1) + a.template NY<CI, CJ + 1>(1)) * volm1;
  auto gu0 = dxm1 * gui + dxm2 * guj;
  auto gv0 = dxm1 * gvi + dxm2 * gvj;
  auto gu1 = dym1 * gui + dym2 * guj;
  auto gv1 = dym1 * gvi + dym2 * gvj;
  using D = decltype(gu0.d);
  return Grad4<D>{gu0, gu1, gv0, gv1};
}

// ---------------------------------------------------------------------------------------------
// inviscid fluxes f (x) and g (y) of one cell, component e (Primitives.F:22-33)
// ---------------------------------------------------------------------------------------------
template <int OI, int OJ, class A>
BC_HD auto flux_f(const A& a, int e) {
  auto m = a.template W<OI, OJ>(1);
  using R = decltype(m * a.template U<OI, OJ>() + a.template P<OI, OJ>());
  switch (e) {
    case 0: return R(m * 1.0);
    case 1: return R(m * a.template U<OI, OJ>() + a.template P<OI, OJ>());
    case 2: return R(m * a.template V<OI, OJ>());
    case 3: return R(m * a.template Wz<OI, OJ>());
    default: return R(m * a.template H<OI, OJ>());
  }
}
template <int OI, int OJ, class A>
BC_HD auto flux_g(const A& a, int e) {
  auto m = a.template W<OI, OJ>(2);
  using R = decltype(m * a.template V<OI, OJ>() + a.template P<OI, OJ>());
  switch (e) {
    case 0: return R(m * 1.0);
    case 1: return R(m * a.template U<OI, OJ>());
    case 2: return R(m * a.template V<OI, OJ>() + a.template P<OI, OJ>());
    case 3: return R(m * a.template Wz<OI, OJ>());
    default: return R(m * a.template H<OI, OJ>());
  }
}


// ---------------------------------------------------------------------------------------------
// Pieces of a face flux that the block-Jacobian assembly (facejac.cuh) also needs on their own.
// face_flux() below is written in terms of them, so both paths evaluate the same expressions.
// ---------------------------------------------------------------------------------------------

// normals of the face-centred dual cell used by the viscous Green-Gauss gradients
// (flux_visqueux_o4_{i,j}.F / flux_visqueux_o2_{i,j}.F): A = along the face normal direction, C = across
struct DualNormals {
  double volm1;
  double nAp_x, nAm_x, nCp_x, nCm_x, nAp_y, nAm_y, nCp_y, nCm_y;
};
template <int DIR, class A>
BC_HD DualNormals dual_normals(const A& a) {
  constexpr int kA = DIR, kC = 1 - DIR;
  DualNormals n;
  n.volm1 = a.template VOLF<0, 0>(DIR);
  n.nAp_x = 0.5 * (a.template NX<AT(1, 0)>(kA) + a.template NX<0, 0>(kA));
  n.nAm_x = -0.5 * (a.template NX<AT(-1, 0)>(kA) + a.template NX<0, 0>(kA));
  n.nCp_x = 0.5 * (a.template NX<AT(-1, 1)>(kC) + a.template NX<AT(0, 1)>(kC));
  n.nCm_x = -0.5 * (a.template NX<AT(-1, 0)>(kC) + a.template NX<0, 0>(kC));
  n.nAp_y = 0.5 * (a.template NY<AT(1, 0)>(kA) + a.template NY<0, 0>(kA));
  n.nAm_y = -0.5 * (a.template NY<AT(-1, 0)>(kA) + a.template NY<0, 0>(kA));
  n.nCp_y = 0.5 * (a.template NY<AT(-1, 1)>(kC) + a.template NY<AT(0, 1)>(kC));
  n.nCm_y = -0.5 * (a.template NY<AT(-1, 0)>(kC) + a.template NY<0, 0>(kC));
  return n;
}

// face values of the velocity / temperature gradients and of u, v, w, mu (compact o4, or o2 near a wall)
template <class A1, class A2, class A3, class A4, class A5, class A6, class A7, class A8, class A9, class A10, class A11, class A12>
struct ViscScalarsT {
  A1 ux; A2 uy; A3 vx; A4 vy; A5 wx; A6 wy; A7 tx; A8 ty; A9 uu; A10 vv; A11 ww; A12 mmu;
};
template <class... Ts>
BC_HD ViscScalarsT<Ts...> make_visc_scalars(Ts... v) {
  return ViscScalarsT<Ts...>{v...};
}
template <int DIR, bool VISC_O2, class A>
BC_HD auto visc_scalars(const A& a, const DualNormals& n) {
  const double volm1 = n.volm1;
  // Green-Gauss gradient of one scalar from its four dual-cell side values; the reference sums
  // N,S,O,E which is (A+,A-,C+,C-) for i-faces and (C+,C-,A+,A-) for j-faces.
  auto gg = [&](auto vAp, auto vAm, auto vCp, auto vCm, double nAp, double nAm, double nCp, double nCm) {
    if constexpr (DIR == 0)
      return (vAp * nAp + vAm * nAm + vCp * nCp + vCm * nCm) * volm1;
    else
      return (vCp * nCp + vCm * nCm + vAp * nAp + vAm * nAm) * volm1;
  };
  constexpr double TWENTYFOURTH = 1.0 / 24.0;
  constexpr double TWELFTH = 0.25 / 3.0;
  constexpr double ccross = TWELFTH * 0.0625;
  (void)TWENTYFOURTH;
  (void)ccross;
#define BC_O4_ROW(Q, T_) (-a.template Q<AT(-2, T_)>() + 9.0 * a.template Q<AT(-1, T_)>() + 9.0 * a.template Q<AT(0, T_)>() - a.template Q<AT(1, T_)>())
#define BC_O4_SIDES(Q)                                                                                              \
  auto Q##_Ap = TWENTYFOURTH * (-a.template Q<AT(1, 0)>() + 26.0 * a.template Q<AT(0, 0)>() - a.template Q<AT(-1, 0)>());  \
  auto Q##_Am = TWENTYFOURTH * (-a.template Q<AT(0, 0)>() + 26.0 * a.template Q<AT(-1, 0)>() - a.template Q<AT(-2, 0)>()); \
  auto Q##_Cm = ccross * (-BC_O4_ROW(Q, -2) + 7.0 * BC_O4_ROW(Q, -1) + 7.0 * BC_O4_ROW(Q, 0) - BC_O4_ROW(Q, 1));       \
  auto Q##_Cp = ccross * (-BC_O4_ROW(Q, -1) + 7.0 * BC_O4_ROW(Q, 0) + 7.0 * BC_O4_ROW(Q, 1) - BC_O4_ROW(Q, 2));
#define BC_O2_SIDES(Q)                                                                                                          \
  auto Q##_Ap = a.template Q<AT(0, 0)>();                                                                                       \
  auto Q##_Am = a.template Q<AT(-1, 0)>();                                                                                      \
  auto Q##_Cm = 0.25 * (a.template Q<AT(0, 0)>() + a.template Q<AT(0, -1)>() + a.template Q<AT(-1, 0)>() + a.template Q<AT(-1, -1)>()); \
  auto Q##_Cp = 0.25 * (a.template Q<AT(0, 0)>() + a.template Q<AT(0, 1)>() + a.template Q<AT(-1, 0)>() + a.template Q<AT(-1, 1)>());
#define BC_GRADS(Q, GX, GY)                                                \
  auto GX = gg(Q##_Ap, Q##_Am, Q##_Cp, Q##_Cm, n.nAp_x, n.nAm_x, n.nCp_x, n.nCm_x); \
  auto GY = gg(Q##_Ap, Q##_Am, Q##_Cp, Q##_Cm, n.nAp_y, n.nAm_y, n.nCp_y, n.nCm_y);
#define BC_VS_RETURN return make_visc_scalars(ux, uy, vx, vy, wx, wy, tx, ty, uu, vv, ww, mmu);
  if constexpr (!VISC_O2) {
    BC_O4_SIDES(U) BC_GRADS(U, ux, uy)
    BC_O4_SIDES(V) BC_GRADS(V, vx, vy)
    BC_O4_SIDES(Wz) BC_GRADS(Wz, wx, wy)
    BC_O4_SIDES(T) BC_GRADS(T, tx, ty)
    auto uu = 0.0625 * BC_O4_ROW(U, 0);
    auto vv = 0.0625 * BC_O4_ROW(V, 0);
    auto ww = 0.0625 * BC_O4_ROW(Wz, 0);
    auto mmu = 0.0625 * BC_O4_ROW(Mu, 0);
    BC_VS_RETURN
  } else {
    BC_O2_SIDES(U) BC_GRADS(U, ux, uy)
    BC_O2_SIDES(V) BC_GRADS(V, vx, vy)
    BC_O2_SIDES(Wz) BC_GRADS(Wz, wx, wy)
    BC_O2_SIDES(T) BC_GRADS(T, tx, ty)
    auto uu = 0.5 * (a.template U<0, 0>() + a.template U<AT(-1, 0)>());
    auto vv = 0.5 * (a.template V<0, 0>() + a.template V<AT(-1, 0)>());
    auto ww = 0.5 * (a.template Wz<0, 0>() + a.template Wz<AT(-1, 0)>());
    auto mmu = 0.5 * (a.template Mu<0, 0>() + a.template Mu<AT(-1, 0)>());
    BC_VS_RETURN
  }
#undef BC_O4_ROW
#undef BC_O4_SIDES
#undef BC_O2_SIDES
#undef BC_GRADS
#undef BC_VS_RETURN
}

// viscous stresses and heat flux from the face scalars: f[1..4], g[1..4] (mass components unused)
template <class VS>
BC_HD auto visc_stress(const VS& s, const SchemeConsts& c) {
  constexpr double TWOTHIRD = 2.0 / 3.0;
  auto lambda = s.mmu * c.cpprandtl;
  auto fvrou = TWOTHIRD * s.mmu * (2.0 * s.ux - s.vy);
  auto fvrov = s.mmu * (s.uy + s.vx);
  auto fvrow = s.mmu * s.wx;
  auto fvroe = lambda * s.tx + s.uu * fvrou + s.vv * fvrov + s.ww * fvrow;
  auto gvrou = s.mmu * (s.uy + s.vx);
  auto gvrov = TWOTHIRD * s.mmu * (-s.ux + 2.0 * s.vy);
  auto gvrow = s.mmu * s.wy;
  auto gvroe = lambda * s.ty + s.uu * gvrou + s.vv * gvrov + s.ww * gvrow;
  using D = decltype((fvroe + gvroe).d);
  struct R {
    Var<D> f[5], g[5];
  };
  R r;
  r.f[1] = promote<D>(fvrou); r.f[2] = promote<D>(fvrov); r.f[3] = promote<D>(fvrow); r.f[4] = promote<D>(fvroe);
  r.g[1] = promote<D>(gvrou); r.g[2] = promote<D>(gvrov); r.g[3] = promote<D>(gvrow); r.g[4] = promote<D>(gvroe);
  return r;
}

// Roe-averaged spectral radius of a face (spectralradius_{i,j}.F); r = right (face cell), l = left.
// The cell velocities are taken from the primitives (w * (1/rho); the reference divides w by rho again here: 1 ulp).
template <class R0, class R1, class R2, class TR, class L0, class L1, class L2, class TL>
BC_HD auto spectral_radius(Var<R0> rhomr, Var<R1> ur, Var<R2> vr, Var<TR> tr, Var<L0> rhoml, Var<L1> ul, Var<L2> vl, Var<TL> tl,
                           double nxf, double nyf, const SchemeConsts& c) {
  auto c2r = c.gam * c.rgaz * tr;
  auto c2l = c.gam * c.rgaz * tl;
  auto r = sqrt(rhomr / rhoml);
  auto rr = 1.0 / (1.0 + r);
  auto omrr = 1.0 - rr;
  auto u = ul * rr + ur * omrr;
  auto v = vl * rr + vr * omrr;
  auto c2x = c2l * rr + c2r * omrr;
  const double nx2 = nxf * nxf + nyf * nyf;
  auto ab = fabs(nxf * u + nyf * v);
  auto sq = sqrt(c2x * nx2);
  return ab + sq;
}

// per-cell part of the sensor: dilatation and the Ducros ratio from the velocity gradients (ducrosfordnc_{i,j}.F)
template <class DV, class DU>
struct CellSens {
  Var<DV> divu;
  Var<DU> ducros;
};
template <class A, class B>
BC_HD auto sens_from_divu_vort(Var<A> divu, Var<B> vort) {
  auto divu2 = divu * divu;
  auto vort2 = vort * vort;
  auto ducros = divu2 / (divu2 + vort2 + 1e-15);
  return CellSens<A, decltype(ducros.d)>{divu, ducros};
}
template <class G>
BC_HD auto sens_from_grad(const G& gr) {
  return sens_from_divu_vort(gr.u0 + gr.v1, gr.v0 - gr.u1);
}

// Jameson pressure sensor x Ducros x dilatation switch (ducrosfordnc_{i,j}.F): the factor `coef` of eps2.
// s0 / s1: CellSens of the face cell and of its along-neighbour -1.  The dilatation switch
// dxm = (1 - tanh(x)) / 2 is decreasing in x, so max(dxm0, dxm1) is evaluated as ONE tanh of the smaller argument
// (same branch rule as the reference's max: the second operand is taken iff the first is smaller).
template <class P2, class P1, class P0, class PP, class S0, class S1, class C0, class C1>
BC_HD auto sensor_coef(Var<P2> p_m2, Var<P1> p_m1, Var<P0> p_0, Var<PP> p_p1, const S0& s0, const S1& s1, double vol0, double vol1,
                       Var<C0> c2r, Var<C1> c2l, double nx2) {
  auto k_sensor1 = fabs(p_m1 - 2.0 * p_0 + p_p1) / fabs(p_m1 + 2.0 * p_0 + p_p1);
  auto k_sensor2 = fabs(p_m2 - 2.0 * p_m1 + p_0) / fabs(p_m2 + 2.0 * p_m1 + p_0);
  auto x0 = 2.5 + 10.0 * vol0 / (sqrt(c2r * nx2) + 1e-15) * s0.divu;
  auto x1 = 2.5 + 10.0 * vol1 / (sqrt(c2l * nx2) + 1e-15) * s1.divu;
  const bool take1 = x0.v > x1.v;   // dxm(x0) < dxm(x1)
  const Var<decltype(t_sel(take1, x1.d, x0.d))> xs{take1 ? x1.v : x0.v, t_sel(take1, x1.d, x0.d)};
  auto dxm = 0.5 * (1.0 - tanh(xs));
  return fmax(k_sensor1, k_sensor2) * fmax(s0.ducros, s1.ducros) * dxm;
}

template <int N>
template <int OI, int OJ>
BC_HD auto GlobalAcc<N>::SENS() const {
  return sens_from_grad(GR<OI, OJ>());
}

enum FaceMode { FACE_MAIN = 0, FACE_NEAR5 = 1, FACE_NEAR3 = 2, FACE_WALL = 3 };

// ---------------------------------------------------------------------------------------------
// One face flux hn(1:5) in direction DIR.
//   VISC_O2 : 2nd-order viscous gradients (rows j <= 2 of the wall scheme) instead of the compact o4
//   MODE    : FACE_MAIN  centred 6-point Euler flux
//             FACE_NEAR5 / FACE_NEAR3  off-centred wall-adjacent Euler flux (j-faces at j = 3 / 2)
//             FACE_WALL  wall flux (j-face at j = 1)             -- DIR must be 1 for the last three
// Result type RD is the accessor's widest tangent type.
// ---------------------------------------------------------------------------------------------
// centred Euler flux of component e over the 2 GH cells along the face normal (euler_o{4,6,8,10}_{i,j}.F)
template <int DIR, int ORD, class A, int... K>
BC_HD auto euler_centred(const A& a, int e, double nxf, double nyf, ISeq<K...>) {
  return (... + (EulerC<ORD, K>::v * (flux_f<AT(K, 0)>(a, e) + flux_f<AT(-K - 1, 0)>(a, e)))) * nxf +
         (... + (EulerC<ORD, K>::v * (flux_g<AT(K, 0)>(a, e) + flux_g<AT(-K - 1, 0)>(a, e)))) * nyf;
}
// off-centred Euler flux of the j-face ROW next to a wall: cell rows 1 .. NP = offsets 1 - ROW .. NP - ROW from the face cell
template <int DIR, int ORD, int ROW, class A, int... R>
BC_HD auto euler_near_wall(const A& a, int e, double nxf, double nyf, ISeq<R...>) {
  return (... + (NearC<ORD, ROW, R + 1>::v * flux_f<AT(R + 1 - ROW, 0)>(a, e))) * nxf +
         (... + (NearC<ORD, ROW, R + 1>::v * flux_g<AT(R + 1 - ROW, 0)>(a, e))) * nyf;
}
// predictor_{5,7,9,11}p_{i,j}.F over the offsets -GH .. GH - 1
template <int DIR, int ORD, class A, int... Q>
BC_HD auto predictor_diff(const A& a, int e, ISeq<Q...>) {
  return (... + (PredC<ORD, Q - SchemeOrd<ORD>::GH>::v * a.template W<AT(Q - SchemeOrd<ORD>::GH, 0)>(e)));
}

// j-face rows 2 .. GH of a wall block for orders other than 5: MODE = FACE_NEAR_ROW + row
constexpr int FACE_NEAR_ROW = 10;

template <int DIR, bool VISC_O2, int MODE, int ORD = 5, class A, class RD>
BC_HD void face_flux(const A& a, const SchemeConsts& c, Var<RD> (&hn)[5]) {
  const double nxf = a.template NX<0, 0>(DIR);
  const double nyf = a.template NY<0, 0>(DIR);

  if constexpr (MODE == FACE_WALL) {
    // fluxwall.F:3-50  (ct0 = 9/8, ct1 = -1/8: flux_num_dnc5.F90:196-197)
    auto pw = 1.125 * a.template P<0, 0>() + (-0.125) * a.template P<AT(1, 0)>();
    auto mmu = a.template Mu<0, 0>();
    const double vf = a.template VOLF<0, 0>(DIR);
    auto ux = 2.0 * a.template U<0, 0>() * nxf * vf;
    auto vx = 2.0 * a.template V<0, 0>() * nxf * vf;
    auto wx = 2.0 * a.template Wz<0, 0>() * nxf * vf;
    auto uy = 2.0 * a.template U<0, 0>() * nyf * vf;
    auto vy = 2.0 * a.template V<0, 0>() * nyf * vf;
    auto wy = 2.0 * a.template Wz<0, 0>() * nyf * vf;
    constexpr double TWOTHIRD = 2.0 / 3.0;
    auto fvrou = TWOTHIRD * mmu * (2.0 * ux - vy);
    auto fvrov = mmu * (uy + vx);
    auto fvrow = mmu * wx;
    auto gvrou = mmu * (uy + vx);
    auto gvrov = TWOTHIRD * mmu * (-ux + 2.0 * vy);
    auto gvrow = mmu * wy;
    hn[0] = promote<RD>(cst(0.0));
    hn[1] = promote<RD>(pw * nxf - (fvrou * nxf + gvrou * nyf));
    hn[2] = promote<RD>(pw * nyf - (fvrov * nxf + gvrov * nyf));
    hn[3] = promote<RD>(-(fvrow * nxf + gvrow * nyf));
    if (c.wall_iso) {
      // fluxwall_iso.F:29-53.  Kept as the reference computes it: no factor 2 in the temperature gradient, and `lambda` is the
      // value the i-face viscous fragment of the same cell left behind (flux_visqueux_o2_i.F:61: the AVERAGE of mu over the
      // cells (i-1, 1) and (i, 1), not mu(i, 1)).
      auto lambda = 0.5 * (a.template Mu<0, 0>() + a.template Mu<AT(0, -1)>()) * c.cpprandtl;
      auto tx = (a.template T<0, 0>() - c.twall) * nxf * vf;
      auto ty = (a.template T<0, 0>() - c.twall) * nyf * vf;
      hn[4] = promote<RD>(-(lambda * tx * nxf + lambda * ty * nyf));
    } else {
      hn[4] = promote<RD>(cst(0.0));
    }
    return;
  } else {
    // ---- viscous face gradients, stresses --------------------------------------------------------
    const DualNormals dn = dual_normals<DIR>(a);
    const auto vsc = visc_scalars<DIR, VISC_O2>(a, dn);
    const auto vs = visc_stress(vsc, c);

    // ---- scalar dissipation: Roe spectral radius (spectralradius_{i,j}.F) -------------------------
    auto rspec = spectral_radius(a.template W<0, 0>(0), a.template U<0, 0>(), a.template V<0, 0>(), a.template T<0, 0>(),
                                 a.template W<AT(-1, 0)>(0), a.template U<AT(-1, 0)>(), a.template V<AT(-1, 0)>(),
                                 a.template T<AT(-1, 0)>(), nxf, nyf, c);
    const double nx2 = nxf * nxf + nyf * nyf;

    // ---- Jameson / Ducros / dilatation sensor (ducrosfordnc_{i,j}.F) ------------------------------
    auto c2r = c.gam * c.rgaz * a.template T<0, 0>();
    auto c2l = c.gam * c.rgaz * a.template T<AT(-1, 0)>();
    auto coef = sensor_coef(a.template P<AT(-2, 0)>(), a.template P<AT(-1, 0)>(), a.template P<AT(0, 0)>(), a.template P<AT(1, 0)>(),
                            a.template SENS<0, 0>(), a.template SENS<AT(-1, 0)>(), a.template VOL<0, 0>(), a.template VOL<AT(-1, 0)>(),
                            c2r, c2l, nx2);
    auto eps2 = c.k2 * coef;
    auto eps4 = fmax(0.0, c.k4 - eps2 * 12.0);

    // ---- assembly (fluxnumassembly_{i,j}.F) -------------------------------------------------------
    const double sn = ::sqrt(nxf * nxf + nyf * nyf);
    const double invsn = 1.0 / sn;
    const double nxloc = nxf * invsn;
    const double nyloc = nyf * invsn;

#pragma unroll
    for (int e = 0; e < 5; ++e) {
      // Euler flux
      auto euler = [&]() {
        if constexpr (MODE == FACE_MAIN) {
          return euler_centred<DIR, ORD>(a, e, nxf, nyf, MakeISeq<SchemeOrd<ORD>::GH>{});
        } else {
          // FACE_NEAR5 / FACE_NEAR3: rows 3 / 2 of the order-5 scheme (nearbndfluxes5demi_7p.F, nearbndfluxes3demi_7p.F)
          constexpr int row = MODE == FACE_NEAR5 ? 3 : MODE == FACE_NEAR3 ? 2 : MODE - FACE_NEAR_ROW;
          return euler_near_wall<DIR, ORD, row>(a, e, nxf, nyf, MakeISeq<SchemeOrd<ORD>::NP>{});
        }
      };
      auto fx = euler();
      // predictor_7p_{i,j}.F
      auto pred = predictor_diff<DIR, ORD>(a, e, MakeISeq<2 * SchemeOrd<ORD>::GH>{});
      auto diff = 0.5 * (a.template W<AT(0, 0)>(e) - a.template W<AT(-1, 0)>(e));
      auto diss = rspec * (eps2 * diff + eps4 * pred);
      if (e == 0)
        hn[e] = promote<RD>(fx - diss);
      else
        hn[e] = promote<RD>(fx - diss - (vs.f[e] * nxloc + vs.g[e] * nyloc) * sn);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Which face formula the reference's driver applies on row j of a block (flux_num_dnc{3,5,7,9}.F90: main loop from j = GH + 1,
// off-centred Euler j-fluxes on rows GH .. 2, wall flux on row 1; 2nd-order viscous gradients on rows <= 2 and everywhere for
// order 3).  `wall` false = the _nowall variants (main loop from j = 1).
// ---------------------------------------------------------------------------------------------
template <int DIR, int ORD, int ROW, class A, class RD>
BC_HD void face_near_rows(const A& a, const SchemeConsts& c, int j, Var<RD> (&hn)[5]) {
  if constexpr (ROW >= 3) {
    if (j == ROW) {
      face_flux<1, !SchemeOrd<ORD>::VISC_O4, FACE_NEAR_ROW + ROW, ORD>(a, c, hn);
      return;
    }
    face_near_rows<DIR, ORD, ROW - 1>(a, c, j, hn);
  } else {
    face_flux<1, true, FACE_NEAR_ROW + 2, ORD>(a, c, hn);   // row 2
  }
}

template <int DIR, int ORD, class A, class RD>
BC_HD void face_by_row(const A& a, const SchemeConsts& c, bool wall, int j, Var<RD> (&hn)[5]) {
  constexpr bool o2 = !SchemeOrd<ORD>::VISC_O4;
  if constexpr (DIR == 0) {
    if (o2 || (wall && j <= 2))
      face_flux<0, true, FACE_MAIN, ORD>(a, c, hn);
    else
      face_flux<0, false, FACE_MAIN, ORD>(a, c, hn);
  } else {
    if (wall && j == 1)
      face_flux<1, true, FACE_WALL, ORD>(a, c, hn);
    else if (wall && j <= SchemeOrd<ORD>::GH)
      face_near_rows<1, ORD, SchemeOrd<ORD>::GH>(a, c, j, hn);
    else
      face_flux<1, o2, FACE_MAIN, ORD>(a, c, hn);
  }
}

}  // namespace bcast
