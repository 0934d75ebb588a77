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

template <int CI, int CJ, class A>
BC_HD auto cell_gradients(const A& a) {
  constexpr double b1 = 8.0 * (1.0 / 12.0);
  constexpr double b2 = -(1.0 / 12.0);
  auto gui = b1 * (a.template U<CI + 1, CJ>() - a.template U<CI - 1, CJ>()) + b2 * (a.template U<CI + 2, CJ>() - a.template U<CI - 2, CJ>());
  auto gvi = b1 * (a.template V<CI + 1, CJ>() - a.template V<CI - 1, CJ>()) + b2 * (a.template V<CI + 2, CJ>() - a.template V<CI - 2, CJ>());
  auto guj = b1 * (a.template U<CI, CJ + 1>() - a.template U<CI, CJ - 1>()) + b2 * (a.template U<CI, CJ + 2>() - a.template U<CI, CJ - 2>());
  auto gvj = b1 * (a.template V<CI, CJ + 1>() - a.template V<CI, CJ - 1>()) + b2 * (a.template V<CI, CJ + 2>() - a.template V<CI, CJ - 2>());
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
template <int DIR, bool VISC_O2, int MODE, class A, class RD>
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
    constexpr double denom = 1.0 / 60.0;
    constexpr double c1 = 37.0 * denom, c2 = -8.0 * denom, c3 = denom;
    constexpr double d1 = 10.0 * denom, d2 = 5.0 * denom, d3 = denom;

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
          // euler_o6_{i,j}.F
          return (c1 * (flux_f<AT(0, 0)>(a, e) + flux_f<AT(-1, 0)>(a, e)) + c2 * (flux_f<AT(1, 0)>(a, e) + flux_f<AT(-2, 0)>(a, e)) +
                  c3 * (flux_f<AT(2, 0)>(a, e) + flux_f<AT(-3, 0)>(a, e))) * nxf +
                 (c1 * (flux_g<AT(0, 0)>(a, e) + flux_g<AT(-1, 0)>(a, e)) + c2 * (flux_g<AT(1, 0)>(a, e) + flux_g<AT(-2, 0)>(a, e)) +
                  c3 * (flux_g<AT(2, 0)>(a, e) + flux_g<AT(-3, 0)>(a, e))) * nyf;
        } else {
          // nearbndfluxes5demi_7p.F (face j = 3: rows 1..5 = offsets -2..2) / nearbndfluxes3demi_7p.F
          // (face j = 2: rows 1..5 = offsets -1..3); coefficients coefnearbnd_7p.F
          constexpr bool five = (MODE == FACE_NEAR5);
          constexpr double k0 = (five ? -3.0 : 12.0) * denom, k1 = (five ? 27.0 : 77.0) * denom, k2_ = (five ? 47.0 : -43.0) * denom,
                           k3 = (five ? -13.0 : 17.0) * denom, k4_ = (five ? 2.0 : -3.0) * denom;
          constexpr int o = five ? -2 : -1;
          return (k0 * flux_f<AT(o, 0)>(a, e) + k1 * flux_f<AT(o + 1, 0)>(a, e) + k2_ * flux_f<AT(o + 2, 0)>(a, e) +
                  k3 * flux_f<AT(o + 3, 0)>(a, e) + k4_ * flux_f<AT(o + 4, 0)>(a, e)) * nxf +
                 (k0 * flux_g<AT(o, 0)>(a, e) + k1 * flux_g<AT(o + 1, 0)>(a, e) + k2_ * flux_g<AT(o + 2, 0)>(a, e) +
                  k3 * flux_g<AT(o + 3, 0)>(a, e) + k4_ * flux_g<AT(o + 4, 0)>(a, e)) * nyf;
        }
      };
      auto fx = euler();
      // predictor_7p_{i,j}.F
      auto pred = -d3 * a.template W<AT(-3, 0)>(e) + d2 * a.template W<AT(-2, 0)>(e) - d1 * a.template W<AT(-1, 0)>(e) +
                  d1 * a.template W<AT(0, 0)>(e) - d2 * a.template W<AT(1, 0)>(e) + d3 * a.template W<AT(2, 0)>(e);
      auto diff = 0.5 * (a.template W<AT(0, 0)>(e) - a.template W<AT(-1, 0)>(e));
      auto diss = rspec * (eps2 * diff + eps4 * pred);
      if (e == 0)
        hn[e] = promote<RD>(fx - diss);
      else
        hn[e] = promote<RD>(fx - diss - (vs.f[e] * nxloc + vs.g[e] * nyloc) * sn);
    }
  }
}

}  // namespace bcast
