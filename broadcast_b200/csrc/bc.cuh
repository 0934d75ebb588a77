// Ghost-cell boundary fills of the BROADCAST hot path, one thread per boundary-line cell, written
// once for passive (D = Zero) and tangent (D = Tan<N>) arithmetic.
//
// Reference (restated): srcfv/borders/init_2d.F:1-44 (interface decoding),
//   bc_wall_viscous.F90:2-99, bc_no_reflexion.F90:8-152, bc_supandsubinlet.F90:2-171,
//   bc_extrapolate.F90:39-71, jn_match.F90:3-66, bc_wall_viscous_iso.F90:1-108, bc_symmetry.F90:1-79 (the last two: SURVEY.md 8(f3),
//   the sensitivity driver's walls / half-domain cards) and their Tapenade tangents in srcfv/tangent/.
// The tangent routines of the reference update BOTH w and wd ghosts, except the extrapolation whose
// tangent writes wd only (tangent/bc_extrapolateo2_d.f90:63-65).
#pragma once
#include "grid.cuh"

namespace bcast {

// decoded interface (init_2d.F): line of lmax cells starting at (imin,jmin), ghost direction (-i0,-j0)
struct BcLine {
  int imin, jmin, i0, j0, kdir, high, lmax;
};

inline bool decode_interface(const char* loc, const int interf[4] /* imin,jmin,imax,jmax */, BcLine& b) {
  b.imin = interf[0];
  b.jmin = interf[1];
  const int imax = interf[2], jmax = interf[3];
  b.i0 = 0;
  b.j0 = 0;
  b.high = 0;
  b.kdir = 0;
  b.lmax = 0;
  if (loc[0] == 'I' && loc[1] == 'l' && loc[2] == 'o') {
    b.kdir = 1; b.i0 = 1; b.lmax = jmax - b.jmin + 1;
  } else if (loc[0] == 'I' && loc[1] == 'h' && loc[2] == 'i') {
    b.kdir = 1; b.i0 = -1; b.lmax = jmax - b.jmin + 1; b.high = 1;
  } else if (loc[0] == 'J' && loc[1] == 'l' && loc[2] == 'o') {
    b.kdir = 2; b.j0 = 1; b.lmax = imax - b.imin + 1;
  } else if (loc[0] == 'J' && loc[1] == 'h' && loc[2] == 'i') {
    b.kdir = 2; b.j0 = -1; b.lmax = imax - b.imin + 1; b.high = 1;
  } else {
    return false;
  }
  return true;
}

// read/write view of the state (and its tangents) with run-time Fortran indices
template <int N>
struct StateRW {
  using DT = TanOf<N>;
  double* w;
  double* wd;  // [n][5 planes]
  GridDesc g;
  __device__ __forceinline__ Var<DT> get(int i, int j, int e) const {
    const long long k = g.cidx(i, j);
    Var<DT> r;
    r.v = w[e * g.sc + k];
    if constexpr (N > 0) {
#pragma unroll
      for (int q = 0; q < N; ++q) r.d.d[q] = wd[(long long)(q * 5 + e) * g.sc + k];
    }
    return r;
  }
  __device__ __forceinline__ void set(int i, int j, int e, const Var<DT>& x, bool primal = true) const {
    const long long k = g.cidx(i, j);
    if (primal) w[e * g.sc + k] = x.v;
    if constexpr (N > 0) {
#pragma unroll
      for (int q = 0; q < N; ++q) wd[(long long)(q * 5 + e) * g.sc + k] = x.d.d[q];
    }
  }
};

// ---------------------------------------------------------------------------------------------
// adiabatic viscous wall (bc_wall_viscous.F90:46-97)
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ void bc_wall_viscous_adia_line(const StateRW<N>& s, const BcLine& b, double gam, int l /*0-based*/) {
  using DT = TanOf<N>;
  using VT = Var<DT>;
  const int i = b.imin + l * b.j0 * b.j0;
  const int j = b.jmin + l * b.i0 * b.i0;
  const int i0 = b.i0, j0 = b.j0, gh = s.g.gh;
  const double gam1 = gam - 1.0, gami = 1.0 / gam;
  const double THIRD = 1.0 / 3.0;

  VT roe = s.get(i, j, 0);
  VT roem1 = 1.0 / roe;
  VT ue = s.get(i, j, 1) * roem1;
  VT ve = s.get(i, j, 2) * roem1;
  VT we = s.get(i, j, 3) * roem1;
  VT ve2 = ue * ue + ve * ve;
  VT pe = gam1 * (s.get(i, j, 4) - 0.5 * roe * ve2);

  VT roe1 = s.get(i + i0, j + j0, 0);
  VT roe1m1 = 1.0 / roe1;
  VT ue1 = s.get(i + i0, j + j0, 1) * roe1m1;
  VT ve1 = s.get(i + i0, j + j0, 2) * roe1m1;
  VT ve21 = ue1 * ue1 + ve1 * ve1;
  VT pe1 = gam1 * (s.get(i + i0, j + j0, 4) - 0.5 * roe1 * ve21);

  VT pw = 1.125 * pe + (-0.125) * pe1;
  VT pi = THIRD * (4.0 * pw - pe);
  VT roi = pow(pi / pe * pow(roe, gam), gami);
  VT ui = -ue, vi = -ve, wi = -we;
  VT roiei = pi / gam1 + 0.5 * roi * (ui * ui + vi * vi);

  for (int de = 1; de <= gh; ++de) {
    s.set(i - de * i0, j - de * j0, 0, roi);
    s.set(i - de * i0, j - de * j0, 1, roi * ui);
    s.set(i - de * i0, j - de * j0, 2, roi * vi);
    s.set(i - de * i0, j - de * j0, 3, roi * wi);
    s.set(i - de * i0, j - de * j0, 4, roiei);
    roi = s.get(i + de * i0, j + de * j0, 0);
    VT rm1 = 1.0 / roi;
    VT u1 = s.get(i + de * i0, j + de * j0, 1) * rm1;
    VT v1 = s.get(i + de * i0, j + de * j0, 2) * rm1;
    VT w1 = s.get(i + de * i0, j + de * j0, 3) * rm1;
    roiei = s.get(i + de * i0, j + de * j0, 4);
    ui = -u1;
    vi = -v1;
    wi = -w1;
  }
}

// ---------------------------------------------------------------------------------------------
// isothermal viscous wall (bc_wall_viscous_iso.F90:37-105; tangent/bc_wall_viscous_iso_d.f90): wall pressure as for the adiabatic
// wall, first ghost density from the wall temperature (roi = 2 pw / (rgaz twall) - roe), deeper ghost densities by linear
// extrapolation of the two previous layers, velocities mirrored, ghost internal energy pi / (gam - 1) for every layer
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ void bc_wall_viscous_iso_line(const StateRW<N>& s, const BcLine& b, double twall, double gam, double rgaz, int l /*0-based*/) {
  using DT = TanOf<N>;
  using VT = Var<DT>;
  const int i = b.imin + l * b.j0 * b.j0;
  const int j = b.jmin + l * b.i0 * b.i0;
  const int i0 = b.i0, j0 = b.j0, gh = s.g.gh;
  const double gam1 = gam - 1.0;
  const double THIRD = 1.0 / 3.0;

  VT roe = s.get(i, j, 0);
  VT roem1 = 1.0 / roe;
  VT ue = s.get(i, j, 1) * roem1;
  VT ve = s.get(i, j, 2) * roem1;
  VT we = s.get(i, j, 3) * roem1;
  VT ve2 = ue * ue + ve * ve + we * we;
  VT pe = gam1 * (s.get(i, j, 4) - 0.5 * roe * ve2);

  VT roe1 = s.get(i + i0, j + j0, 0);
  VT roe1m1 = 1.0 / roe1;
  VT ue1 = s.get(i + i0, j + j0, 1) * roe1m1;
  VT ve1 = s.get(i + i0, j + j0, 2) * roe1m1;
  VT we1 = s.get(i + i0, j + j0, 3) * roe1m1;
  VT ve21 = ue1 * ue1 + ve1 * ve1 + we1 * we1;
  VT pe1 = gam1 * (s.get(i + i0, j + j0, 4) - 0.5 * roe1 * ve21);

  VT pw = 1.125 * pe + (-0.125) * pe1;
  VT roi = 2.0 * pw / (rgaz * twall) - roe;
  VT pi = THIRD * (4.0 * pw - pe);
  VT roiei = pi / gam1;
  VT ui = -ue, vi = -ve, wi = -we;
  VT prev = roe;   // density one layer towards the interior of the ghost being written (da = de - 1)

  for (int de = 1; de <= gh; ++de) {
    s.set(i - de * i0, j - de * j0, 0, roi);
    s.set(i - de * i0, j - de * j0, 1, roi * ui);
    s.set(i - de * i0, j - de * j0, 2, roi * vi);
    s.set(i - de * i0, j - de * j0, 3, roi * wi);
    s.set(i - de * i0, j - de * j0, 4, roiei + 0.5 * roi * (ui * ui + vi * vi + wi * wi));
    VT next = 2.0 * roi - prev;
    prev = roi;
    roi = next;
    VT rm1 = 1.0 / s.get(i + de * i0, j + de * j0, 0);
    ui = -(s.get(i + de * i0, j + de * j0, 1) * rm1);
    vi = -(s.get(i + de * i0, j + de * j0, 2) * rm1);
    wi = -(s.get(i + de * i0, j + de * j0, 3) * rm1);
  }
}

// ---------------------------------------------------------------------------------------------
// profile walls of the sensitivity driver (BROADCAST_npz_sens.py:1763; card_bl2d_fv_npz_sens.py:105).  `prof` (lm values along the line)
// and the gas constants are ACTIVE inputs of the shipped tangents (tangent/bc_wall_blow_profile_d.f90: velprofd, gamd;
// tangent/bc_wall_viscous_iso_profile_d.f90: twallprofd, gamd, rgazd): profd / gamd / rgazd are their tangents along direction 0
// (null / 0 = passive); further directions of a vector-mode pass see them as passive.
//   BLOW = true : bc_wall_blow_profile.F90:36-95      wall-normal blowing velocity profile, isentropic ghost density, all layers alike
//   BLOW = false: bc_wall_viscous_iso_profile.F90     isothermal wall with a wall-temperature profile (as bc_wall_viscous_iso_line)
// ---------------------------------------------------------------------------------------------
template <int N, bool BLOW>
__device__ void bc_wall_profile_line(const StateRW<N>& s, const BcLine& b, const double* __restrict__ prof, const double* __restrict__ profd,
                                     double gam_, double gamd, double rgaz_, double rgazd, int l) {
  using DT = TanOf<N>;
  using VT = Var<DT>;
  const int i = b.imin + l * b.j0 * b.j0;
  const int j = b.jmin + l * b.i0 * b.i0;
  const int i0 = b.i0, j0 = b.j0, gh = s.g.gh;
  auto active = [](double v, double d) {
    VT r{v, DT{}};
    if constexpr (N > 0) r.d.d[0] = d;
    return r;
  };
  const VT gam = active(gam_, gamd), rgaz = active(rgaz_, rgazd), pr = active(prof[l], profd ? profd[l] : 0.0);
  const VT gam1 = gam - 1.0, gami = 1.0 / gam;
  const double THIRD = 1.0 / 3.0;

  const VT roe = s.get(i, j, 0);
  const VT roem1 = 1.0 / roe;
  const VT ue = s.get(i, j, 1) * roem1, ve = s.get(i, j, 2) * roem1, we = s.get(i, j, 3) * roem1;
  const VT pe = gam1 * (s.get(i, j, 4) - 0.5 * roe * (ue * ue + ve * ve + we * we));
  const VT roe1 = s.get(i + i0, j + j0, 0);
  const VT roe1m1 = 1.0 / roe1;
  const VT ue1 = s.get(i + i0, j + j0, 1) * roe1m1, ve1 = s.get(i + i0, j + j0, 2) * roe1m1, we1 = s.get(i + i0, j + j0, 3) * roe1m1;
  const VT pe1 = gam1 * (s.get(i + i0, j + j0, 4) - 0.5 * roe1 * (ue1 * ue1 + ve1 * ve1 + we1 * we1));
  const VT pw = 1.125 * pe + (-0.125) * pe1;
  const VT pi = THIRD * (4.0 * pw - pe);
  const VT roiei = pi / gam1;
  VT roi, ui = -ue, vi, wi = -we;
  if constexpr (BLOW) {
    roi = pow(pi * pow(roe, gam) / pe, gami);
    vi = 2.0 * pr - ve;
  } else {
    roi = 2.0 * pw / (rgaz * pr) - roe;
    vi = -ve;
  }
  VT prev = roe;
  for (int de = 1; de <= gh; ++de) {
    s.set(i - de * i0, j - de * j0, 0, roi);
    s.set(i - de * i0, j - de * j0, 1, roi * ui);
    s.set(i - de * i0, j - de * j0, 2, roi * vi);
    s.set(i - de * i0, j - de * j0, 3, roi * wi);
    s.set(i - de * i0, j - de * j0, 4, roiei + 0.5 * roi * (ui * ui + vi * vi + wi * wi));
    if constexpr (!BLOW) {
      const VT next = 2.0 * roi - prev;
      prev = roi;
      roi = next;
    }
    const VT rm1 = 1.0 / s.get(i + de * i0, j + de * j0, 0);
    ui = -(s.get(i + de * i0, j + de * j0, 1) * rm1);
    const VT vn = s.get(i + de * i0, j + de * j0, 2) * rm1;
    if constexpr (BLOW) vi = 2.0 * pr - vn; else vi = -vn;
    wi = -(s.get(i + de * i0, j + de * j0, 3) * rm1);
  }
}

// ---------------------------------------------------------------------------------------------
// symmetry plane (bc_symmetry.F90:36-76; tangent/bc_symmetry_d.f90): ghost de mirrors interior layer de - 1, the velocity reflected
// about the boundary-face normal, total energy corrected by the change of in-plane kinetic energy; rho w (plane 4) is NOT written.
// ANTI: bc_antisymmetry.F90:40-49 (tangent/bc_antisymmetry_d.f90) -- same reflection, ghost = (-rho, -(rho u'), rho v', -(E'))
// ---------------------------------------------------------------------------------------------
template <int N, bool ANTI = false>
__device__ void bc_symmetry_line(const StateRW<N>& s, const BcLine& b, const double* __restrict__ nx, const double* __restrict__ ny, int l) {
  using DT = TanOf<N>;
  using VT = Var<DT>;
  const int i1 = b.i0 * b.i0, j1 = b.j0 * b.j0;
  const int i = b.imin + l * j1;
  const int j = b.jmin + l * i1;
  const int i0 = b.i0, j0 = b.j0, gh = s.g.gh;
  const double sens = (double)(i0 + j0);
  const long long kn = s.g.nidx(i + b.high * i1, j + b.high * j1) + (long long)(b.kdir - 1) * s.g.sn;
  const double nxloc = nx[kn], nyloc = ny[kn];
  const double nsumi = 1.0 / ::sqrt(nxloc * nxloc + nyloc * nyloc);
  const double nxnorm = nxloc * nsumi * sens, nynorm = nyloc * nsumi * sens;
  for (int de = 1; de <= gh; ++de) {
    const int da = de - 1;
    const VT rho = s.get(i + da * i0, j + da * j0, 0), m1 = s.get(i + da * i0, j + da * j0, 1), m2 = s.get(i + da * i0, j + da * j0, 2);
    const VT rhoinv = 1.0 / rho;
    VT velx = m1 * rhoinv, vely = m2 * rhoinv;
    const VT veln = velx * nxnorm + vely * nynorm;
    velx = velx - 2.0 * veln * nxnorm;
    vely = vely - 2.0 * veln * nynorm;
    if constexpr (!ANTI) {
      const VT g1 = rho * velx, g2 = rho * vely;
      s.set(i - de * i0, j - de * j0, 0, rho);
      s.set(i - de * i0, j - de * j0, 1, g1);
      s.set(i - de * i0, j - de * j0, 2, g2);
      s.set(i - de * i0, j - de * j0, 4, s.get(i + da * i0, j + da * j0, 4) - 0.5 * ((m1 * m1 + m2 * m2) / rho) + 0.5 * ((g1 * g1 + g2 * g2) / rho));
    } else {
      const VT g0 = -rho, g1 = -(rho * velx), g2 = rho * vely;
      s.set(i - de * i0, j - de * j0, 0, g0);
      s.set(i - de * i0, j - de * j0, 1, g1);
      s.set(i - de * i0, j - de * j0, 2, g2);
      s.set(i - de * i0, j - de * j0, 4, -(s.get(i + da * i0, j + da * j0, 4) - 0.5 * ((m1 * m1 + m2 * m2) / rho) + 0.5 * ((g1 * g1 + g2 * g2) / g0)));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// pressure outlet (bc_pressure.F90:30-112; tangent/bc_pressure_d.f90): exterior state from the imposed pressure pext along the
// outgoing characteristics of the boundary cell; with noref the characteristic blend of bc_no_reflexion between the cell state and
// that exterior state (switches piecewise constant, not differentiated); every ghost layer gets the same state
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ void bc_pressure_line(const StateRW<N>& s, const BcLine& b, double pext, bool noref, double gam, const double* __restrict__ nx,
                                 const double* __restrict__ ny, int l) {
  using DT = TanOf<N>;
  using VT = Var<DT>;
  const int i1 = b.i0 * b.i0, j1 = b.j0 * b.j0;
  const int i = b.imin + l * j1;
  const int j = b.jmin + l * i1;
  const int i0 = b.i0, j0 = b.j0, gh = s.g.gh;
  const double gam1 = gam - 1.0;
  const double sens = (double)(i0 + j0);
  const long long kn = s.g.nidx(i + b.high * i1, j + b.high * j1) + (long long)(b.kdir - 1) * s.g.sn;
  const double nxloc = nx[kn], nyloc = ny[kn];
  const double nsumi = 1.0 / ::sqrt(nxloc * nxloc + nyloc * nyloc);
  const double nxnorm = nxloc * nsumi * sens, nynorm = nyloc * nsumi * sens;

  const VT ro0 = s.get(i, j, 0), w2 = s.get(i, j, 1), w3 = s.get(i, j, 2), w4 = s.get(i, j, 3), w5 = s.get(i, j, 4);
  const VT ro0m1 = 1.0 / ro0;
  const VT uu = w2 * ro0m1, vv = w3 * ro0m1, ww = w4 * ro0m1;
  const VT p0 = gam1 * (w5 - 0.5 * ro0 * (uu * uu + vv * vv + ww * ww));
  const VT c20 = gam * p0 * ro0m1;
  const VT c20m1 = 1.0 / c20;
  const VT roc0 = ro0 * sqrt(c20);
  const VT roc0m1 = 1.0 / roc0;
  const VT ros = ro0, us = uu, vs = vv, ws = ww;
  const VT ps = gam1 * (w5 - ros * 0.5 * (us * us + vs * vs + ws * ws));
  const VT vns = us * nxnorm + vs * nynorm;
  const VT uts = us - vns * nxnorm, vts = vs - vns * nynorm;
  const VT dp = pext - ps;
  const VT rod = ros + dp * c20m1;
  const VT vnd = vns + dp * roc0m1;
  const VT ud = uts + vnd * nxnorm, vd = vts + vnd * nynorm;
  VT ro, rou, rov, row, roe;
  if (noref) {
    const VT rovn0 = w2 * nxnorm + w3 * nynorm;
    const double epsm = 0.5 + fsign(0.5, roc0.v - rovn0.v);
    const double eps0 = 0.5 + fsign(0.5, -rovn0.v);
    const double epsp = 0.5 + fsign(0.5, -roc0.v - rovn0.v);
    const VT ut = eps0 * uts + (1.0 - eps0) * uts;
    const VT vt = eps0 * vts + (1.0 - eps0) * vts;
    const VT wt = eps0 * ws + (1.0 - eps0) * ws;
    const VT am = epsm * (ps - roc0 * vns) + (1.0 - epsm) * (pext - roc0 * vnd);
    const VT ap = epsp * (ps + roc0 * vns) + (1.0 - epsp) * (pext + roc0 * vnd);
    const VT vn = (ap - am) * 0.5 * roc0m1;
    const VT p = (ap + am) * 0.5;
    const VT bs = (p - ps) * ro0 * ro0 * roc0m1 * roc0m1 + ros;
    const VT b0 = (p - pext) * ro0 * ro0 * roc0m1 * roc0m1 + rod;
    ro = eps0 * bs + (1.0 - eps0) * b0;
    roe = p / gam1;
    rou = ro * (ut + vn * nxnorm);
    rov = ro * (vt + vn * nynorm);
    row = ro * wt;
  } else {
    ro = rod;
    rou = ro * ud;
    rov = ro * vd;
    row = ro * ws;
    roe = VT{pext / gam1, DT{}};
  }
  const VT rom1 = 1.0 / ro;
  const VT etot = roe + 0.5 * rom1 * (rou * rou + rov * rov + row * row);
  for (int de = 1; de <= gh; ++de) {
    s.set(i - de * i0, j - de * j0, 0, ro);
    s.set(i - de * i0, j - de * j0, 1, rou);
    s.set(i - de * i0, j - de * j0, 2, rov);
    s.set(i - de * i0, j - de * j0, 3, row);
    s.set(i - de * i0, j - de * j0, 4, etot);
  }
}

// ---------------------------------------------------------------------------------------------
// non-reflecting characteristic condition (bc_no_reflexion.F90:49-150)
//   wbd: (lm,5) reference state along the line, passive
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ void bc_no_reflexion_line(const StateRW<N>& s, const BcLine& b, const double* __restrict__ wbd, int lm,
                                     const double* __restrict__ nx, const double* __restrict__ ny, double gam, int l) {
  using DT = TanOf<N>;
  using VT = Var<DT>;
  const int i1 = b.i0 * b.i0, j1 = b.j0 * b.j0;
  const int i = b.imin + l * j1;
  const int j = b.jmin + l * i1;
  const int i0 = b.i0, j0 = b.j0, gh = s.g.gh;
  const double gam1 = gam - 1.0;
  const double sens = (double)(i0 + j0);

  const long long kn = s.g.nidx(i + b.high * i1, j + b.high * j1) + (long long)(b.kdir - 1) * s.g.sn;
  const double nxloc = nx[kn], nyloc = ny[kn];
  const double nsum = ::sqrt(nxloc * nxloc + nyloc * nyloc);
  const double nsumi = 1.0 / nsum;
  const double nxnorm = nxloc * nsumi * sens;
  const double nynorm = nyloc * nsumi * sens;

  // 0-state
  VT ro0 = s.get(i, j, 0);
  VT ro0m1 = 1.0 / ro0;
  VT w2 = s.get(i, j, 1), w3 = s.get(i, j, 2), w4 = s.get(i, j, 3), w5 = s.get(i, j, 4);
  VT uu = w2 * ro0m1, vv = w3 * ro0m1, ww = w4 * ro0m1;
  VT roe1 = w5 - 0.5 * ro0 * (uu * uu + vv * vv + ww * ww);
  VT p0 = gam1 * roe1;
  VT roc0 = ro0 * sqrt(gam * p0 * ro0m1);
  VT roc0m1 = 1.0 / roc0;
  VT rovn0 = w2 * nxnorm + w3 * nynorm;
  // wave-direction switches: piecewise constant, not differentiated
  const double epsm = 0.5 + fsign(0.5, roc0.v - rovn0.v);
  const double eps0 = 0.5 + fsign(0.5, -rovn0.v);
  const double epsp = 0.5 + fsign(0.5, -roc0.v - rovn0.v);

  // d-state (passive)
  const double rod = wbd[l + 0 * lm];
  const double rodm1 = 1.0 / rod;
  const double ud = wbd[l + 1 * lm] * rodm1;
  const double vd = wbd[l + 2 * lm] * rodm1;
  const double wdi = wbd[l + 3 * lm] * rodm1;
  const double pd = gam1 * (wbd[l + 4 * lm] - rod * 0.5 * (ud * ud + vd * vd + wdi * wdi));
  const double vnd = ud * nxnorm + vd * nynorm;
  const double utd = ud - vnd * nxnorm;
  const double vtd = vd - vnd * nynorm;

  // scheme state
  VT ros = ro0;
  VT rosm1 = 1.0 / ros;
  VT us = w2 * rosm1, vs = w3 * rosm1, ws = w4 * rosm1;
  VT ps = gam1 * (w5 - ros * 0.5 * (us * us + vs * vs + ws * ws));
  VT vns = us * nxnorm + vs * nynorm;
  VT uts = us - vns * nxnorm;
  VT vts = vs - vns * nynorm;

  VT ut = eps0 * uts + (1.0 - eps0) * utd;
  VT vt = eps0 * vts + (1.0 - eps0) * vtd;
  VT wt = eps0 * ws + (1.0 - eps0) * wdi;

  VT am = epsm * (ps - roc0 * vns) + (1.0 - epsm) * (pd - roc0 * vnd);
  VT ap = epsp * (ps + roc0 * vns) + (1.0 - epsp) * (pd + roc0 * vnd);
  VT vn = (ap - am) * 0.5 * roc0m1;
  VT p = (ap + am) * 0.5;

  VT bs = (p - ps) * ro0 * ro0 * roc0m1 * roc0m1 + ros;
  VT b0 = (p - pd) * ro0 * ro0 * roc0m1 * roc0m1 + rod;
  VT ro = eps0 * bs + (1.0 - eps0) * b0;
  VT rom1 = 1.0 / ro;
  VT roe = p / gam1;
  VT rou = ro * (ut + vn * nxnorm);
  VT rov = ro * (vt + vn * nynorm);
  VT row = ro * wt;
  VT roe2 = roe + 0.5 * rom1 * (rou * rou + rov * rov + row * row);

  for (int de = 1; de <= gh; ++de) {
    const int da = de - 1;
    s.set(i - de * i0, j - de * j0, 0, ro);
    s.set(i - de * i0, j - de * j0, 1, rou);
    s.set(i - de * i0, j - de * j0, 2, rov);
    s.set(i - de * i0, j - de * j0, 3, row);
    s.set(i - de * i0, j - de * j0, 4, roe2);
    ro = 2.0 * ro - s.get(i - da * i0, j - da * j0, 0);
    rou = 2.0 * rou - s.get(i - da * i0, j - da * j0, 1);
    rov = 2.0 * rov - s.get(i - da * i0, j - da * j0, 2);
    row = 2.0 * row - s.get(i - da * i0, j - da * j0, 3);
    roe2 = 2.0 * roe2 - s.get(i - da * i0, j - da * j0, 4);
  }
}

// ---------------------------------------------------------------------------------------------
// supersonic / subsonic inlet (bc_supandsubinlet.F90:39-169), quirks kept:
//   pr/pl without the factor 1/2 (:49,:57), field(j,i,.) indexing of the face state (:52-58),
//   velo2 = uu^2 + vv^2 + ww (:69), row = field(j,de,4) (:160)
//   field: (lm, gh, 5) passive Dirichlet table
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ void bc_supandsubinlet_line(const StateRW<N>& s, const BcLine& b, const double* __restrict__ field, int lm,
                                       const double* __restrict__ nx, const double* __restrict__ ny, double gam, int l) {
  using DT = TanOf<N>;
  using VT = Var<DT>;
  const int i = b.imin + l * b.j0 * b.j0;
  const int j = b.jmin + l * b.i0 * b.i0;
  const int i0 = b.i0, j0 = b.j0, gh = s.g.gh;
  const double gam1 = gam - 1.0;
  auto fld = [&](int a, int bb, int c) { return field[(a - 1) + (long long)(bb - 1) * lm + (long long)(c - 1) * lm * gh]; };

  // Roe face state: passive (only feeds the Mach test and the sign switches)
  double epsm = 0.0, eps0 = 0.0, epsp = 0.0, nxnorm = 0.0, nynorm = 0.0;
  bool subsonic;
  {
    const double wr1 = s.w[0 * s.g.sc + s.g.cidx(i, j)];
    const double rom1 = 1.0 / wr1;
    const double wr2 = s.w[1 * s.g.sc + s.g.cidx(i, j)] * rom1;
    const double wr3 = s.w[2 * s.g.sc + s.g.cidx(i, j)] * rom1;
    const double wr4 = s.w[3 * s.g.sc + s.g.cidx(i, j)] * rom1;
    const double w5 = s.w[4 * s.g.sc + s.g.cidx(i, j)];
    const double pr = gam1 * (w5 - wr1 * (wr2 * wr2 + wr3 * wr3 + wr4 * wr4));
    const double hr = (w5 + pr) * rom1;
    const double wl1 = fld(j, i, 1);
    const double rosm1 = 1.0 / wl1;
    const double wl2 = fld(j, i, 2) * rosm1;
    const double wl3 = fld(j, i, 3) * rosm1;
    const double wl4 = fld(j, i, 4) * rosm1;
    const double pl = gam1 * (fld(j, i, 5) - wl1 * (wl2 * wl2 + wl3 * wl3 + wl4 * wl4));
    const double hl = (fld(j, i, 5) + pl) * rosm1;
    const double r = ::sqrt(wr1 / wl1);
    const double rr = ::sqrt(wr1 * wl1);
    const double oneonrplusone = 1.0 / (r + 1.0);
    const double uu = (wr2 * r + wl2) * oneonrplusone;
    const double vv = (wr3 * r + wl3) * oneonrplusone;
    const double ww = (wr4 * r + wl4) * oneonrplusone;
    const double hh = (hr * r + hl) * oneonrplusone;
    const double velo2 = uu * uu + vv * vv + ww;
    const double ee = 0.5 * velo2;
    const double sound2 = gam1 * (hh - ee);
    const double mach = ::sqrt(velo2 / sound2);
    subsonic = mach < 1.0;
    if (subsonic) {
      const long long kn = s.g.nidx(i, j) + (long long)(b.kdir - 1) * s.g.sn;
      const double nxloc = nx[kn], nyloc = ny[kn];
      const double nsum = 1.0 / ::sqrt(nxloc * nxloc + nyloc * nyloc);
      nxnorm = nxloc * nsum;
      nynorm = nyloc * nsum;
      const double roc0 = rr * ::sqrt(sound2);
      const double rovn0 = rr * (uu * nxnorm + vv * nynorm);
      epsm = 0.5 + fsign(0.5, roc0 - rovn0);
      eps0 = 0.5 + fsign(0.5, -rovn0);
      epsp = 0.5 + fsign(0.5, -roc0 - rovn0);
    }
  }

  for (int de = 1; de <= gh; ++de) {
    const int ig = i - de * i0, jg = j - de * j0;
    VT q[5];
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      q[e].v = fld(l + 1, de, e + 1);
      q[e].d = TanTraits<DT>::zero();
      s.set(ig, jg, e, q[e]);
    }
    const int da = de - 1;
    if (subsonic) {
      // l-state: the Dirichlet ghost just written (tangent zero)
      VT rod = q[0];
      VT rom1 = 1.0 / rod;
      VT ud = q[1] * rom1, vd = q[2] * rom1, wd_ = q[3] * rom1;
      VT pd = gam1 * (q[4] - rod * 0.5 * (ud * ud + vd * vd + wd_ * wd_));
      VT vnd = ud * nxnorm + vd * nynorm;
      VT utd = ud - vnd * nxnorm;
      VT vtd = vd - vnd * nynorm;
      VT hl = (q[4] + pd) * rom1;
      // r-state: previous cell towards the interior
      const int ir = i - da * i0, jr = j - da * j0;
      VT ros = s.get(ir, jr, 0);
      VT rosm1 = 1.0 / ros;
      VT us = s.get(ir, jr, 1) * rosm1, vs = s.get(ir, jr, 2) * rosm1, ws = s.get(ir, jr, 3) * rosm1;
      VT e5 = s.get(ir, jr, 4);
      VT ps = gam1 * (e5 - ros * 0.5 * (us * us + vs * vs + ws * ws));
      VT vns = us * nxnorm + vs * nynorm;
      VT uts = us - vns * nxnorm;
      VT vts = vs - vns * nynorm;
      VT hr = (e5 + ps) * rosm1;

      VT r = sqrt(ros * rom1);
      VT rr = sqrt(ros * rod);
      VT oneonrplusone = 1.0 / (r + 1.0);
      VT uu = (us * r + ud) * oneonrplusone;
      VT vv = (vs * r + vd) * oneonrplusone;
      VT ww = (ws * r + wd_) * oneonrplusone;
      VT ee = 0.5 * (uu * uu + vv * vv + ww * ww);
      VT hh = (hr * r + hl) * oneonrplusone;
      VT sound2 = gam1 * (hh - ee);
      VT roc0 = rr * sqrt(sound2);
      VT roc0m1 = 1.0 / roc0;

      VT ut = eps0 * uts + (1.0 - eps0) * utd;
      VT vt = eps0 * vts + (1.0 - eps0) * vtd;
      VT am = epsm * (ps - roc0 * vns) + (1.0 - epsm) * (pd - roc0 * vnd);
      VT ap = epsp * (ps + roc0 * vns) + (1.0 - epsp) * (pd + roc0 * vnd);
      VT vn = (ap - am) * 0.5 * roc0m1;
      VT p = (ap + am) * 0.5;
      VT bs = (p - ps) * rr * rr * roc0m1 * roc0m1 + ros;
      VT b0 = (p - pd) * rr * rr * roc0m1 * roc0m1 + rod;
      VT ro = eps0 * bs + (1.0 - eps0) * b0;
      VT roe = p / gam1;
      VT rou = ro * (ut + vn * nxnorm);
      VT rov = ro * (vt + vn * nynorm);
      VT row;
      row.v = fld(j, de, 4);
      row.d = TanTraits<DT>::zero();
      s.set(ig, jg, 0, ro);
      s.set(ig, jg, 1, rou);
      s.set(ig, jg, 2, rov);
      s.set(ig, jg, 3, row);
      s.set(ig, jg, 4, roe + 0.5 * (rou * rou + rov * rov + row * row) / ro);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// linear extrapolation (bc_extrapolate.F90:56-69); tangent mode writes wd only
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ void bc_extrapolate_o2_line(const StateRW<N>& s, const BcLine& b, int l) {
  const int i = b.imin + l * b.j0 * b.j0;
  const int j = b.jmin + l * b.i0 * b.i0;
  const int i0 = b.i0, j0 = b.j0, gh = s.g.gh;
  for (int de = 1; de <= gh; ++de) {
    const int d1 = de - 1, d2 = de - 2;
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      auto x = 2.0 * s.get(i - i0 * d1, j - j0 * d1, e) - s.get(i - i0 * d2, j - j0 * d2, e);
      s.set(i - i0 * de, j - j0 * de, e, x, /*primal=*/N == 0);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Dirichlet fill from a table (bc_general.F90:6-31; tangent/bc_general_d.f90): ghost layer de of line cell l takes field(l, de, :);
// the tangent routine zeroes the ghost tangents and leaves w alone.  field: (lm, gh, 5), Fortran order.
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ void bc_general_line(const StateRW<N>& s, const BcLine& b, const double* __restrict__ field, int lm, int l) {
  using DT = TanOf<N>;
  const int i = b.imin + l * b.j0 * b.j0;
  const int j = b.jmin + l * b.i0 * b.i0;
  const int i0 = b.i0, j0 = b.j0, gh = s.g.gh;
  for (int de = 1; de <= gh; ++de) {
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      Var<DT> x{};
      x.v = field[l + (long long)(de - 1) * lm + (long long)e * lm * gh];
      s.set(i - i0 * de, j - j0 * de, e, x, /*primal=*/N == 0);
    }
  }
}

}  // namespace bcast
