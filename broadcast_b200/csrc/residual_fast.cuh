// Second-generation fused primal residual (order 5, gh = 3): the tile algorithm and its face formulas, written as
// host+device phase functions so that the very same code runs inside k_residual_fast (residual_fast.cu) and, on a
// machine without a GPU, inside tests/host/residual_fast_host.cpp against the oracle.
//
// What changed against k_residual_tile (residual_tile.cu), which evaluates the reference-shaped templates of
// scheme.cuh face by face (ncu, profiles/r1_c_summary.md: shared-memory wavefronts 59 %, FP64 pipe 44 %, ~450 LDS.64
// and ~1 400 FP64 instructions per cell):
//   * convective flux as (u nx + v ny) * [rho, rho u, rho v, rho w, rho E + p] + p n, one pass over the six stencil
//     cells shared with the 5th-difference operand (euler_o6_{i,j}.F + predictor_7p_{i,j}.F): 8 loads per cell;
//   * the normal-direction interpolation (-1, 9, 9, -1) of u, v, w, T that the compact viscous gradients apply on five
//     cross rows (flux_visqueux_o4_{i,j}.F) is a per-FACE quantity: it is computed once per face into shared memory
//     and read by the four cross neighbours (84 -> ~36 loads per face, ~216 -> ~110 flops);
//   * sqrt(rho) and the sound speed are per-cell quantities: the Roe weights become sl/(sl+sr), the sensor
//     denominators a*|n| (spectralradius_{i,j}.F, ducrosfordnc_{i,j}.F): 5 -> 2 square roots per face;
//   * max(k1, k2) of the Jameson sensor and min(x0, x1) of the dilatation switch are selected by cross-multiplication
//     before dividing; (1 - tanh x)/2 = 1/(1 + exp 2x); reciprocals by MUFU seed + two Newton steps;
//   * the face metric scalings (1/24, 1/(12*16), 1/vol_face, 1/2) are folded into the eight dual-cell normals;
//     (fv nx/|n| + gv ny/|n|) |n| is evaluated as fv nx + gv ny (fluxnumassembly_{i,j}.F).
// All of these are re-associations: results agree with the reference to a few ulp of the largest term (tests: 1e-12 of
// the plane maximum).  The three wall rows (o2 viscous fluxes, off-centred fluxes, wall flux: flux_num_dnc5.F90:161-220)
// keep the reference-shaped templates of scheme.cuh through an accessor over the same shared arrays.
//
// Tile: 32 x OJ output cells per CTA of 32 (OJ + 1) threads: OJ x 33 i-faces (OJ warps x 32 left faces + OJ lanes of the
// last warp for the last column) and (OJ + 1) x 32 j-faces (all threads) -- no face is evaluated twice inside a CTA.
// OJ = 9 (320 threads, 10 warps): two CTAs put 5 warps on every SM sub-partition, the most that 96 registers allow.
#pragma once
#include "scheme.cuh"

// The tile height is a compile-time constant; a translation unit that wants another one (residual_fast_tma.cu: OJ = 8, so that
// two w buffers still leave room for two CTAs per SM) defines BCAST_RF_OJ and its own namespace name BCAST_RF_NS.
#ifndef BCAST_RF_NS
#define BCAST_RF_NS rf
#endif

// A translation unit that defines BCAST_RF_DUAL gets the same algorithm in forward-mode tangent arithmetic (value + ONE tangent
// direction per scalar: residual_tangent.cu, the strip / colour-loop tangent kernel): `real` is then a dual number, every state-
// dependent scalar below is a `real`, the mesh metrics stay double.  Non-smooth intrinsics follow the conventions of the
// reference's Tapenade tangent (dual.cuh): abs' by the sign of x (x >= 0 -> +), max takes the second operand iff the first is
// smaller, sqrt'(0) = 0.
namespace bcast {
namespace BCAST_RF_NS {

#ifdef BCAST_RF_DUAL
struct Dual {
  double v, d;
};
using real = Dual;
BC_HD Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
BC_HD Dual operator+(Dual a, double b) { return {a.v + b, a.d}; }
BC_HD Dual operator+(double a, Dual b) { return {a + b.v, b.d}; }
BC_HD Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
BC_HD Dual operator-(Dual a, double b) { return {a.v - b, a.d}; }
BC_HD Dual operator-(double a, Dual b) { return {a - b.v, -b.d}; }
BC_HD Dual operator-(Dual a) { return {-a.v, -a.d}; }
BC_HD Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
BC_HD Dual operator*(Dual a, double b) { return {a.v * b, a.d * b}; }
BC_HD Dual operator*(double a, Dual b) { return {a * b.v, a * b.d}; }
BC_HD Dual operator/(Dual a, double b) { return {a.v / b, a.d / b}; }
BC_HD Dual& operator+=(Dual& a, Dual b) { a.v += b.v; a.d += b.d; return a; }
BC_HD double rf_val(Dual x) { return x.v; }
BC_HD double rf_out(Dual x) { return x.d; }   // what a tangent kernel stores
BC_HD Dual rf_const(double x) { return {x, 0.0}; }
BC_HD Dual rf_sqrt(Dual x) {
  const double s = ::sqrt(x.v);
  return {s, x.v == 0.0 ? 0.0 : x.d / (2.0 * s)};
}
BC_HD Dual rf_abs(Dual x) { return {::fabs(x.v), x.v >= 0.0 ? x.d : -x.d}; }
BC_HD Dual rf_max(Dual a, Dual b) { return a.v < b.v ? b : a; }
BC_HD Dual rf_max(double a, Dual b) { return a < b.v ? b : Dual{a, 0.0}; }
BC_HD Dual rf_min(Dual a, double b) { return a.v < b ? a : Dual{b, 0.0}; }
BC_HD Dual rf_exp(Dual x) {
  const double e = ::exp(x.v);
  return {e, e * x.d};
}
using RfDT = Tan<1>;
BC_HD Var<Tan<1>> rf_to_var(Dual x) {
  Var<Tan<1>> r;
  r.v = x.v;
  r.d.d[0] = x.d;
  return r;
}
BC_HD Dual rf_from_var(const Var<Tan<1>>& x) { return {x.v, x.d.d[0]}; }
BC_HD Dual rf_from_var(const Var<Zero>& x) { return {x.v, 0.0}; }
#else
using real = double;
BC_HD double rf_val(double x) { return x; }
BC_HD double rf_out(double x) { return x; }
BC_HD double rf_const(double x) { return x; }
BC_HD double rf_sqrt(double x) { return ::sqrt(x); }
BC_HD double rf_abs(double x) { return ::fabs(x); }
BC_HD double rf_max(double a, double b) { return ::fmax(a, b); }
BC_HD double rf_min(double a, double b) { return ::fmin(a, b); }
BC_HD double rf_exp(double x) { return ::exp(x); }
using RfDT = Zero;
BC_HD PVar rf_to_var(double x) { return PVar{x, {}}; }
BC_HD double rf_from_var(const PVar& x) { return x.v; }
#endif

constexpr int H = 3;
#ifndef BCAST_RF_OJ
#define BCAST_RF_OJ 9
#endif
constexpr int OI = 32, OJ = BCAST_RF_OJ;
constexpr int PI = OI + 2 * H, PJ = OJ + 2 * H;  // staged cells: i0-3 .. i0+34, j0-3 .. j0+OJ+2
constexpr int NC = PI * PJ;
constexpr int NT = OI * (OJ + 1);                // one thread per j-face; OJ rows of i-faces + OJ lanes for the last column
// derived per-cell arrays (the five planes of w live in their own buffer so that TMA can deliver them, double buffered)
enum { A_U = 0, A_V = 1, A_WZ = 2, A_T = 3, A_P = 4, A_MU = 5, A_SR = 6, A_CS = 7, A_DV = 8, A_DU = 9, NARR = 10 };
constexpr int WBUF = (5 * NC + 15) / 16 * 16;    // doubles per w buffer: 5 * NC rounded up to a multiple of 128 bytes
// normal-direction interpolations R_q (q = u, v, w, T) of the faces a CTA's viscous gradients read
constexpr int RI_W = OI + 1, RI_H = OJ + 4;   // i-faces i0 .. i0+32, rows j0-2 .. j0+9
constexpr int RJ_W = OI + 4, RJ_H = OJ + 1;   // j-faces columns i0-2 .. i0+33, rows j0 .. j0+8
constexpr int RQ = RI_W * RI_H;               // 396 >= 324
constexpr int NRB = 4 * RQ;
constexpr int XI_P = OI + 1, XJ_P = OI;       // pitches of the face-flux exchange buffer
constexpr int NXB = 5 * (OJ + 1) * OI;        // >= 5 * OJ * (OI + 1)
constexpr int GW = OI + 2, GH_ = OJ + 2;      // sensor cells: i0-1 .. i0+32, j0-1 .. j0+8 (scratch divu / vort alias X)
constexpr int NSM_REST = NARR * NC + NRB + NXB;
constexpr int NSM = WBUF + NSM_REST;          // doubles of shared memory per CTA, one w buffer (88 128 bytes)
constexpr int NSM_TMA = 2 * WBUF + NSM_REST + 2;   // two w buffers + two mbarriers (109 520 bytes)
// ---- bulk-staged variant (k_residual_fast_bulk): the mesh metrics of the tile in shared memory too ---------------------------------
//   vol   box (OI+2) x (OJ+2)      cells i0-1 .. i0+32, j0-1 .. j0+OJ          (sensor cells)            TMA 2-D box
//   volf  box (OI+4) x (OJ+1) x 2  cells i0-1 .. i0+34, j0 .. j0+OJ            (faces)                   TMA 3-D box
//         (box origins sit on even storage columns: TMA faults on a box whose first byte is not 16-byte aligned)
//   node  4 planes (nx0, nx1, ny0, ny1) x (OJ+3) rows j0-1 .. j0+OJ+1, columns i0-1 .. i0+34.  Node planes have an odd leading
//         dimension ldn on even grids, which no tensor map can describe (strides must be multiples of 16 bytes) -- but the array
//         seen as rows of 2 ldn elements can: node row s of plane k is then the half (t & 1) of "double row" t >> 1 with
//         t = k (nj + 1) + s, and the rows of one parity class of a tile are consecutive double rows.  Two TMA boxes of
//         MN_SLOT x MN_HALF per plane (one per parity class) deliver the window; a box starts at the even element at or below its
//         first element, so its rows sit shifted by `sh` in {0, 1} entries and readers add the shift.  (The first version issued
//         one 1-D bulk copy per row: 48 copies, each an election loop of ~13 instructions on one warp -- 1 300 warp instructions
//         before the first byte moved, profiles/r2_b_summary.md.)
constexpr int MV_W = OI + 2, MV_H = OJ + 2;
constexpr int MF_W = OI + 4, MF_H = OJ + 1;   // from cell i0-1: the first coordinate of a TMA box must be a multiple of 16 bytes
constexpr int MN_ROWS = OJ + 3, MN_HALF = (MN_ROWS + 1) / 2, MN_SLOT = OI + 4;   // columns i0-1 .. i0+33 (+ shift)
constexpr int up16(int n) { return (n + 15) / 16 * 16; }
// the vol box is read by the sensor phase only: it lives in the part of the flux-exchange buffer X that phases 0-1 do not use
// (their divu / vort scratch takes the first 2 GW GH_ entries), at the first 128-byte aligned offset behind that scratch
constexpr int O_X = WBUF + NARR * NC + NRB;      // offset of X() from the start of shared memory
constexpr int O_VOLBOX = up16(O_X + 2 * GW * GH_);
static_assert(O_VOLBOX + MV_W * MV_H <= O_X + NXB, "vol box inside the exchange buffer");
constexpr int MF_PS = MF_W * MF_H;                // plane stride inside the volf box
constexpr int M_VOLF = 0, M_NODE = up16(2 * MF_PS), MN_BOX = up16(MN_HALF * MN_SLOT), NMET = M_NODE + 8 * MN_BOX;   // (a TMA box lands on a 128-byte aligned address)
constexpr int O_MET = up16(NSM);                 // metric region behind the arrays of the LDG kernel (128-byte aligned)
constexpr int NSM_BULK = O_MET + NMET + 2;       // + one mbarrier
static_assert(2 * ((long long)NSM_BULK * 8 + 1024) <= 228 * 1024, "two CTAs per SM");
static_assert((MV_W * 8) % 16 == 0 && (MF_W * 8) % 16 == 0 && (MN_SLOT * 8) % 16 == 0, "box rows: multiples of 16 bytes");
static_assert(OI + 3 + 1 <= MN_SLOT, "node row window: columns a = 0 .. OI + 2 and the shift");
constexpr int BULK_BYTES = (5 * NC + MV_W * MV_H + 2 * MF_W * MF_H + 8 * MN_HALF * MN_SLOT) * 8;   // what one tile receives
static_assert(5 * NC <= WBUF, "w buffer");
static_assert(5 * OJ * XI_P <= NXB, "exchange buffer");
static_assert(OJ <= 16 && NT % 32 == 0 && NT / 32 == OJ + 1 && 2 * OI <= NT && OI == 32, "thread mappings (warp-aligned sensor / R_q tasks)");
static_assert(2 * GW * GH_ <= NXB, "scratch aliasing");
static_assert(RJ_W * RJ_H <= RQ, "R buffer");

// reciprocal: MUFU seed (20 mantissa bits) + two Newton steps (device); plain division on the host build
BC_HD double frcp(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
#else
  return 1.0 / x;
#endif
}

BC_HD real rf_rcp(real x) {
#ifdef BCAST_RF_DUAL
  const double r = frcp(x.v);
  return {r, -r * r * x.d};
#else
  return frcp(x);
#endif
}

struct TileCtx {
  real* wsm;  // the five planes of w of the tile, [e][b][a]  (exactly the box a 3-D TMA load of w delivers)
  real* sm;   // derived arrays, R buffer, exchange buffer
  const GridDesc& g;      // (references: in the kernels these are the kernel parameters, read from the constant bank)
  const SchemeConsts& c;
  double sqgr;  // sqrt(gam * rgaz)
  bool wall;
  const double *w, *nx, *ny, *vol, *volf;
  const double* wd = nullptr;   // tangent build: the five planes of the direction at hand
  double* res;
  int i0, j0;  // first output cell of the tile
  int i1 = 1 << 30, j1 = 1 << 30;   // last cell the tile may write (tangent build: the rows of a rectangle)
  unsigned char* flags = nullptr;   // tangent build: per staged cell, 1 if any of its five tangents is non-zero (face skipping)
  const double* met = nullptr;      // bulk-staged variant: the metric region (volf box, node rows)
  const double* volbox = nullptr;   // bulk-staged variant: the vol box (inside the exchange buffer, dead after the sensor phase)
  int nsh = 0;                      // bulk-staged variant: bit 2 k + p = shift of the rows of node plane k, parity class p (node_shift_mask)
  BC_HD TileCtx(const GridDesc& g_, const SchemeConsts& c_) : g(g_), c(c_) {}
  BC_HD real* arr(int a) const { return sm + a * NC; }
  BC_HD real* RB() const { return sm + NARR * NC; }
  BC_HD real* X() const { return sm + NARR * NC + NRB; }
  // does the tile hold sensor cells of the first ghost layer of a physical boundary?
  BC_HD bool has_ghost_sensor() const {
    return (i0 == 1 && !(g.edges & 1)) || (i0 + OI >= g.im + 1 && !(g.edges & 2)) || j0 == 1 || j0 + OJ >= g.jm + 1;
  }
};

// accessor over the shared arrays for the reference-shaped templates of scheme.cuh (wall rows, gradients)
struct SmemAcc2 {
  using DT = RfDT;
  const real* s;   // sm + shared index of the base cell
  const real* sw;  // wsm + shared index of the base cell
  const double *nx, *ny, *vol, *volf;
  long long c, n;
  int ldc, ldn;
  long long sc, sn;
  using VT = Var<RfDT>;
  template <int OI_, int OJ_> BC_HD real raw(int a) const { return s[a * NC + OI_ + OJ_ * PI]; }
  template <int OI_, int OJ_> BC_HD VT ld(int a) const { return rf_to_var(raw<OI_, OJ_>(a)); }
  template <int OI_, int OJ_> BC_HD real rawW(int e) const { return sw[e * NC + OI_ + OJ_ * PI]; }
  template <int OI_, int OJ_> BC_HD VT W(int e) const { return rf_to_var(rawW<OI_, OJ_>(e)); }
  template <int OI_, int OJ_> BC_HD VT U() const { return ld<OI_, OJ_>(A_U); }
  template <int OI_, int OJ_> BC_HD VT V() const { return ld<OI_, OJ_>(A_V); }
  template <int OI_, int OJ_> BC_HD VT Wz() const { return ld<OI_, OJ_>(A_WZ); }
  template <int OI_, int OJ_> BC_HD VT T() const { return ld<OI_, OJ_>(A_T); }
  template <int OI_, int OJ_> BC_HD VT P() const { return ld<OI_, OJ_>(A_P); }
  template <int OI_, int OJ_> BC_HD VT Mu() const { return ld<OI_, OJ_>(A_MU); }
  template <int OI_, int OJ_> BC_HD VT H() const { return (W<OI_, OJ_>(4) + P<OI_, OJ_>()) * (1.0 / W<OI_, OJ_>(0)); }
  template <int OI_, int OJ_> BC_HD auto SENS() const {   // A_DV holds vol * divu
    return CellSens<RfDT, RfDT>{ld<OI_, OJ_>(A_DV) / VOL<OI_, OJ_>(), ld<OI_, OJ_>(A_DU)};
  }
  template <int OI_, int OJ_> BC_HD double NX(int kk) const { return BC_LDG(nx + kk * sn + n + OI_ + (long long)OJ_ * ldn); }
  template <int OI_, int OJ_> BC_HD double NY(int kk) const { return BC_LDG(ny + kk * sn + n + OI_ + (long long)OJ_ * ldn); }
  template <int OI_, int OJ_> BC_HD double VOL() const { return BC_LDG(vol + c + OI_ + (long long)OJ_ * ldc); }
  template <int OI_, int OJ_> BC_HD double VOLF(int kk) const { return BC_LDG(volf + kk * sc + c + OI_ + (long long)OJ_ * ldc); }
};

BC_HD SmemAcc2 make_acc(const TileCtx& t, int a, int b) {  // shared coordinates (a, b): cell (i0-H+a, j0-H+b)
  SmemAcc2 A;
  A.s = t.sm + a + b * PI;
  A.sw = t.wsm + a + b * PI;
  A.nx = t.nx; A.ny = t.ny; A.vol = t.vol; A.volf = t.volf;
  A.c = t.g.cidx(t.i0 - H + a, t.j0 - H + b);
  A.n = t.g.nidx(t.i0 - H + a, t.j0 - H + b);
  A.ldc = t.g.ldc; A.ldn = t.g.ldn; A.sc = t.g.sc; A.sn = t.g.sn;
  return A;
}

// ---- phase 0: stage w, cell primitives (phys/Primitives.F:2-34, phys/viscosity.F:1) -------------------------------
// STAGED: the five planes of w are already in t.wsm (TMA); otherwise they are loaded from global memory here.
template <bool STAGED>
BC_HD void phase0(const TileCtx& t, int tid) {
  const GridDesc& g = t.g;
  constexpr int NIT = (NC + NT - 1) / NT;
  real q[NIT][5];
  // all loads of the thread first (the only HBM reads of the kernel besides the metrics), then the arithmetic
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = tid + it * NT;
    const int a = idx % PI, b = idx / PI;
    const int gi = t.i0 - H + a, gj = t.j0 - H + b;
    q[it][0] = rf_const(1.0); q[it][1] = rf_const(0.0); q[it][2] = rf_const(0.0); q[it][3] = rf_const(0.0); q[it][4] = rf_const(1.0);
    if (STAGED ? idx < NC : (idx < NC && gi <= g.im + g.gh && gj <= g.jm + g.gh)) {   // cells beyond the padded array get a sane state
#ifdef BCAST_RF_DUAL
      const long long k = g.cidx(gi, gj);
#pragma unroll
      for (int e = 0; e < 5; ++e) q[it][e] = Dual{BC_LDG(t.w + e * g.sc + k), BC_LDG(t.wd + e * g.sc + k)};
#else
      if constexpr (STAGED) {
#pragma unroll
        for (int e = 0; e < 5; ++e) q[it][e] = t.wsm[e * NC + idx];
        if (q[it][0] == 0.0) { q[it][0] = 1.0; q[it][4] = 1.0; }   // TMA zero fill outside the array (a real density is positive)
      } else {
        const double* p = t.w + g.cidx(gi, gj);
#pragma unroll
        for (int e = 0; e < 5; ++e) q[it][e] = BC_LDG(p + e * g.sc);
      }
#endif
    }
  }
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = tid + it * NT;
    if (idx >= NC) break;
    const real q0 = q[it][0], q1 = q[it][1], q2 = q[it][2], q3 = q[it][3], q4 = q[it][4];
#ifdef BCAST_RF_DUAL
    if (t.flags) t.flags[idx] = (q0.d != 0.0 || q1.d != 0.0 || q2.d != 0.0 || q3.d != 0.0 || q4.d != 0.0) ? 1 : 0;
#endif
    const real rom1 = rf_rcp(q0);
    const real u = q1 * rom1, v = q2 * rom1, wz = q3 * rom1;
    const real ec = 0.5 * (u * u + v * v + wz * wz);
    const real eloc = (q4 - ec * q0) * rom1;
    const real tl = eloc * t.c.cvm1;
    const real p = t.c.gam1 * q0 * eloc;
    const real sqt = rf_sqrt(tl);
    if constexpr (!STAGED) {
      real* sw = t.wsm + idx;
      sw[0] = q0; sw[NC] = q1; sw[2 * NC] = q2; sw[3 * NC] = q3; sw[4 * NC] = q4;
    }
    real* s = t.sm + idx;
    s[A_U * NC] = u;
    s[A_V * NC] = v;
    s[A_WZ * NC] = wz;
    s[A_T * NC] = tl;
    s[A_P * NC] = p;
    s[A_MU * NC] = t.c.betas * rf_rcp(tl + t.c.s_suth) * sqt * tl;
    s[A_SR * NC] = rf_sqrt(q0);
    s[A_CS * NC] = t.sqgr * sqt;
  }
}

// ---- phase 1: sensor cells (dilatation, Ducros ratio) and the R_q of the i-faces ---------------------------------
BC_HD real ducros_ratio(real divu, real vort) {
  const real d2 = divu * divu;
  return d2 * rf_rcp(d2 + vort * vort + 1e-15);
}
// R_q(face) = -q(-2) + 9 q(-1) + 9 q(0) - q(1) along the face normal (the 1/16 is applied by the consumer)
BC_HD real rrow(const real* q, int stride) { return 9.0 * (q[-stride] + q[0]) - (q[-2 * stride] + q[stride]); }

// Sensor cells of a tile: rows j0 .. j0+OJ-1 over columns i0-1 .. i0+32 (34 OJ cells, one per thread) and the two rows
// j0-1, j0+OJ over columns i0 .. i0+31 (64 cells, a second round of the last two warps); the corners are never read.
// cell metrics of the 5-point gradient (geom/dxdy.F:1-6), loaded before phase 0 so that their latency is hidden
struct SensGeom {
  double dxm1, dxm2, dym1, dym2, vol;
  bool valid;
};
BC_HD bool sensor_of(const TileCtx& t, int tid, int round, int& ga, int& gb) {   // window coordinates of the thread's sensor cell
  // Warp-aligned mapping (r2_c: the former `tid % 34` mapping let every warp straddle two rows -- 28 % excess shared-memory
  // wavefronts in this phase, ncu r2_15 -- and gave warp OJ - 1 three work items where the others have two):
  //   round 0: warp r < OJ owns row j0 + r, columns i0-1 .. i0+30 (one per lane); the last warp owns the two columns i0+31, i0+32
  //            of all OJ rows (2 OJ lanes)
  //   round 1: the two rows j0-1 and j0+OJ, columns i0 .. i0+31: warps OJ - 1 and OJ (neither has a main R_q task in phase 1)
  if (round == 0) {
    const int wrp = tid >> 5, lane = tid & 31;
    if (wrp < OJ) { ga = lane; gb = 1 + wrp; }
    else {
      if (lane >= 2 * OJ) return false;
      ga = OI + lane / OJ;
      gb = 1 + lane % OJ;
    }
  } else {
    const int u = tid - (NT - 2 * OI);
    if (u < 0) return false;
    ga = 1 + (u & (OI - 1));
    gb = u < OI ? 0 : GH_ - 1;
  }
  const GridDesc& g = t.g;
  const int ci = t.i0 - 1 + ga, cj = t.j0 - 1 + gb;
  return ci >= g.glo() && ci <= g.ghi() && cj >= 1 && cj <= g.jm;   // slab-internal edges: real gradients in the halo column
}
BC_HD SensGeom prefetch_sensor(const TileCtx& t, int tid, int round) {
  SensGeom G{};
  int ga, gb;
  G.valid = sensor_of(t, tid, round, ga, gb);
  if (G.valid) {
    const GridDesc& g = t.g;
    const int ci = t.i0 - 1 + ga, cj = t.j0 - 1 + gb;
    const long long n = g.nidx(ci, cj);
    G.vol = BC_LDG(t.vol + g.cidx(ci, cj));
    const double volm1 = frcp(G.vol);
    G.dxm1 = 0.5 * (BC_LDG(t.nx + n) + BC_LDG(t.nx + n + 1)) * volm1;
    G.dxm2 = 0.5 * (BC_LDG(t.nx + g.sn + n) + BC_LDG(t.nx + g.sn + n + g.ldn)) * volm1;
    G.dym1 = 0.5 * (BC_LDG(t.ny + n) + BC_LDG(t.ny + n + 1)) * volm1;
    G.dym2 = 0.5 * (BC_LDG(t.ny + g.sn + n) + BC_LDG(t.ny + g.sn + n + g.ldn)) * volm1;
  }
  return G;
}
// dilatation, vorticity and Ducros ratio of one sensor cell (gradop_5pi.F, gradop_5pj.F, gradient.F: same operation order
// as cell_gradients() of scheme.cuh)
BC_HD void sensor_cell(const TileCtx& t, int tid, int round, const SensGeom& G) {
  if (!G.valid) return;
  int ga, gb;
  sensor_of(t, tid, round, ga, gb);
  const int k = (ga + H - 1) + (gb + H - 1) * PI;
  const real* U = t.arr(A_U) + k;
  const real* V = t.arr(A_V) + k;
  constexpr double b1 = 8.0 * (1.0 / 12.0), b2 = -(1.0 / 12.0);
  const real gui = b1 * (U[1] - U[-1]) + b2 * (U[2] - U[-2]);
  const real gvi = b1 * (V[1] - V[-1]) + b2 * (V[2] - V[-2]);
  const real guj = b1 * (U[PI] - U[-PI]) + b2 * (U[2 * PI] - U[-2 * PI]);
  const real gvj = b1 * (V[PI] - V[-PI]) + b2 * (V[2 * PI] - V[-2 * PI]);
  const real gu0 = G.dxm1 * gui + G.dxm2 * guj, gv0 = G.dxm1 * gvi + G.dxm2 * gvj;
  const real gu1 = G.dym1 * gui + G.dym2 * guj, gv1 = G.dym1 * gvi + G.dym2 * gvj;
  const real divu = gu0 + gv1, vort = gv0 - gu1;
  real* S0 = t.X();
  S0[ga + gb * GW] = divu;
  S0[GW * GH_ + ga + gb * GW] = vort;
  t.arr(A_DV)[k] = G.vol * divu;
  t.arr(A_DU)[k] = ducros_ratio(divu, vort);
}

BC_HD void phase1(const TileCtx& t, int tid, const SensGeom& G0, const SensGeom& G1) {
  sensor_cell(t, tid, 0, G0);
  sensor_cell(t, tid, 1, G1);
  // R_q of the i-faces (face columns i0 .. i0+32, rows j0-2 .. j0+OJ+1): eight tasks (quantity q, half of the face rows) of 32
  // columns each, one per warp (lane = column: conflict free), on the warps before the last one; the last warp takes the 33rd
  // column (4 RI_H values)
  const int wrp = tid >> 5, lane = tid & 31;
  constexpr int NW = NT / 32, RW = NW - 1 < 8 ? NW - 1 : 8;
  if (wrp < RW) {
    constexpr int HALF = (RI_H + 1) / 2;
    for (int grp = wrp; grp < 8; grp += RW) {
      const int q = grp >> 1, frow0 = (grp & 1) * HALF, nrow = (grp & 1) ? RI_H - HALF : HALF;
      // (restrict: source arrays and R buffer are disjoint parts of shared memory; without it every store orders the loads of the
      //  next row behind it and the loop runs one shared-memory latency per row)
      const real* __restrict__ src = t.arr(A_U + q) + (lane + H) + (frow0 + 1) * PI;  // face (i0 + lane, j0 - 2 + frow0)
      real* __restrict__ dst = t.RB() + q * RQ + frow0 * RI_W + lane;
      real rv[HALF];
#pragma unroll
      for (int n = 0; n < HALF; ++n)
        if (n < nrow) rv[n] = rrow(src + n * PI, 1);
#pragma unroll
      for (int n = 0; n < HALF; ++n)
        if (n < nrow) dst[n * RI_W] = rv[n];
    }
  } else if (wrp == NW - 1) {
    for (int idx = lane; idx < 4 * RI_H; idx += 32) {
      const int q = idx / RI_H, frow = idx % RI_H;
      t.RB()[q * RQ + frow * RI_W + OI] = rrow(t.arr(A_U + q) + (OI + H) + (frow + 1) * PI, 1);
    }
  }
}

// first ghost layer of the sensor cells by linear extrapolation of the gradients (rhs/gradveloingh.F:1-19); divu and the
// vorticity are linear in the gradients, so extrapolating them is extrapolating the gradients
BC_HD void phase1b(const TileCtx& t, int tid) {
  const GridDesc& g = t.g;
  const real* S0 = t.X();
  const real* S1 = S0 + GW * GH_;
  for (int idx = tid; idx < GW * GH_; idx += NT) {
    const int a = idx % GW + (H - 1), b = idx / GW + (H - 1);
    const int ci = t.i0 - H + a, cj = t.j0 - H + b;
    int d = 0;
    if (cj >= 1 && cj <= g.jm) {
      if (ci == 0 && !(g.edges & 1)) d = 1;
      else if (ci == g.im + 1 && !(g.edges & 2)) d = -1;
    } else if (ci >= 1 && ci <= g.im) {
      if (cj == 0) d = GW;
      else if (cj == g.jm + 1) d = -GW;
    }
    if (d != 0) {
      const real divu = 2.0 * S0[idx + d] - S0[idx + 2 * d];
      const real vort = 2.0 * S1[idx + d] - S1[idx + 2 * d];
      const int k = a + b * PI;
      t.arr(A_DV)[k] = BC_LDG(t.vol + g.cidx(ci, cj)) * divu;
      t.arr(A_DU)[k] = ducros_ratio(divu, vort);
    }
  }
}

// R_q of the j-faces (columns i0-2 .. i0+33, rows j0 .. j0+8): thread (fcol = tid % 36, grp = tid / 36) owns quantity grp/2
// and five (even grp) or four (odd grp) consecutive face rows, sliding along j
BC_HD void phase_rj(const TileCtx& t, int tid) {
  // warp-aligned like the R_q of the i-faces: (quantity, half of the rows) x 32 columns per warp, the four remaining columns
  // (i0+30 .. i0+33) on the last warp
  const int wrp = tid >> 5, lane = tid & 31;
  constexpr int NW = NT / 32, RW = NW - 1 < 8 ? NW - 1 : 8;
  constexpr int HALF = (RJ_H + 1) / 2;
  if (wrp < RW) {
    for (int grp = wrp; grp < 8; grp += RW) {
      const int q = grp >> 1, frow0 = (grp & 1) * HALF, nrow = (grp & 1) ? RJ_H - HALF : HALF;
      const real* __restrict__ src = t.arr(A_U + q) + (lane + 1) + (frow0 + H) * PI;  // face (i0 - 2 + lane, j0 + frow0)
      real* __restrict__ dst = t.RB() + q * RQ + frow0 * RJ_W + lane;
      real col[HALF + 3];   // the column of the source array this thread slides along: all loads first
#pragma unroll
      for (int n = 0; n < HALF + 3; ++n)
        if (n < nrow + 3) col[n] = src[(n - 2) * PI];
#pragma unroll
      for (int n = 0; n < HALF; ++n)
        if (n < nrow) dst[n * RJ_W] = 9.0 * (col[n + 1] + col[n + 2]) - (col[n] + col[n + 3]);
    }
  } else if (wrp == NW - 1) {
    constexpr int XC = RJ_W - 32;   // 4
    for (int idx = lane; idx < 4 * RJ_H * XC; idx += 32) {
      const int q = idx / (RJ_H * XC), rem = idx % (RJ_H * XC), frow = rem / XC, fcol = 32 + rem % XC;
      const real* src = t.arr(A_U + q) + (fcol + 1) + (frow + H) * PI;
      t.RB()[q * RQ + frow * RJ_W + fcol] = 9.0 * (src[-PI] + src[0]) - (src[-2 * PI] + src[PI]);
    }
  }
}

// ---- one regular face (FACE_MAIN, compact o4 viscous gradients) ---------------------------------------------------
// s / sw: shared pointers of the face cell in the derived arrays / the w planes; rb: R buffer entry of this face for q = 0 (quantity stride RQ, cross
// stride RC); n / c: node- and cell-layout indices of the face cell in the global metric arrays.
// Face metrics, loaded from global memory BEFORE the phases that precede the face evaluation so that their latency is
// hidden: face normal and the eight dual-cell normals (flux_visqueux_o4_{i,j}.F) with the scalings folded in.
struct FaceGeom {
  double nxf, nyf;
  double nApx, nAmx, nApy, nAmy, nCpx, nCmx, nCpy, nCmy;
};
template <int DIR>
BC_HD FaceGeom load_geom(const TileCtx& t, int fi, int fj) {
  const GridDesc& g = t.g;
  const long long n = g.nidx(fi, fj), c = g.cidx(fi, fj);
  const long long a1 = DIR == 0 ? 1 : g.ldn, c1 = DIR == 0 ? g.ldn : 1;
  const double* nxA = t.nx + DIR * g.sn + n;
  const double* nyA = t.ny + DIR * g.sn + n;
  const double* nxC = t.nx + (1 - DIR) * g.sn + n;
  const double* nyC = t.ny + (1 - DIR) * g.sn + n;
  FaceGeom G;
  G.nxf = BC_LDG(nxA);
  G.nyf = BC_LDG(nyA);
  const double volf = BC_LDG(t.volf + DIR * g.sc + c);
  constexpr double ccross = (0.25 / 3.0) * 0.0625;
  const double sA = (0.5 / 24.0) * volf, sC = (0.5 * ccross) * volf;
  G.nApx = (BC_LDG(nxA + a1) + G.nxf) * sA;
  G.nAmx = -(BC_LDG(nxA - a1) + G.nxf) * sA;
  G.nApy = (BC_LDG(nyA + a1) + G.nyf) * sA;
  G.nAmy = -(BC_LDG(nyA - a1) + G.nyf) * sA;
  G.nCpx = (BC_LDG(nxC - a1 + c1) + BC_LDG(nxC + c1)) * sC;
  G.nCmx = -(BC_LDG(nxC - a1) + BC_LDG(nxC)) * sC;
  G.nCpy = (BC_LDG(nyC - a1 + c1) + BC_LDG(nyC + c1)) * sC;
  G.nCmy = -(BC_LDG(nyC - a1) + BC_LDG(nyC)) * sC;
  return G;
}

// s / sw: shared pointers of the face cell in the derived arrays / the w planes; rb: R buffer entry of this face for q = 0 (quantity stride RQ, cross
// stride RC)
template <int DIR>
BC_HD void face_fast(const TileCtx& t, const real* s, const real* sw, const real* rb, const FaceGeom& G, real (&hn)[5]) {
  constexpr int SA = DIR == 0 ? 1 : PI;
  constexpr int RC = DIR == 0 ? RI_W : 1;
  const SchemeConsts& cs = t.c;
#define RF_LD(A_, K_) s[(A_) * NC + (K_) * SA]
#define RF_LW(E_, K_) sw[(E_) * NC + (K_) * SA]
  const double nxf = G.nxf, nyf = G.nyf;
  const double nApx = G.nApx, nAmx = G.nAmx, nApy = G.nApy, nAmy = G.nAmy;
  const double nCpx = G.nCpx, nCmx = G.nCmx, nCpy = G.nCpy, nCmy = G.nCmy;

  // (order of the blocks chosen for register pressure: the viscous part first, its dual normals and gradients die before
  //  the convective accumulators become live)
  // ---- viscous flux, compact 4th order (flux_visqueux_o4_{i,j}.F) -------------------------------------------------
  real gx[4], gy[4], fv[4];  // gradients and face values of u, v, w, T
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const real q1 = RF_LD(A_U + q, 1), q0 = RF_LD(A_U + q, 0), qm1 = RF_LD(A_U + q, -1), qm2 = RF_LD(A_U + q, -2);
    const real Ap = 26.0 * q0 - (q1 + qm1), Am = 26.0 * qm1 - (q0 + qm2);
    const real* r = rb + q * RQ;
    const real rm2 = r[-2 * RC], rm1 = r[-RC], r0 = r[0], r1 = r[RC], r2 = r[2 * RC];
    const real Cm = 7.0 * (rm1 + r0) - (rm2 + r1), Cp = 7.0 * (r0 + r1) - (rm1 + r2);
    gx[q] = Ap * nApx + Am * nAmx + Cp * nCpx + Cm * nCmx;
    gy[q] = Ap * nApy + Am * nAmy + Cp * nCpy + Cm * nCmy;
    fv[q] = 0.0625 * r0;
  }
  const real mmu = 0.0625 * (9.0 * (RF_LD(A_MU, -1) + RF_LD(A_MU, 0)) - (RF_LD(A_MU, -2) + RF_LD(A_MU, 1)));
  constexpr double TWOTHIRD = 2.0 / 3.0;
  const real lambda = mmu * cs.cpprandtl;
  const real fvrou = TWOTHIRD * mmu * (2.0 * gx[0] - gy[1]);
  const real fvrov = mmu * (gy[0] + gx[1]);
  const real fvrow = mmu * gx[2];
  const real gvrov = TWOTHIRD * mmu * (2.0 * gy[1] - gx[0]);
  const real gvrow = mmu * gy[2];
  const real fvroe = lambda * gx[3] + fv[0] * fvrou + fv[1] * fvrov + fv[2] * fvrow;
  const real gvroe = lambda * gy[3] + fv[0] * fvrov + fv[1] * gvrov + fv[2] * gvrow;
  const real visc[5] = {rf_const(0.0), fvrou * nxf + fvrov * nyf, fvrov * nxf + gvrov * nyf, fvrow * nxf + gvrow * nyf,
                        fvroe * nxf + gvroe * nyf};

  // ---- Roe spectral radius (spectralradius_{i,j}.F) ---------------------------------------------------------------
  const double nx2 = nxf * nxf + nyf * nyf;
  real rspec;
  {
    const real sr = RF_LD(A_SR, 0), sl = RF_LD(A_SR, -1);
    const real inv = rf_rcp(sl + sr);
    const real rr = sl * inv, omrr = sr * inv;  // 1/(1+sqrt(rho_r/rho_l)) and its complement
    const real u = RF_LD(A_U, -1) * rr + RF_LD(A_U, 0) * omrr;
    const real v = RF_LD(A_V, -1) * rr + RF_LD(A_V, 0) * omrr;
    const real c2x = (cs.gam * cs.rgaz) * (RF_LD(A_T, -1) * rr + RF_LD(A_T, 0) * omrr);
    rspec = rf_abs(nxf * u + nyf * v) + rf_sqrt(c2x * nx2);
  }

  // ---- Jameson x Ducros x dilatation sensor (ducrosfordnc_{i,j}.F) -------------------------------------------------
  real eps2 = rf_const(0.0);
  if (cs.k2 != 0.0) {
    const real pm2 = RF_LD(A_P, -2), pm1 = RF_LD(A_P, -1), p0 = RF_LD(A_P, 0), pp1 = RF_LD(A_P, 1);
    const real a1_ = rf_abs(pm1 - 2.0 * p0 + pp1), b1_ = rf_abs(pm1 + 2.0 * p0 + pp1);
    const real a2_ = rf_abs(pm2 - 2.0 * pm1 + p0), b2_ = rf_abs(pm2 + 2.0 * pm1 + p0);
    const bool second = rf_val(a1_) * rf_val(b2_) < rf_val(a2_) * rf_val(b1_);  // k1 < k2: max() takes the second operand
    const real ks = (second ? a2_ : a1_) * rf_rcp(second ? b2_ : b1_);
    const real duc = rf_max(RF_LD(A_DU, 0), RF_LD(A_DU, -1));
    const double sn = ::sqrt(nx2);
    const real t0 = RF_LD(A_DV, 0), t1 = RF_LD(A_DV, -1);
    const real d0 = RF_LD(A_CS, 0) * sn + 1e-15, d1 = RF_LD(A_CS, -1) * sn + 1e-15;
    // x0 > x1: the dilatation switch is decreasing, max() is at the smaller argument
    const bool take1 = rf_val(t0) * rf_val(d1) > rf_val(t1) * rf_val(d0);
    const real xs = 2.5 + 10.0 * (take1 ? t1 : t0) * rf_rcp(take1 ? d1 : d0);
    const real dxm = rf_rcp(1.0 + rf_exp(rf_min(2.0 * xs, 700.0)));  // (1 - tanh x) / 2
    eps2 = cs.k2 * (ks * duc * dxm);
  }
  const real eps4 = rf_max(0.0, cs.k4 - eps2 * 12.0);

  // ---- convective flux + 5th-difference operand: one pass over cells -3 .. 2 -----------------------------------
  constexpr double denom = 1.0 / 60.0;
  constexpr double ck[6] = {denom, -8.0 * denom, 37.0 * denom, 37.0 * denom, -8.0 * denom, denom};
  constexpr double dk[6] = {-denom, 5.0 * denom, -10.0 * denom, 10.0 * denom, -5.0 * denom, denom};
  real fx[5], pr[5];
#pragma unroll
  for (int e = 0; e < 5; ++e) fx[e] = pr[e] = rf_const(0.0);
  real pbar = rf_const(0.0);
#pragma unroll
  for (int k = -3; k <= 2; ++k) {
    const real u = RF_LD(A_U, k), v = RF_LD(A_V, k), p = RF_LD(A_P, k);
    const real cv = ck[k + 3] * (u * nxf + v * nyf);
    pbar += ck[k + 3] * p;
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      const real we = RF_LW(e, k);
      fx[e] += cv * (e == 4 ? we + p : we);
      pr[e] += dk[k + 3] * we;
    }
  }
  fx[1] += pbar * nxf;
  fx[2] += pbar * nyf;

  // ---- assembly (dissipation_ducros_{i,j}.F, fluxnumassembly_{i,j}.F) ----------------------------------------------
  const real e2h = 0.5 * eps2;
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    const real diff = RF_LW(e, 0) - RF_LW(e, -1);
    hn[e] = fx[e] - rspec * (e2h * diff + eps4 * pr[e]) - visc[e];
  }
#undef RF_LD
#undef RF_LW
}

// ---- bulk-staged variant: metrics from shared memory ---------------------------------------------------------------------------
// The copies of one tile as a list of tensor boxes (the host emulation executes the same list): op 0 = w, 1 = vol, 2 = volf,
// 3 + 2 pl + p = the rows of parity class p (r = p, p + 2, ...) of node plane pl (0 nx0, 1 nx1, 2 ny0, 3 ny1).
struct BulkOp {
  int kind;            // 0 w (3-D box PI x PJ x 5), 1 vol (2-D box), 2 volf (3-D box), 3 node box of nx, 4 node box of ny
  int dst;             // offset in doubles from the start of shared memory
  int x, y;            // tensor coordinates of the box origin (always even x: the first byte of a box must be 16-byte aligned)
};
constexpr int NBULK = 11;
// node planes as rows of 2 ldn elements: double row and half of storage row s of plane k
BC_HD int node_t(const GridDesc& g, int k, int s) { return k * (g.jm + 2 * g.gh + 1) + s; }
BC_HD int node_x(const GridDesc& g, int i0, int t) { return (t & 1) * g.ldn + i0 + 1; }   // element of node column i0-1 inside its double row
BC_HD BulkOp bulk_op(const GridDesc& g, int i0, int j0, int op) {
  BulkOp o;
  o.kind = -1; o.dst = 0; o.x = 0; o.y = 0;
  if (op == 0) { o.kind = 0; o.x = i0 - 1; o.y = j0 - 1; }
  else if (op == 1) { o.kind = 1; o.dst = O_VOLBOX; o.x = i0 + 1; o.y = j0 + 1; }
  else if (op == 2) { o.kind = 2; o.dst = O_MET + M_VOLF; o.x = i0 + 1; o.y = j0 + 2; }
  else if (op < NBULK) {
    const int pl = (op - 3) >> 1, p = (op - 3) & 1, k = pl & 1;
    const int t = node_t(g, k, j0 + 1 + p);                        // first row of the class: storage row of node row j0-1+p
    o.kind = 3 + (pl >> 1);
    o.x = node_x(g, i0, t) & ~1;
    o.y = t >> 1;
    o.dst = O_MET + M_NODE + (op - 3) * MN_BOX;
  }
  return o;
}
// value of node plane (isy ? ny : nx)[k] at window coordinates (a, r) = node (i0-1+a, j0-1+r)
struct NodeView {
  const double* m;   // met + M_NODE
  int nsh;           // bit 2 k + p: shift of the rows of plane k, parity class p
  BC_HD const double* row(int isy, int k, int r) const {
    const int c = 2 * k + (r & 1);
    return m + (4 * isy + c) * MN_BOX + (r >> 1) * MN_SLOT + ((nsh >> c) & 1);
  }
};
BC_HD int node_shift_mask(const GridDesc& g, int i0, int j0) {   // once per tile
  int mask = 0;
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int p = 0; p < 2; ++p) mask |= (node_x(g, i0, node_t(g, k, j0 + 1 + p)) & 1) << (2 * k + p);
  return mask;
}
BC_HD NodeView node_view(const TileCtx& t) { return NodeView{t.met + M_NODE, t.nsh}; }
template <int DIR>
BC_HD FaceGeom load_geom_sm(const TileCtx& t, int fi, int fj) {
  const NodeView nv = node_view(t);
  const int a = fi - (t.i0 - 1), r = fj - (t.j0 - 1);
  const double volf = t.met[M_VOLF + DIR * MF_PS + (fi - t.i0 + 1) + (fj - t.j0) * MF_W];
  constexpr double ccross = (0.25 / 3.0) * 0.0625;
  const double sA = (0.5 / 24.0) * volf, sC = (0.5 * ccross) * volf;
  FaceGeom G;
  if (DIR == 0) {
    const double* ax = nv.row(0, 0, r) + a;
    const double* ay = nv.row(1, 0, r) + a;
    const double *cx0 = nv.row(0, 1, r) + a, *cx1 = nv.row(0, 1, r + 1) + a;
    const double *cy0 = nv.row(1, 1, r) + a, *cy1 = nv.row(1, 1, r + 1) + a;
    G.nxf = ax[0]; G.nyf = ay[0];
    G.nApx = (ax[1] + G.nxf) * sA; G.nAmx = -(ax[-1] + G.nxf) * sA;
    G.nApy = (ay[1] + G.nyf) * sA; G.nAmy = -(ay[-1] + G.nyf) * sA;
    G.nCpx = (cx1[-1] + cx1[0]) * sC; G.nCmx = -(cx0[-1] + cx0[0]) * sC;
    G.nCpy = (cy1[-1] + cy1[0]) * sC; G.nCmy = -(cy0[-1] + cy0[0]) * sC;
  } else {
    const double *axm = nv.row(0, 1, r - 1) + a, *ax0 = nv.row(0, 1, r) + a, *axp = nv.row(0, 1, r + 1) + a;
    const double *aym = nv.row(1, 1, r - 1) + a, *ay0 = nv.row(1, 1, r) + a, *ayp = nv.row(1, 1, r + 1) + a;
    const double *cxm = nv.row(0, 0, r - 1) + a, *cx0 = nv.row(0, 0, r) + a;
    const double *cym = nv.row(1, 0, r - 1) + a, *cy0 = nv.row(1, 0, r) + a;
    G.nxf = ax0[0]; G.nyf = ay0[0];
    G.nApx = (axp[0] + G.nxf) * sA; G.nAmx = -(axm[0] + G.nxf) * sA;
    G.nApy = (ayp[0] + G.nyf) * sA; G.nAmy = -(aym[0] + G.nyf) * sA;
    G.nCpx = (cxm[1] + cx0[1]) * sC; G.nCmx = -(cxm[0] + cx0[0]) * sC;
    G.nCpy = (cym[1] + cy0[1]) * sC; G.nCmy = -(cym[0] + cy0[0]) * sC;
  }
  return G;
}
BC_HD SensGeom sensor_geom_sm(const TileCtx& t, int tid, int round) {
  SensGeom G{};
  int ga, gb;
  G.valid = sensor_of(t, tid, round, ga, gb);
  if (G.valid) {   // window coordinates of the sensor window = those of the vol box and of the node rows
    const NodeView nv = node_view(t);
    G.vol = t.volbox[ga + gb * MV_W];
    const double volm1 = frcp(G.vol);
    const double* x0 = nv.row(0, 0, gb) + ga;
    const double* y0 = nv.row(1, 0, gb) + ga;
    G.dxm1 = 0.5 * (x0[0] + x0[1]) * volm1;
    G.dxm2 = 0.5 * (nv.row(0, 1, gb)[ga] + nv.row(0, 1, gb + 1)[ga]) * volm1;
    G.dym1 = 0.5 * (y0[0] + y0[1]) * volm1;
    G.dym2 = 0.5 * (nv.row(1, 1, gb)[ga] + nv.row(1, 1, gb + 1)[ga]) * volm1;
  }
  return G;
}

// ---- phase 2: i-faces (i0 + col, j0 + row) -------------------------------------------------------------------------
// warps 0..7: row = warp, col = lane (the left faces of the tile's cells); warp 8, lanes 0..7: the right faces of the
// last column (col = 32, row = lane).  Both mappings are bank-conflict free on the 38-double pitch.
struct FaceId {
  int row, col, fi, fj;
  bool active, generic;
};
BC_HD FaceId iface_of(const TileCtx& t, int tid) {
  FaceId f;
  if (tid < OI * OJ) { f.row = tid / OI; f.col = tid % OI; }
  else { f.row = tid - OI * OJ; f.col = OI; }
  f.fi = t.i0 + f.col;
  f.fj = t.j0 + f.row;
  f.active = f.row < OJ && f.fi <= t.g.im + 1 && f.fj <= t.g.jm;
  f.generic = t.wall && f.fj <= 2;
  return f;
}
BC_HD FaceId jface_of(const TileCtx& t, int tid) {
  FaceId f;
  f.row = tid / OI; f.col = tid % OI;
  f.fi = t.i0 + f.col;
  f.fj = t.j0 + f.row;
  f.active = f.fi <= t.g.im && f.fj <= t.g.jm + 1;
  f.generic = t.wall && f.fj <= 3;
  return f;
}
BC_HD FaceGeom prefetch_iface(const TileCtx& t, int tid) {
  const FaceId f = iface_of(t, tid);
  if (f.active && !f.generic) return load_geom<0>(t, f.fi, f.fj);
  return FaceGeom{};
}
BC_HD FaceGeom prefetch_jface(const TileCtx& t, int tid) {
  const FaceId f = jface_of(t, tid);
  if (f.active && !f.generic) return load_geom<1>(t, f.fi, f.fj);
  return FaceGeom{};
}

BC_HD FaceGeom geom_iface_sm(const TileCtx& t, int tid) {
  const FaceId f = iface_of(t, tid);
  if (f.active && !f.generic) return load_geom_sm<0>(t, f.fi, f.fj);
  return FaceGeom{};
}
BC_HD FaceGeom geom_jface_sm(const TileCtx& t, int tid) {
  const FaceId f = jface_of(t, tid);
  if (f.active && !f.generic) return load_geom_sm<1>(t, f.fi, f.fj);
  return FaceGeom{};
}

// tangent build: does any cell of the face's stencil carry a tangent?  (along -3 .. 2 on its own row -- up to +3 for the
// off-centred wall flux of face j = 2 --, along -2 .. 1 on the cross rows +-1, +-2; the tangents of the two sensor cells cover
// the extrapolated gradients of the first ghost layer)
template <int DIR>
BC_HD bool face_has_tangent(const TileCtx& t, int a, int b, int hi) {
#ifdef BCAST_RF_DUAL
  if (!t.flags) return true;
  constexpr int SA = DIR == 0 ? 1 : PI, SC = DIR == 0 ? PI : 1;
  const unsigned char* f = t.flags + a + b * PI;
  bool act = false;
  for (int s_ = -3; s_ <= hi; ++s_) act = act || f[s_ * SA] != 0;
  for (int t_ = -2; t_ <= 2; ++t_)
    for (int s_ = -2; s_ <= 1; ++s_) act = act || f[s_ * SA + t_ * SC] != 0;
  const int k = a + b * PI;
  act = act || t.arr(A_DV)[k].d != 0.0 || t.arr(A_DU)[k].d != 0.0 || t.arr(A_DV)[k - SA].d != 0.0 || t.arr(A_DU)[k - SA].d != 0.0;
  return act;
#else
  return true;
#endif
}

BC_HD void phase2(const TileCtx& t, int tid, const FaceGeom& G) {
  const FaceId f = iface_of(t, tid);
  if (!f.active) return;
  real hn[5];
  const int a = f.col + H, b = f.row + H;
  if (!face_has_tangent<0>(t, a, b, 2)) {
#pragma unroll
    for (int e = 0; e < 5; ++e) hn[e] = rf_const(0.0);
  } else if (f.generic) {
    Var<RfDT> h[5];
    face_flux<0, true, FACE_MAIN>(make_acc(t, a, b), t.c, h);
#pragma unroll
    for (int e = 0; e < 5; ++e) hn[e] = rf_from_var(h[e]);
  } else {
    face_fast<0>(t, t.sm + a + b * PI, t.wsm + a + b * PI, t.RB() + (f.row + 2) * RI_W + f.col, G, hn);
  }
  real* X = t.X();
#pragma unroll
  for (int e = 0; e < 5; ++e) X[(e * OJ + f.row) * XI_P + f.col] = hn[e];
}

// ---- phase 3: j-faces ----------------------------------------------------------------------------------------------
BC_HD void phase3(const TileCtx& t, int tid, const FaceGeom& G) {
  const FaceId f = jface_of(t, tid);
  if (!f.active) return;
  real hn[5];
  const int a = f.col + H, b = f.row + H;
  if (!face_has_tangent<1>(t, a, b, (t.wall && f.fj == 2) ? 3 : 2)) {
#pragma unroll
    for (int e = 0; e < 5; ++e) hn[e] = rf_const(0.0);
  } else if (f.generic) {
    Var<RfDT> h[5];
    const SmemAcc2 A = make_acc(t, a, b);
    if (f.fj == 1) face_flux<1, true, FACE_WALL>(A, t.c, h);
    else if (f.fj == 2) face_flux<1, true, FACE_NEAR3>(A, t.c, h);
    else face_flux<1, false, FACE_NEAR5>(A, t.c, h);
#pragma unroll
    for (int e = 0; e < 5; ++e) hn[e] = rf_from_var(h[e]);
  } else {
    face_fast<1>(t, t.sm + a + b * PI, t.wsm + a + b * PI, t.RB() + f.row * RJ_W + f.col + 2, G, hn);
  }
  real* X = t.X();
#pragma unroll
  for (int e = 0; e < 5; ++e) X[(e * (OJ + 1) + f.row) * XJ_P + f.col] = hn[e];
}

// ---- balance (rhs/balance.F:2-15) ------------------------------------------------------------------------------------
BC_HD bool owns_cell(const TileCtx& t, int tid) {
  const int i = t.i0 + tid % OI, j = t.j0 + tid / OI;
  return tid < OI * OJ && i <= t.g.im && j <= t.g.jm && i <= t.i1 && j <= t.j1;
}
BC_HD void balance_i(const TileCtx& t, int tid, real (&r)[5]) {
  if (!owns_cell(t, tid)) return;
  const int cx = tid % OI, cy = tid / OI;
  const real* X = t.X();
#pragma unroll
  for (int e = 0; e < 5; ++e) r[e] = -(X[(e * OJ + cy) * XI_P + cx + 1] - X[(e * OJ + cy) * XI_P + cx]);
}
BC_HD void balance_j_store(const TileCtx& t, int tid, const real (&r)[5]) {
  if (!owns_cell(t, tid)) return;
  const int cx = tid % OI, cy = tid / OI;
  const real* X = t.X();
  const long long k = t.g.cidx(t.i0 + cx, t.j0 + cy);
#pragma unroll
  for (int e = 0; e < 5; ++e)
    t.res[e * t.g.sc + k] = rf_out(r[e] - (X[(e * (OJ + 1) + cy + 1) * XJ_P + cx] - X[(e * (OJ + 1) + cy) * XJ_P + cx]));
}

}  // namespace BCAST_RF_NS
}  // namespace bcast
