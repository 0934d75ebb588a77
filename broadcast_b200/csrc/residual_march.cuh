// Third-generation fused primal residual (order 5, gh = 3): a persistent CTA MARCHES along j over a strip of 32 columns.
//
// Why (ncu of k_residual_fast, profiles/r1_i_summary.md): the 32 x 9 tile kernel stages 38 x 15 cells per 32 x 9 outputs (1.98
// staged cells per output cell), every CTA starts with exposed HBM loads of w and of ~30 metric values per thread, and 57 % of its
// stall samples sit in those staging phases.  Here (2.5-D blocking):
//   * a work item is (strip of 32 columns) x (segment of SEG rows); the CTA advances RB = 4 rows per step and keeps the rows it
//     still needs in shared-memory RINGS, so a cell is staged 38/32 = 1.19 times and its primitives are computed once;
//   * the rows of the NEXT step are in flight while the faces of the current step are evaluated, all of them by the bulk-copy
//     engine with completion on an mbarrier: the five planes of w, vol and volf as TMA tensor boxes (38 x 1 x planes per row);
//     the node-layout planes nx, ny (odd leading dimension: global strides must be multiples of 16 bytes for a tensor map) as
//     1-D bulk copies of one row each, started at the 16-byte aligned element below the row's first element -- the row then sits
//     shifted by `shift` in {0, 1} entries in its slot and readers add the shift.  No face waits on global memory, no metric sits
//     in registers across a phase, no thread computes a global address in the steady state except the 32 lanes that issue the copies;
//   * i-faces, j-faces and the four right-edge i-faces of a step run through ONE copy of the face code (runtime shared-memory
//     offsets instead of a template parameter: the three inlined copies of the first version of this kernel thrashed the
//     instruction cache, 41 % of the face-phase stall samples were `no_instruction`, profiles/r2_a_summary.md);
//   * three barriers per step: [sensors, R_i, R_j of the new rows] | [all faces] | [balance + store, primitives of the next rows].
// The face formulas are those of residual_fast.cuh (same re-associations); the three wall rows keep the reference-shaped templates
// of scheme.cuh behind a call that is not inlined.
//
// Geometry of an item: output columns i0 .. i0+31, rows j0 .. j1.  Staged columns a = 0 .. 37 <-> cell i0-3+a.  Rows are
// addressed by q = j - (j0-3) (q = 0 is the lowest staged row); ring position = q mod ring depth.
//   cells   ring 13 rows, row block [w0..w4, u, v, wz, T, p, mu, sqrt(rho), c][38]  (the TMA box of w lands on arrays 0..4)
//   metrics ring 10 rows, row block vol | volf0 volf1 | nx0 nx1 ny0 ny1 (slots of 40: shifted rows)   (copy destinations aligned)
//   sensor  ring  8 rows [vol*divu, ducros, divu, vort][34]   columns i0-1 .. i0+32
//   R       per quantity u, v, w, T: R_i ring 8 rows x 34 (i-faces i0 .. i0+32) followed by R_j 4 rows x 36 (columns i0-2 .. i0+33)
//   X       per equation: i-face fluxes 4 x 33 followed by the j-face flux ring 8 x 32
// Step s (output rows J = j0+4s .. J+3, qJ = 3+4s):  sensors J+1 .. J+4, R_i J+2 .. J+5, R_j faces J+1 .. J+4 | i-faces rows
// J .. J+3, j-faces J+1 .. J+4 | balance + store of rows J .. J+3, primitives of rows J+7 .. J+10.  Copies of rows J+7 .. J+10
// (cells) and J+6 .. J+9 (metrics) are issued at the top of the step.  A prologue provides rows j0-3 .. j0+6 and the bottom faces.
// Reference: srcfv/rhs/flux_num_dnc5.F90:7-226.
#pragma once
#include "scheme.cuh"

#if defined(__CUDACC__)
#define RM_COLD __host__ __device__ __noinline__
#else
#define RM_COLD static __attribute__((noinline))
#endif

namespace bcast {
namespace rm {

constexpr int H = 3, W = 32, RB = 4, PC = W + 2 * H;   // 38 staged columns
constexpr int NT = 288;                                 // 9 compute warps (the kernel adds a tenth: the copy-issuing warp)
constexpr int NT_LAUNCH = NT + 32;
// cell ring
constexpr int NRC = 13, CROW = 496;                     // 13 arrays x 38 = 494, padded to a multiple of 128 bytes
enum { C_W = 0, C_U = 5, C_V = 6, C_WZ = 7, C_T = 8, C_P = 9, C_MU = 10, C_SR = 11, C_CS = 12 };
// metric ring: vol (38 -> 48) | volf [2][38] (-> 80) | four node planes in slots of 40 (a row may sit shifted by one entry)
constexpr int NRM = 10, MROW = 288, MSLOT = 40;
enum { M_VOL = 0, M_VF0 = 48, M_VF1 = 86, M_NX0 = 128, M_NX1 = 168, M_NY0 = 208, M_NY1 = 248, M_NYO = 80 /* ny - nx */ };
// sensor ring
constexpr int SW = W + 2, NRS = 8, SARR = NRS * SW;
enum { S_DV = 0, S_DU = 1, S_DIV = 2, S_VORT = 3 };
// R buffer: per quantity [R_i: 8 x 34][R_j: 4 x 36]
constexpr int RIW = 34, NRI = 8, RJW = W + 4, NRJ = 4;
constexpr int R_J = NRI * RIW, RQS = R_J + NRJ * RJW;   // offset of R_j inside a quantity slab, slab size
// flux exchange: per equation [X_i: 4 x 33][X_j ring: 8 x 32]
constexpr int XIW = 33, NXJ = 8;
constexpr int X_J = RB * XIW, XES = X_J + NXJ * W;
// shared-memory map (doubles)
constexpr int O_CELLS = 0;
constexpr int O_MET = O_CELLS + NRC * CROW;
constexpr int O_SENS = O_MET + NRM * MROW;
constexpr int O_R = O_SENS + 4 * SARR;
constexpr int O_X = O_R + 4 * RQS;
constexpr int O_BAR = O_X + 5 * XES;
constexpr int NSM = O_BAR + 2;
static_assert((CROW * 8) % 128 == 0 && (MROW * 8) % 128 == 0 && (O_MET * 8) % 128 == 0 && (M_VF0 * 8) % 128 == 0, "TMA destinations");
static_assert((M_NX0 * 8) % 16 == 0 && (MSLOT * 8) % 16 == 0, "bulk-copy destinations");
static_assert(M_NY0 - M_NX0 == M_NYO && M_NY1 - M_NX1 == M_NYO, "ny follows nx");

BC_HD double frcp(double x) {   // reciprocal: MUFU seed + two Newton steps (device); plain division on the host build
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

struct Item {
  int i0, j0, j1;   // first output column, first / last output row
};

struct MCtx {
  double* sm;
  const GridDesc& g;
  const SchemeConsts& c;
  double sqgr;
  bool wall;
  const double *w, *nx, *ny, *vol, *volf;
  double* res;
  Item it;
  BC_HD MCtx(const GridDesc& g_, const SchemeConsts& c_) : g(g_), c(c_) {}
  BC_HD int jrow(int q) const { return it.j0 - H + q; }   // Fortran row of staged row q
  BC_HD int qlast() const { return it.j1 - it.j0 + 2 * H; }   // last staged row any face of the item reads (row j1+3)
  BC_HD bool row_exists(int q) const { return q <= qlast() && jrow(q) <= g.jm + g.gh; }
  // node planes: first element (in its array) of staged row q of plane k (0: i-face normals, 1: j-face normals); a row is copied
  // from the 16-byte aligned element at or below it, so it sits shifted by the parity of that index
  BC_HD long long node_first(int q, int k) const { return (long long)(it.i0 - 1) + (long long)(jrow(q) - 1 + g.gh) * g.ldn + (long long)k * g.sn; }
  BC_HD int shift(int q, int k) const {   // == node_first(q, k) & 1, in 32-bit parity arithmetic
    return ((it.i0 - 1) + ((jrow(q) - 1 + g.gh) & g.ldn) + (k & (int)g.sn)) & 1;
  }
};
BC_HD int cpos(int q) { return O_CELLS + (q % NRC) * CROW; }
BC_HD int mpos(int q) { return O_MET + (q % NRM) * MROW; }
// ring positions of consecutive rows without a division each: next position / position + r (0 <= r < n)
BC_HD int ring_inc(int p, int n) { return p + 1 == n ? 0 : p + 1; }
BC_HD int ring_add(int p, int r, int n) { p += r; return p >= n ? p - n : p; }
// metric row block of row q with the shift of node plane k folded in: nx_k(a) = sm[mnode(t, q, k) + M_NX0 + k * MSLOT + a]
BC_HD int mnode(const MCtx& t, int q, int k) { return mpos(q) + t.shift(q, k); }

// ---- the asynchronous copies of one generation, as a flat list of operations (one per lane and round) -----------------------
// op < cn: w of cell row cq0 + op | then vol, volf of the mn metric rows | then the 4 node planes of each metric row
struct CopyOp {
  int kind;          // 0 w (3-D box), 1 vol (2-D box), 2 volf (3-D box), 3 node row (1-D bulk copy), -1 none
  int dst;           // offset into sm
  int x, y;          // tensor coordinates (kinds 0-2)
  const double* src; // kind 3
  int bytes;
};
BC_HD int copy_count(int cn, int mn) { return cn + 6 * mn; }
BC_HD CopyOp copy_op(const MCtx& t, int op, int cq0, int cn, int mq0, int mn) {
  CopyOp o;
  o.kind = -1; o.dst = 0; o.x = t.it.i0 - 1; o.y = 0; o.src = nullptr; o.bytes = 0;
  if (op < cn) {
    const int q = cq0 + op;
    if (!t.row_exists(q)) return o;
    o.kind = 0; o.dst = cpos(q); o.y = t.jrow(q) - 1 + t.g.gh; o.bytes = 5 * PC * 8;
    return o;
  }
  op -= cn;
  if (op < 2 * mn) {
    const int q = mq0 + (op >> 1);
    if (!t.row_exists(q)) return o;
    o.y = t.jrow(q) - 1 + t.g.gh;
    if (op & 1) { o.kind = 2; o.dst = mpos(q) + M_VF0; o.bytes = 2 * PC * 8; }
    else { o.kind = 1; o.dst = mpos(q) + M_VOL; o.bytes = PC * 8; }
    return o;
  }
  op -= 2 * mn;
  if (op < 4 * mn) {
    const int q = mq0 + (op >> 2), pl = op & 3, k = pl & 1;
    if (!t.row_exists(q)) return o;
    const long long first = t.node_first(q, k), lo = first & ~1LL;
    long long cnt = PC + (first - lo);
    cnt = (cnt + 1) & ~1LL;                       // 38 or 40 elements
    const long long total = 2 * t.g.sn;
    if (lo + cnt > total) cnt = total - lo;       // last row of the array (even: total and lo are)
    o.kind = 3;
    o.src = ((pl & 2) ? t.ny : t.nx) + lo;
    o.dst = mpos(q) + M_NX0 + pl * MSLOT;
    o.bytes = (int)cnt * 8;
  }
  return o;
}

// ---- prims of staged rows q0 .. q0+n-1 (phys/Primitives.F:2-34, phys/viscosity.F:1); tasks task0, task0 + stride, ... ------
BC_HD void prims_rows(const MCtx& t, int task0, int stride, int q0, int n) {
  for (int task = task0; task < n * PC; task += stride) {
    const int r = task / PC, a = task - r * PC;
    const int q = q0 + r;
    if (!t.row_exists(q)) continue;
    double* cr = t.sm + cpos(q) + a;
    double q0_ = cr[(C_W + 0) * PC], q1 = cr[(C_W + 1) * PC], q2 = cr[(C_W + 2) * PC], q3 = cr[(C_W + 3) * PC], q4 = cr[(C_W + 4) * PC];
    if (t.it.i0 - H + a > t.g.im + t.g.gh) {   // beyond the padded array: a sane state (never used by an active face)
      q0_ = 1.0; q1 = 0.0; q2 = 0.0; q3 = 0.0; q4 = 1.0;
    }
    const double rom1 = frcp(q0_);
    const double u = q1 * rom1, v = q2 * rom1, wz = q3 * rom1;
    const double ec = 0.5 * (u * u + v * v + wz * wz);
    const double eloc = (q4 - ec * q0_) * rom1;
    const double tl = eloc * t.c.cvm1;
    const double p = t.c.gam1 * q0_ * eloc;
    const double sqt = ::sqrt(tl);
    cr[C_U * PC] = u;
    cr[C_V * PC] = v;
    cr[C_WZ * PC] = wz;
    cr[C_T * PC] = tl;
    cr[C_P * PC] = p;
    cr[C_MU * PC] = t.c.betas * frcp(tl + t.c.s_suth) * sqt * tl;
    cr[C_SR * PC] = ::sqrt(q0_);
    cr[C_CS * PC] = t.sqgr * sqt;
  }
}

// ---- sensor cells, R_i, R_j -----------------------------------------------------------------------------------------------------
BC_HD double ducros_ratio(double divu, double vort) {
  const double d2 = divu * divu;
  return d2 * frcp(d2 + vort * vort + 1e-15);
}
// dilatation, vorticity and Ducros ratio of sensor cell (ga, q): gradop_5pi.F, gradop_5pj.F, gradient.F, geom/dxdy.F
BC_HD void sensor_cell(const MCtx& t, int q, int ga) {
  const GridDesc& g = t.g;
  const int ci = t.it.i0 - 1 + ga, cj = t.jrow(q);
  if (!(ci >= g.glo() && ci <= g.ghi() && cj >= 1 && cj <= g.jm)) return;   // slab-internal edges: real gradients in the halo column
  const int a = ga + H - 1;
  const double* sm = t.sm;
  const int pm = q % NRM;
  const int m = O_MET + pm * MROW + a, mq1 = O_MET + ring_inc(pm, NRM) * MROW + a;
  const int n0 = m + t.shift(q, 0), n1 = m + t.shift(q, 1) + M_NX1, n1p = mq1 + t.shift(q + 1, 1) + M_NX1;
  const double vol = sm[m + M_VOL];
  const double volm1 = 1.0 / vol;
  const double dxm1 = 0.5 * (sm[n0 + M_NX0] + sm[n0 + M_NX0 + 1]) * volm1;
  const double dxm2 = 0.5 * (sm[n1] + sm[n1p]) * volm1;
  const double dym1 = 0.5 * (sm[n0 + M_NY0] + sm[n0 + M_NY0 + 1]) * volm1;
  const double dym2 = 0.5 * (sm[n1 + M_NYO] + sm[n1p + M_NYO]) * volm1;
  int pc = (q - 2) % NRC;
  const double* cm2 = sm + O_CELLS + pc * CROW + a;
  pc = ring_inc(pc, NRC);
  const double* cm1 = sm + O_CELLS + pc * CROW + a;
  pc = ring_inc(pc, NRC);
  const double* c0 = sm + O_CELLS + pc * CROW + a;
  pc = ring_inc(pc, NRC);
  const double* cp1 = sm + O_CELLS + pc * CROW + a;
  pc = ring_inc(pc, NRC);
  const double* cp2 = sm + O_CELLS + pc * CROW + a;
  constexpr double b1 = 8.0 * (1.0 / 12.0), b2 = -(1.0 / 12.0);
  constexpr int U = C_U * PC, V = C_V * PC;
  const double gui = b1 * (c0[U + 1] - c0[U - 1]) + b2 * (c0[U + 2] - c0[U - 2]);
  const double gvi = b1 * (c0[V + 1] - c0[V - 1]) + b2 * (c0[V + 2] - c0[V - 2]);
  const double guj = b1 * (cp1[U] - cm1[U]) + b2 * (cp2[U] - cm2[U]);
  const double gvj = b1 * (cp1[V] - cm1[V]) + b2 * (cp2[V] - cm2[V]);
  const double gu0 = dxm1 * gui + dxm2 * guj, gv0 = dxm1 * gvi + dxm2 * gvj;
  const double gu1 = dym1 * gui + dym2 * guj, gv1 = dym1 * gvi + dym2 * gvj;
  const double divu = gu0 + gv1, vort = gv0 - gu1;
  double* s = t.sm + O_SENS + (q & (NRS - 1)) * SW + ga;
  s[S_DIV * SARR] = divu;
  s[S_VORT * SARR] = vort;
  s[S_DV * SARR] = vol * divu;
  s[S_DU * SARR] = ducros_ratio(divu, vort);
}
// first ghost layer of the sensor cells by linear extrapolation of the gradients (rhs/gradveloingh.F:1-19); divu and the
// vorticity are linear in the gradients.  Corners are never read.
BC_HD void sensor_ghost(const MCtx& t, int q, int ga) {
  const GridDesc& g = t.g;
  const int ci = t.it.i0 - 1 + ga, cj = t.jrow(q);
  int q1 = q, q2 = q, da = 0;
  if (cj >= 1 && cj <= g.jm) {
    if (ci == 0 && !(g.edges & 1)) da = 1;
    else if (ci == g.im + 1 && !(g.edges & 2)) da = -1;
    else return;
  } else if (ci >= 1 && ci <= g.im) {
    if (cj == 0) { q1 = q + 1; q2 = q + 2; }
    else if (cj == g.jm + 1) { q1 = q - 1; q2 = q - 2; }
    else return;
  } else {
    return;
  }
  const double* s1 = t.sm + O_SENS + (q1 & (NRS - 1)) * SW + ga + da;
  const double* s2 = t.sm + O_SENS + (q2 & (NRS - 1)) * SW + ga + 2 * da;
  const double divu = 2.0 * s1[S_DIV * SARR] - s2[S_DIV * SARR];
  const double vort = 2.0 * s1[S_VORT * SARR] - s2[S_VORT * SARR];
  double* s = t.sm + O_SENS + (q & (NRS - 1)) * SW + ga;
  s[S_DV * SARR] = t.sm[mpos(q) + M_VOL + ga + H - 1] * divu;
  s[S_DU * SARR] = ducros_ratio(divu, vort);
}
BC_HD bool item_has_ghost_sensor(const MCtx& t, int q0, int n) {   // CTA-uniform
  const GridDesc& g = t.g;
  if ((t.it.i0 == 1 && !(g.edges & 1)) || (t.it.i0 + W >= g.im + 1 && !(g.edges & 2))) return true;
  const int ja = t.jrow(q0), jb = t.jrow(q0 + n - 1);
  return (ja <= 0 && jb >= 0) || (ja <= g.jm + 1 && jb >= g.jm + 1);
}
// R_q(face) = -q(-2) + 9 q(-1) + 9 q(0) - q(1) along the face normal (the 1/16 is applied by the consumer)
BC_HD void ri_face(const MCtx& t, int q, int fc) {   // i-face i0+fc of row q
  if (!t.row_exists(q)) return;
  const double* c0 = t.sm + cpos(q) + fc + H;
  double* dst = t.sm + O_R + (q & (NRI - 1)) * RIW + fc;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double* s = c0 + (C_U + k) * PC;
    dst[k * RQS] = 9.0 * (s[-1] + s[0]) - (s[-2] + s[1]);
  }
}
BC_HD void rj_face(const MCtx& t, int qf, int fc) {   // j-face of staged row qf (between rows qf-1 and qf), column i0-2+fc
  if (!t.row_exists(qf + 1)) return;
  const int a = fc + 1;
  int pc = (qf - 2) % NRC;
  const double* cm2 = t.sm + O_CELLS + pc * CROW + a;
  pc = ring_inc(pc, NRC);
  const double* cm1 = t.sm + O_CELLS + pc * CROW + a;
  pc = ring_inc(pc, NRC);
  const double* c0 = t.sm + O_CELLS + pc * CROW + a;
  pc = ring_inc(pc, NRC);
  const double* cp1 = t.sm + O_CELLS + pc * CROW + a;
  double* dst = t.sm + O_R + R_J + (qf & (NRJ - 1)) * RJW + fc;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int o = (C_U + k) * PC;
    dst[k * RQS] = 9.0 * (cm1[o] + c0[o]) - (cm2[o] + cp1[o]);
  }
}
// one phase: sensor rows [sq0, sq0+sn), R_i rows [iq0, iq0+in), R_j face rows [jq0, jq0+jn)
constexpr int NT_SENS = RB * SW;   // 136 threads on sensor cells, the other 152 on the R tasks
BC_HD void phase_sens_r(const MCtx& t, int tid, int sq0, int sn, int iq0, int in, int jq0, int jn) {
  if (tid < NT_SENS) {
    for (int task = tid; task < sn * SW; task += NT_SENS) {
      const int r = task / SW;
      sensor_cell(t, sq0 + r, task - r * SW);
    }
  } else {
    const int nri = in * XIW, nrj = jn * RJW;
    for (int task = tid - NT_SENS; task < nri + nrj; task += NT - NT_SENS) {
      if (task < nri) {
        const int r = task / XIW;
        ri_face(t, iq0 + r, task - r * XIW);
      } else {
        const int u = task - nri;
        const int r = u / RJW;
        rj_face(t, jq0 + r, u - r * RJW);
      }
    }
  }
}
BC_HD void phase_sens_ghost(const MCtx& t, int tid, int sq0, int sn) {
  for (int task = tid; task < sn * SW; task += NT) {
    const int r = task / SW;
    sensor_ghost(t, sq0 + r, task - r * SW);
  }
}

// ---- one regular face (compact o4 viscous gradients): the formulas of rf::face_fast on the ring layout -----------------------
// Everything is addressed by offsets into sm, so that i-faces and j-faces share one instruction stream.
struct FaceArgs {
  int rc[6];   // the face cell's column (array 0) in the cell row blocks of along-normal offsets -3 .. 2
  int rp[5];   // R entries (quantity u) of cross offsets -2 .. 2
  int s0, s1;  // sensor entries (S_DV) of the face cell and of its -1 neighbour
  int gA0, gAp, gAm;          // nx of the face, of the next / previous face along the normal   (ny at + M_NYO)
  int gC0, gCp, gCm, gCmp;    // nx of the cross faces: (0), (+c1), (-a1), (-a1 + c1)
  int gvf;     // volf of the face
  int xo;      // X entry of equation 0 (equation stride XES)
};
BC_HD void face_eval(const SchemeConsts& cs, double* sm, const FaceArgs& A) {
#define RM_LD(A_, K_) sm[A.rc[(K_) + 3] + (A_) * PC]
  const double nxf = sm[A.gA0], nyf = sm[A.gA0 + M_NYO];
  // ---- viscous flux, compact 4th order (flux_visqueux_o4_{i,j}.F); metric scalings folded into the eight dual-cell normals ----
  double gx[4], gy[4], fv[4];
  {
    const double volf = sm[A.gvf];
    constexpr double ccross = (0.25 / 3.0) * 0.0625;
    const double sA = (0.5 / 24.0) * volf, sC = (0.5 * ccross) * volf;
    const double nApx = (sm[A.gAp] + nxf) * sA, nAmx = -(sm[A.gAm] + nxf) * sA;
    const double nApy = (sm[A.gAp + M_NYO] + nyf) * sA, nAmy = -(sm[A.gAm + M_NYO] + nyf) * sA;
    const double nCpx = (sm[A.gCmp] + sm[A.gCp]) * sC, nCmx = -(sm[A.gCm] + sm[A.gC0]) * sC;
    const double nCpy = (sm[A.gCmp + M_NYO] + sm[A.gCp + M_NYO]) * sC, nCmy = -(sm[A.gCm + M_NYO] + sm[A.gC0 + M_NYO]) * sC;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double q1 = RM_LD(C_U + q, 1), q0 = RM_LD(C_U + q, 0), qm1 = RM_LD(C_U + q, -1), qm2 = RM_LD(C_U + q, -2);
      const double Ap = 26.0 * q0 - (q1 + qm1), Am = 26.0 * qm1 - (q0 + qm2);
      const double rm2 = sm[A.rp[0] + q * RQS], rm1 = sm[A.rp[1] + q * RQS], r0 = sm[A.rp[2] + q * RQS], r1 = sm[A.rp[3] + q * RQS],
                   r2 = sm[A.rp[4] + q * RQS];
      const double Cm = 7.0 * (rm1 + r0) - (rm2 + r1), Cp = 7.0 * (r0 + r1) - (rm1 + r2);
      gx[q] = Ap * nApx + Am * nAmx + Cp * nCpx + Cm * nCmx;
      gy[q] = Ap * nApy + Am * nAmy + Cp * nCpy + Cm * nCmy;
      fv[q] = 0.0625 * r0;
    }
  }
  const double mmu = 0.0625 * (9.0 * (RM_LD(C_MU, -1) + RM_LD(C_MU, 0)) - (RM_LD(C_MU, -2) + RM_LD(C_MU, 1)));
  constexpr double TWOTHIRD = 2.0 / 3.0;
  const double lambda = mmu * cs.cpprandtl;
  const double fvrou = TWOTHIRD * mmu * (2.0 * gx[0] - gy[1]);
  const double fvrov = mmu * (gy[0] + gx[1]);
  const double fvrow = mmu * gx[2];
  const double gvrov = TWOTHIRD * mmu * (2.0 * gy[1] - gx[0]);
  const double gvrow = mmu * gy[2];
  const double fvroe = lambda * gx[3] + fv[0] * fvrou + fv[1] * fvrov + fv[2] * fvrow;
  const double gvroe = lambda * gy[3] + fv[0] * fvrov + fv[1] * gvrov + fv[2] * gvrow;
  const double visc[5] = {0.0, fvrou * nxf + fvrov * nyf, fvrov * nxf + gvrov * nyf, fvrow * nxf + gvrow * nyf, fvroe * nxf + gvroe * nyf};

  // ---- Roe spectral radius (spectralradius_{i,j}.F) ---------------------------------------------------------------
  const double nx2 = nxf * nxf + nyf * nyf;
  double rspec;
  {
    const double sr = RM_LD(C_SR, 0), sl = RM_LD(C_SR, -1);
    const double inv = frcp(sl + sr);
    const double rr = sl * inv, omrr = sr * inv;
    const double u = RM_LD(C_U, -1) * rr + RM_LD(C_U, 0) * omrr;
    const double v = RM_LD(C_V, -1) * rr + RM_LD(C_V, 0) * omrr;
    const double c2x = (cs.gam * cs.rgaz) * (RM_LD(C_T, -1) * rr + RM_LD(C_T, 0) * omrr);
    rspec = ::fabs(nxf * u + nyf * v) + ::sqrt(c2x * nx2);
  }

  // ---- Jameson x Ducros x dilatation sensor (ducrosfordnc_{i,j}.F) -------------------------------------------------
  double eps2 = 0.0;
  if (cs.k2 != 0.0) {
    const double pm2 = RM_LD(C_P, -2), pm1 = RM_LD(C_P, -1), p0 = RM_LD(C_P, 0), pp1 = RM_LD(C_P, 1);
    const double a1_ = ::fabs(pm1 - 2.0 * p0 + pp1), b1_ = ::fabs(pm1 + 2.0 * p0 + pp1);
    const double a2_ = ::fabs(pm2 - 2.0 * pm1 + p0), b2_ = ::fabs(pm2 + 2.0 * pm1 + p0);
    const bool second = a1_ * b2_ < a2_ * b1_;   // k1 < k2: max() takes the second operand
    const double ks = (second ? a2_ : a1_) * frcp(second ? b2_ : b1_);
    const double duc = ::fmax(sm[A.s0 + S_DU * SARR], sm[A.s1 + S_DU * SARR]);
    const double sn = ::sqrt(nx2);
    const double t0 = sm[A.s0 + S_DV * SARR], t1 = sm[A.s1 + S_DV * SARR];
    const double d0 = RM_LD(C_CS, 0) * sn + 1e-15, d1 = RM_LD(C_CS, -1) * sn + 1e-15;
    const bool take1 = t0 * d1 > t1 * d0;   // x0 > x1: the dilatation switch is decreasing
    const double xs = 2.5 + 10.0 * (take1 ? t1 : t0) * frcp(take1 ? d1 : d0);
    const double dxm = frcp(1.0 + ::exp(::fmin(2.0 * xs, 700.0)));   // (1 - tanh x) / 2
    eps2 = cs.k2 * (ks * duc * dxm);
  }
  const double eps4 = ::fmax(0.0, cs.k4 - eps2 * 12.0);

  // ---- convective flux + 5th-difference operand: one pass over cells -3 .. 2 -----------------------------------
  constexpr double denom = 1.0 / 60.0;
  constexpr double ck[6] = {denom, -8.0 * denom, 37.0 * denom, 37.0 * denom, -8.0 * denom, denom};
  constexpr double dk[6] = {-denom, 5.0 * denom, -10.0 * denom, 10.0 * denom, -5.0 * denom, denom};
  double fx[5], pr[5];
#pragma unroll
  for (int e = 0; e < 5; ++e) fx[e] = pr[e] = 0.0;
  double pbar = 0.0;
#pragma unroll
  for (int k = -3; k <= 2; ++k) {
    const double u = RM_LD(C_U, k), v = RM_LD(C_V, k), p = RM_LD(C_P, k);
    const double cv = ck[k + 3] * (u * nxf + v * nyf);
    pbar += ck[k + 3] * p;
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      const double we = RM_LD(C_W + e, k);
      fx[e] += cv * (e == 4 ? we + p : we);
      pr[e] += dk[k + 3] * we;
    }
  }
  fx[1] += pbar * nxf;
  fx[2] += pbar * nyf;

  // ---- assembly (dissipation_ducros_{i,j}.F, fluxnumassembly_{i,j}.F) ----------------------------------------------
  const double e2h = 0.5 * eps2;
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    const double diff = RM_LD(C_W + e, 0) - RM_LD(C_W + e, -1);
    sm[A.xo + e * XES] = fx[e] - rspec * (e2h * diff + eps4 * pr[e]) - visc[e];
  }
#undef RM_LD
}

// ---- accessor over the rings for the reference-shaped templates of scheme.cuh (the three wall rows: first step of a wall
// segment, rows unwrapped because q < ring depth) ---------------------------------------------------------------------------------
struct MarchAcc {
  using DT = Zero;
  using VT = PVar;
  const double* s;    // cells + q * CROW + a
  const double* ss;   // sens + q * SW + ga   (ga = a - 2)
  const double *nx, *ny, *vol, *volf;
  long long c, n;
  int ldc, ldn;
  long long sc, sn;
  template <int OI_, int OJ_> BC_HD double raw(int arr) const { return s[arr * PC + OI_ + OJ_ * CROW]; }
  template <int OI_, int OJ_> BC_HD VT ld(int arr) const { return PVar{raw<OI_, OJ_>(arr), {}}; }
  template <int OI_, int OJ_> BC_HD VT W(int e) const { return ld<OI_, OJ_>(C_W + e); }
  template <int OI_, int OJ_> BC_HD VT U() const { return ld<OI_, OJ_>(C_U); }
  template <int OI_, int OJ_> BC_HD VT V() const { return ld<OI_, OJ_>(C_V); }
  template <int OI_, int OJ_> BC_HD VT Wz() const { return ld<OI_, OJ_>(C_WZ); }
  template <int OI_, int OJ_> BC_HD VT T() const { return ld<OI_, OJ_>(C_T); }
  template <int OI_, int OJ_> BC_HD VT P() const { return ld<OI_, OJ_>(C_P); }
  template <int OI_, int OJ_> BC_HD VT Mu() const { return ld<OI_, OJ_>(C_MU); }
  template <int OI_, int OJ_> BC_HD VT H() const { return (W<OI_, OJ_>(4) + P<OI_, OJ_>()) * (1.0 / W<OI_, OJ_>(0)); }
  template <int OI_, int OJ_> BC_HD auto SENS() const {   // S_DV holds vol * divu
    const double* p = ss + OI_ + OJ_ * SW;
    return CellSens<Zero, Zero>{PVar{p[S_DV * SARR], {}} / VOL<OI_, OJ_>(), PVar{p[S_DU * SARR], {}}};
  }
  template <int OI_, int OJ_> BC_HD double NX(int kk) const { return BC_LDG(nx + kk * sn + n + OI_ + (long long)OJ_ * ldn); }
  template <int OI_, int OJ_> BC_HD double NY(int kk) const { return BC_LDG(ny + kk * sn + n + OI_ + (long long)OJ_ * ldn); }
  template <int OI_, int OJ_> BC_HD double VOL() const { return BC_LDG(vol + c + OI_ + (long long)OJ_ * ldc); }
  template <int OI_, int OJ_> BC_HD double VOLF(int kk) const { return BC_LDG(volf + kk * sc + c + OI_ + (long long)OJ_ * ldc); }
};
BC_HD MarchAcc make_acc(const MCtx& t, int q, int a) {   // cell (i0-3+a, row of q); q + offsets must stay below the ring depths
  MarchAcc A;
  A.s = t.sm + O_CELLS + q * CROW + a;
  A.ss = t.sm + O_SENS + q * SW + (a - (H - 1));
  A.nx = t.nx; A.ny = t.ny; A.vol = t.vol; A.volf = t.volf;
  A.c = t.g.cidx(t.it.i0 - H + a, t.jrow(q));
  A.n = t.g.nidx(t.it.i0 - H + a, t.jrow(q));
  A.ldc = t.g.ldc; A.ldn = t.g.ldn; A.sc = t.g.sc; A.sn = t.g.sn;
  return A;
}
// the faces of the three wall rows (flux_num_dnc5.F90:161-220): i-faces of rows j <= 2 (o2 viscous gradients), j-faces j = 1
// (wall flux), 2 and 3 (off-centred fluxes).  Cold: first step of the segments that start at the wall.
RM_COLD void wall_face(const MarchAcc A, const SchemeConsts& c, double* X /* entry of equation 0 */, int dir, int fj) {
  PVar h[5];
  if (dir == 0) {
    face_flux<0, true, FACE_MAIN>(A, c, h);
  } else {
    if (fj == 1) face_flux<1, true, FACE_WALL>(A, c, h);
    else if (fj == 2) face_flux<1, true, FACE_NEAR3>(A, c, h);
    else face_flux<1, false, FACE_NEAR5>(A, c, h);
  }
#pragma unroll
  for (int e = 0; e < 5; ++e) X[e * XES] = h[e].v;
}

// ---- the face phase ---------------------------------------------------------------------------------------------------------------
// dir 0: i-face (i0 + col, row q) -> X_i[e][q - qJ][col];  dir 1: j-face (i0 + col, face row q between staged rows q-1 and q)
// -> X_j[e][q & 7][col]
BC_HD void face_task(const MCtx& t, int dir, int q, int col, int qJ) {
  const GridDesc& g = t.g;
  const int fi = t.it.i0 + col, fj = t.jrow(q);
  const int a = col + H;
  FaceArgs A;
  bool wallrow;
  if (dir == 0) {
    if (fi > g.im + 1 || fj > g.jm || fj > t.it.j1 || fj < t.it.j0) return;
    wallrow = t.wall && fj <= 2;
    A.xo = O_X + (q - qJ) * XIW + col;
    const int pc = cpos(q) + a;
#pragma unroll
    for (int k = 0; k < 6; ++k) A.rc[k] = pc + k - 3;
    const int r0 = O_R + col;
#pragma unroll
    for (int k = 0; k < 5; ++k) A.rp[k] = r0 + ((q + k - 2) & (NRI - 1)) * RIW;
    A.s1 = O_SENS + (q & (NRS - 1)) * SW + col;   // sensor cell i0-1+col = fi-1
    A.s0 = A.s1 + 1;
    const int pm = q % NRM;
    const int m0 = O_MET + pm * MROW + a, n0 = m0 + t.shift(q, 0), n1 = m0 + t.shift(q, 1);
    const int n1p = O_MET + ring_inc(pm, NRM) * MROW + a + t.shift(q + 1, 1);
    A.gvf = m0 + M_VF0;
    A.gA0 = n0 + M_NX0; A.gAp = A.gA0 + 1; A.gAm = A.gA0 - 1;
    A.gC0 = n1 + M_NX1; A.gCm = A.gC0 - 1; A.gCp = n1p + M_NX1; A.gCmp = A.gCp - 1;
  } else {
    if (fi > g.im || fj > g.jm + 1 || fj > t.it.j1 + 1 || fj < t.it.j0) return;
    wallrow = t.wall && fj <= 3;
    A.xo = O_X + X_J + (q & (NXJ - 1)) * W + col;
    int p = (q - 3) % NRC;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      A.rc[k] = O_CELLS + p * CROW + a;
      p = p + 1 == NRC ? 0 : p + 1;
    }
    const int r0 = O_R + R_J + (q & (NRJ - 1)) * RJW + col + 2;
#pragma unroll
    for (int k = 0; k < 5; ++k) A.rp[k] = r0 + k - 2;
    A.s0 = O_SENS + (q & (NRS - 1)) * SW + col + 1;
    A.s1 = O_SENS + ((q - 1) & (NRS - 1)) * SW + col + 1;
    int pm = (q - 1) % NRM;
    const int mm = O_MET + pm * MROW + a;
    pm = ring_inc(pm, NRM);
    const int m0 = O_MET + pm * MROW + a;
    pm = ring_inc(pm, NRM);
    const int mp = O_MET + pm * MROW + a;
    A.gvf = m0 + M_VF1;
    A.gA0 = m0 + t.shift(q, 1) + M_NX1; A.gAp = mp + t.shift(q + 1, 1) + M_NX1; A.gAm = mm + t.shift(q - 1, 1) + M_NX1;
    A.gC0 = m0 + t.shift(q, 0) + M_NX0; A.gCp = A.gC0 + 1; A.gCm = mm + t.shift(q - 1, 0) + M_NX0; A.gCmp = A.gCm + 1;
  }
  if (wallrow) wall_face(make_acc(t, q, a), t.c, t.sm + A.xo, dir, fj);
  else face_eval(t.c, t.sm, A);
}
// step faces: i-faces of rows qJ .. qJ+3 (warps 0-3, and the right-edge column on four lanes of warp 8), j-faces qJ+1 .. qJ+4
BC_HD void phase_faces(const MCtx& t, int tid, int qJ) {
  const int wp = tid >> 5, lane = tid & 31;
  int dir, q, col;
  if (wp < RB) { dir = 0; q = qJ + wp; col = lane; }
  else if (wp < 2 * RB) { dir = 1; q = qJ + 1 + (wp - RB); col = lane; }
  else if (lane < RB) { dir = 0; q = qJ + lane; col = W; }
  else return;
  face_task(t, dir, q, col, qJ);
}
// balance (rhs/balance.F:2-15) + coalesced store of rows qJ .. qJ+3: threads 0 .. 127
BC_HD void phase_balance(const MCtx& t, int tid, int qJ) {
  if (tid >= RB * W) return;
  const int r = tid >> 5, cx = tid & 31;
  const int q = qJ + r;
  const int i = t.it.i0 + cx, j = t.jrow(q);
  if (i > t.g.im || j > t.g.jm || j > t.it.j1 || j < t.it.j0) return;
  const double* XI = t.sm + O_X + r * XIW + cx;
  const double* XJ0 = t.sm + O_X + X_J + (q & (NXJ - 1)) * W + cx;
  const double* XJ1 = t.sm + O_X + X_J + ((q + 1) & (NXJ - 1)) * W + cx;
  const long long k = t.g.cidx(i, j);
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    const double ri_ = -(XI[e * XES + 1] - XI[e * XES]);
    t.res[e * t.g.sc + k] = ri_ - (XJ1[e * XES] - XJ0[e * XES]);
  }
}

// rows the prologue touches (q of the first output row of step s: qJ = 3 + 4 s)
constexpr int PRO_CELL_ROWS = 10;                 // cell rows q = 0 .. 9 staged and their primitives computed
constexpr int PRO_MET_Q0 = 2, PRO_MET_ROWS = 7;   // metric rows q = 2 .. 8
constexpr int NT_BAL = RB * W;                    // threads 0 .. 127 balance, threads 128 .. 287 primitives of the next rows

}  // namespace rm
}  // namespace bcast
