// Generic residual / tangent pipeline, instantiated once per tangent width BCAST_N by
// generic_n0.cu / generic_n1.cu / generic_n5.cu (separate translation units keep build times low).
#include "kernels.cuh"
#include <cstdlib>

// scheme order of this translation unit (row f2 of SURVEY.md section 8): 5 unless the unit says otherwise.  The colour-loop helpers
// and the spanwise operators below exist for order 5 only; the other orders get the residual / tangent pipeline.
#ifndef BCAST_ORD
#define BCAST_ORD 5
#endif

namespace bcast {

// ---------------------------------------------------------------------------------------------
// primitives over every cell including ghosts (rhs/primvisc.F:2-9)
// ---------------------------------------------------------------------------------------------
template <int N>
__global__ void k_prims(GridDesc g, SchemeConsts c, RectList rl /* storage (0-based) index windows, inclusive */,
                        const double* __restrict__ w, const double* __restrict__ wd, double* __restrict__ prim,
                        double* __restrict__ primd) {
  using DT = TanOf<N>;
  const Rect rc = rl.r[blockIdx.z];
  const int ii = blockIdx.x * blockDim.x + threadIdx.x + rc.i0;
  const int jj = blockIdx.y * blockDim.y + threadIdx.y + rc.j0;
  if (ii > rc.i1 || jj > rc.j1) return;
  const long long k = ii + (long long)jj * g.ldc;
  Var<DT> q[5];
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    q[e].v = w[e * g.sc + k];
    if constexpr (N > 0) {
#pragma unroll
      for (int n = 0; n < N; ++n) q[e].d.d[n] = wd[(long long)(n * 5 + e) * g.sc + k];
    }
  }
  const CellPrims<DT> p = cell_prims(q, c);
  const Var<DT> out[NPRIM] = {p.u, p.v, p.w, p.t, p.p, p.mu, p.h};
#pragma unroll
  for (int s = 0; s < NPRIM; ++s) {
    prim[s * g.sc + k] = out[s].v;
    if constexpr (N > 0) {
#pragma unroll
      for (int n = 0; n < N; ++n) primd[(long long)(n * NPRIM + s) * g.sc + k] = out[s].d.d[n];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// gradients of velx, vely on interior cells (flux_num_dnc5.F90:124-137)
// ---------------------------------------------------------------------------------------------
template <int N, int ORD>
__global__ void k_grads(GridDesc g, FieldPtrs f, RectList rl /* interior cells, Fortran indices */, double* __restrict__ grad,
                        double* __restrict__ gradd) {
  const Rect rc = rl.r[blockIdx.z];
  const int i = blockIdx.x * blockDim.x + threadIdx.x + rc.i0;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + rc.j0;
  if (i > rc.i1 || j > rc.j1) return;
  GlobalAcc<N> a(f, g, i, j);
  const auto r = cell_gradients<0, 0, ORD>(a);
  const long long k = g.cidx(i, j);
  const decltype(r.u0) out[NGRAD] = {r.u0, r.u1, r.v0, r.v1};
#pragma unroll
  for (int s = 0; s < NGRAD; ++s) {
    grad[s * g.sc + k] = out[s].v;
    if constexpr (N > 0) {
#pragma unroll
      for (int n = 0; n < N; ++n) gradd[(long long)(n * NGRAD + s) * g.sc + k] = out[s].d.d[n];
    }
  }
}

// first ghost layer of the gradients by linear extrapolation (rhs/gradveloingh.F:1-19).  Only layer
// h = 1 at non-corner positions is ever read by the faces (sensor at faces i = 1, im+1, j = 1, jm+1).
static __global__ void k_grad_ghost(GridDesc g, double* __restrict__ grad, int nplanes_total) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = blockIdx.y;
  if (plane >= nplanes_total) return;
  double* p = grad + (long long)plane * g.sc;
  const int im = g.im, jm = g.jm;
  if (t < im) {  // j sides
    const int i = t + 1;
    p[g.cidx(i, 0)] = 2.0 * p[g.cidx(i, 1)] - p[g.cidx(i, 2)];
    p[g.cidx(i, jm + 1)] = 2.0 * p[g.cidx(i, jm)] - p[g.cidx(i, jm - 1)];
  } else if (t < im + jm) {  // i sides (physical ones only: a slab-internal edge holds computed gradients)
    const int j = t - im + 1;
    if (!(g.edges & 1)) p[g.cidx(0, j)] = 2.0 * p[g.cidx(1, j)] - p[g.cidx(2, j)];
    if (!(g.edges & 2)) p[g.cidx(im + 1, j)] = 2.0 * p[g.cidx(im, j)] - p[g.cidx(im - 1, j)];
  }
}

// ---------------------------------------------------------------------------------------------
// cell-centred balance (rhs/balance.F:2-15) of the four face fluxes of each cell
// ---------------------------------------------------------------------------------------------
template <int N, int DIR, class RD, int ORD = 5>
__device__ __forceinline__ void face_dispatch(const FieldPtrs& f, const GridDesc& g, const SchemeConsts& c, bool wall, int i, int j,
                                              Var<RD> (&hn)[5]) {
  GlobalAcc<N> a(f, g, i, j);
  face_by_row<DIR, ORD>(a, c, wall, j, hn);
}

// (ORD is part of the kernel's name: every order has its own translation unit, and two instantiations that differ only through a
// macro would be ONE symbol to the linker)
template <int N, int ORD>
__global__ void __launch_bounds__(128) k_balance(GridDesc g, SchemeConsts c, FieldPtrs f, bool wall, Rect rc, double* __restrict__ out) {
  using DT = TanOf<N>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x + rc.i0;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + rc.j0;
  if (i > rc.i1 || j > rc.j1) return;
  Var<DT> a0[5], a1[5], b0[5], b1[5];
  face_dispatch<N, 0, DT, ORD>(f, g, c, wall, i, j, a0);
  face_dispatch<N, 0, DT, ORD>(f, g, c, wall, i + 1, j, a1);
  face_dispatch<N, 1, DT, ORD>(f, g, c, wall, i, j, b0);
  face_dispatch<N, 1, DT, ORD>(f, g, c, wall, i, j + 1, b1);
  const long long k = g.cidx(i, j);
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    const Var<DT> r = -(a1[e] - a0[e]) - (b1[e] - b0[e]);
    if constexpr (N == 0) {
      out[e * g.sc + k] = r.v;
    } else {
#pragma unroll
      for (int n = 0; n < N; ++n) out[(long long)(n * 5 + e) * g.sc + k] = r.d.d[n];
    }
  }
}

template <int N, int ORD>
cudaError_t residual_generic_t(const GridDesc& g, const SchemeArgs& a, bool wall, double* out, const double* w, const double* wd,
                                      const double* nx, const double* ny, const double* vol, const double* volf, const Rect* rect,
                                      cudaStream_t st) {
  const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
  double* prim = scratch_doubles(0, (size_t)g.sc * NPRIM);
  double* grad = scratch_doubles(1, (size_t)g.sc * NGRAD);
  double* primd = N ? scratch_doubles(2, (size_t)g.sc * NPRIM * N) : nullptr;
  double* gradd = N ? scratch_doubles(3, (size_t)g.sc * NGRAD * N) : nullptr;
  if (!prim || !grad || (N && (!primd || !gradd))) return cudaErrorMemoryAllocation;
  dim3 blk(32, 4);
  Rect rc = rect ? *rect : Rect{1, g.im, 1, g.jm};
  if (rc.i1 < rc.i0 || rc.j1 < rc.j0) return cudaSuccess;
  // the balance of a cell reads gradients of its 4-neighbourhood (sensor) and primitives up to the scheme's GH
  // cells away; a gradient reads primitives NB cells away: restrict both passes to what `rect` needs
  const Rect rg{max(g.glo(), rc.i0 - 1), min(g.ghi(), rc.i1 + 1), max(1, rc.j0 - 1), min(g.jm, rc.j1 + 1)};
  // order 5: +-4 (wall row 1 reads row 5); other orders: the off-centred wall rows read cell rows 1 .. NP
  constexpr int PM = ORD == 5 ? 4 : (SchemeOrd<ORD>::NP > SchemeOrd<ORD>::GH + 1 ? SchemeOrd<ORD>::NP : SchemeOrd<ORD>::GH + 1);
  const Rect rp{max(0, rc.i0 - 1 - PM + g.gh), min(g.ni() - 1, rc.i1 - 1 + PM + g.gh), max(0, rc.j0 - 1 - PM + g.gh), min(g.nj() - 1, rc.j1 - 1 + PM + g.gh)};
  dim3 gall((rp.i1 - rp.i0 + 32) / 32, (rp.j1 - rp.j0 + 4) / 4);
  k_prims<N><<<gall, blk, 0, st>>>(g, c, one_rect(rp), w, wd, prim, primd);
  FieldPtrs f{w, prim, grad, nx, ny, vol, volf, wd, primd, gradd};
  dim3 gint((rg.i1 - rg.i0 + 32) / 32, (rg.j1 - rg.j0 + 4) / 4);
  k_grads<N, ORD><<<gint, blk, 0, st>>>(g, f, one_rect(rg), grad, gradd);
  {
    const int nt = g.im + g.jm;
    k_grad_ghost<<<dim3((nt + 127) / 128, NGRAD), 128, 0, st>>>(g, grad, NGRAD);
    if (N) k_grad_ghost<<<dim3((nt + 127) / 128, NGRAD * N), 128, 0, st>>>(g, gradd, NGRAD * N);
  }
  {
    dim3 gb((rc.i1 - rc.i0 + 32) / 32, (rc.j1 - rc.j0 + 4) / 4);
    k_balance<N, ORD><<<gb, blk, 0, st>>>(g, c, f, wall, rc, out);
  }
  return cudaGetLastError();
}


#define BCAST_CAT2(a, b) a##b
#define BCAST_CAT(a, b) BCAST_CAT2(a, b)

// ---------------------------------------------------------------------------------------------
// spanwise operators (srcfv/dz/coeffs_5p_dz.F90:10-174, coeffs_5p_dz2.F90): per interior cell
//   dz_out(e) = sum_k coeffs_k(w)(e) * d func_k(w)[wd](e)
// coefficient tables dz/coeffs_dz.F, dz/coeffs_dz2.F; functions matrix_dz/function_dz.F,
// matrix_dz2/function_dz2.F (their tangents: dz/function_5p_dz_d.f90:155-401, function_5p_dz2_d.f90).
// WHICH = 1: d/dz rows (cell + 5-point cross), WHICH = 2: d2/dz2 rows (cell-local).
// ---------------------------------------------------------------------------------------------
#if BCAST_N > 0 && BCAST_ORD == 5
template <int N, int WHICH>
__global__ void __launch_bounds__(128) k_dz(GridDesc g, SchemeConsts c, FieldPtrs f, Rect rc, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + rc.i0;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + rc.j0;
  if (i > rc.i1 || j > rc.j1) return;
  using DT = Tan<N>;
  const GlobalAcc<N> a(f, g, i, j);
  const long long k = g.cidx(i, j);
  const Var<DT> u = a.template U<0, 0>(), v = a.template V<0, 0>(), wz = a.template Wz<0, 0>();
  const Var<DT> mu = a.template Mu<0, 0>();
  constexpr double TWOTHIRD = 2.0 / 3.0;
  double r[5][N];
  if constexpr (WHICH == 2) {
    const Var<DT> t = a.template T<0, 0>();
    const double c1[5] = {0.0, -mu.v, -mu.v, -2.0 * TWOTHIRD * mu.v, -mu.v * c.cpprandtl};
    const double c2 = -mu.v * u.v, c3 = -mu.v * v.v, c4 = -2.0 * TWOTHIRD * mu.v * wz.v;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      r[0][n] = c1[0] * 0.0;
      r[1][n] = c1[1] * u.d.d[n];
      r[2][n] = c1[2] * v.d.d[n];
      r[3][n] = c1[3] * wz.d.d[n];
      r[4][n] = c1[4] * t.d.d[n] + c2 * u.d.d[n] + c3 * v.d.d[n] + c4 * wz.d.d[n];
    }
  } else {
    const Var<DT> p = a.template P<0, 0>();
    Var<DT> q[5];
#pragma unroll
    for (int e = 0; e < 5; ++e) q[e] = a.template W<0, 0>(e);
    // 5-point gradients (gradop_5pi.F, gradop_5pj.F, gradient.F, dxdy.F)
    constexpr double b1 = 8.0 * (1.0 / 12.0), b2 = -(1.0 / 12.0);
    const double volm1 = 1.0 / a.template VOL<0, 0>();
    const double dxm1 = 0.5 * (a.template NX<0, 0>(0) + a.template NX<1, 0>(0)) * volm1;
    const double dxm2 = 0.5 * (a.template NX<0, 0>(1) + a.template NX<0, 1>(1)) * volm1;
    const double dym1 = 0.5 * (a.template NY<0, 0>(0) + a.template NY<1, 0>(0)) * volm1;
    const double dym2 = 0.5 * (a.template NY<0, 0>(1) + a.template NY<0, 1>(1)) * volm1;
#define BC_GRAD(Q, G0, G1)                                                                                                        \
  auto G0##_i = b1 * (a.template Q<1, 0>() - a.template Q<-1, 0>()) + b2 * (a.template Q<2, 0>() - a.template Q<-2, 0>());       \
  auto G0##_j = b1 * (a.template Q<0, 1>() - a.template Q<0, -1>()) + b2 * (a.template Q<0, 2>() - a.template Q<0, -2>());       \
  auto G0 = dxm1 * G0##_i + dxm2 * G0##_j;                                                                                        \
  auto G1 = dym1 * G0##_i + dym2 * G0##_j;
    BC_GRAD(U, gu0, gu1)
    BC_GRAD(V, gv0, gv1)
    BC_GRAD(Wz, gw0, gw1)
    BC_GRAD(Mu, gm0, gm1)
#undef BC_GRAD
    (void)gu1;
    (void)gv0;
    const auto divu = gu0 + gv1;
    // functions (matrix_dz/function_dz.F) in tangent arithmetic
    const auto f0_1 = q[1] * wz - mu * gw0;
    const auto f0_2 = q[2] * wz - mu * gw1;
    const auto f0_3 = q[3] * wz + p + TWOTHIRD * mu * divu;
    const auto f0_4 = (q[4] + p) * wz;
    const auto f10 = mu * gw0;
    const auto f12 = mu * gw1;
    // coefficients (dz/coeffs_dz.F), passive
    const double muv = mu.v, uv = u.v, vv = v.v, wv = wz.v, gm0v = gm0.v, gm1v = gm1.v;
    const double k2_[5] = {0.0, TWOTHIRD * gm0v, TWOTHIRD * gm1v, -gm0v, TWOTHIRD * (gm0v * uv + muv * gu0.v)};
    const double k3_[5] = {0.0, TWOTHIRD * muv, TWOTHIRD * muv, -gm1v, TWOTHIRD * muv * uv};
    const double k4_3 = -muv, k4_4 = -(gm0v * wv + muv * gw0.v);
    const double k5 = -muv * wv;
    const double k6 = TWOTHIRD * (gm1v * vv + muv * gv1.v);
    const double k7 = TWOTHIRD * muv * vv;
    const double k8 = -(gm1v * wv + muv * gw1.v);
    const double k9 = -muv * wv;
    const double k10 = -muv * gw0.v;
    const double k11 = -uv;
    const double k12 = -muv * gw1.v;
    const double k13 = -vv;
    const double k14 = TWOTHIRD * muv * divu.v;
    const double k15 = TWOTHIRD * wv * divu.v;
    const double k16 = TWOTHIRD * muv * wv;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      r[0][n] = q[3].d.d[n];
      r[1][n] = f0_1.d.d[n] + k2_[1] * wz.d.d[n] + k3_[1] * gw0.d.d[n];
      r[2][n] = f0_2.d.d[n] + k2_[2] * wz.d.d[n] + k3_[2] * gw1.d.d[n];
      r[3][n] = f0_3.d.d[n] + k2_[3] * u.d.d[n] + k3_[3] * v.d.d[n] + k4_3 * divu.d.d[n];
      r[4][n] = f0_4.d.d[n] + k2_[4] * wz.d.d[n] + k3_[4] * gw0.d.d[n] + k4_4 * u.d.d[n] + k5 * gu0.d.d[n] + k6 * wz.d.d[n] +
                k7 * gw1.d.d[n] + k8 * v.d.d[n] + k9 * gv1.d.d[n] + k10 * u.d.d[n] + k11 * f10.d.d[n] + k12 * v.d.d[n] +
                k13 * f12.d.d[n] + k14 * wz.d.d[n] + k15 * mu.d.d[n] + k16 * divu.d.d[n];
    }
  }
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int e = 0; e < 5; ++e) out[(long long)(n * 5 + e) * g.sc + k] = r[e][n];
}

cudaError_t BCAST_CAT(dz_generic_, BCAST_N)(const GridDesc& g, const SchemeArgs& a, int which, double* out, const double* w, const double* wd,
                                           const double* nx, const double* ny, const double* vol, const double* volf, const Rect* rect,
                                           cudaStream_t st) {
  constexpr int N = BCAST_N;
  const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
  double* prim = scratch_doubles(0, (size_t)g.sc * NPRIM);
  double* primd = scratch_doubles(2, (size_t)g.sc * NPRIM * N);
  if (!prim || !primd) return cudaErrorMemoryAllocation;
  Rect rc = rect ? *rect : Rect{1, g.im, 1, g.jm};
  if (rc.i1 < rc.i0 || rc.j1 < rc.j0) return cudaSuccess;
  const Rect rp{max(0, rc.i0 - 3 + g.gh), min(g.ni() - 1, rc.i1 + 1 + g.gh), max(0, rc.j0 - 3 + g.gh), min(g.nj() - 1, rc.j1 + 1 + g.gh)};
  dim3 blk(32, 4);
  dim3 gall((rp.i1 - rp.i0 + 32) / 32, (rp.j1 - rp.j0 + 4) / 4);
  k_prims<N><<<gall, blk, 0, st>>>(g, c, one_rect(rp), w, wd, prim, primd);
  FieldPtrs f{w, prim, nullptr, nx, ny, vol, volf, wd, primd, nullptr};
  dim3 gb((rc.i1 - rc.i0 + 32) / 32, (rc.j1 - rc.j0 + 4) / 4);
  if (which == 1)
    k_dz<N, 1><<<gb, blk, 0, st>>>(g, c, f, rc, out);
  else
    k_dz<N, 2><<<gb, blk, 0, st>>>(g, c, f, rc, out);
  return cudaGetLastError();
}
#endif

#if BCAST_N == 5 && BCAST_ORD == 5
// ---------------------------------------------------------------------------------------------
// Tangent of the rows of up to four rectangles (the boundary strips of the Jacobian assembly) in ONE pass:
// block = 32 cells x 4 faces, thread (x, z) evaluates ONE face flux of cell x in 5-direction tangent arithmetic
// (a quarter of k_balance<5>'s per-thread work and no register spills: the strip launches are latency bound),
// the four faces of a cell are combined through shared memory exactly as rhs/balance.F does.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_balance_faces5(GridDesc g, SchemeConsts c, FieldPtrs f, bool wall, RectList rl, double* __restrict__ out) {
  // blockIdx.z = rect * 5 + direction: each thread differentiates ONE face in ONE direction (Tan<1>).
  // The 32 cells of a block are consecutive cells of the rectangle in row-major order (i fastest), so that the
  // gh-wide column strips (3 x jm cells) fill their warps just like the row strips (im x 3).
  __shared__ double sh[4][5][32];
  const int dir = blockIdx.z % 5;
  const Rect rc = rl.r[blockIdx.z / 5];
  const int wi = rc.i1 - rc.i0 + 1, wj = rc.j1 - rc.j0 + 1;
  const int n = blockIdx.x * 32 + threadIdx.x;
  const bool act = n < wi * wj;
  const int i = rc.i0 + n % wi;
  const int j = rc.j0 + n / wi;
  const int face = threadIdx.z;
  // direction `dir` of the 5-direction arrays seen as a 1-direction field set
  FieldPtrs f1 = f;
  f1.wd = f.wd + (long long)dir * 5 * g.sc;
  f1.primd = f.primd + (long long)dir * NPRIM * g.sc;
  f1.gradd = f.gradd + (long long)dir * NGRAD * g.sc;
  if (act) {
    Var<Tan<1>> hn[5];
    if (face == 0) face_dispatch<1, 0>(f1, g, c, wall, i, j, hn);
    else if (face == 1) face_dispatch<1, 0>(f1, g, c, wall, i + 1, j, hn);
    else if (face == 2) face_dispatch<1, 1>(f1, g, c, wall, i, j, hn);
    else face_dispatch<1, 1>(f1, g, c, wall, i, j + 1, hn);
#pragma unroll
    for (int e = 0; e < 5; ++e) sh[face][e][threadIdx.x] = hn[e].d.d[0];
  }
  __syncthreads();
  if (!act) return;
  const long long k = g.cidx(i, j);
  for (int e = face; e < 5; e += 4) {
    const double r = -(sh[1][e][threadIdx.x] - sh[0][e][threadIdx.x]) - (sh[3][e][threadIdx.x] - sh[2][e][threadIdx.x]);
    out[(long long)(dir * 5 + e) * g.sc + k] = r;
  }
}

// ---------------------------------------------------------------------------------------------
cudaError_t launch_tangent_tile5(const GridDesc& g, const SchemeArgs& a, bool wall, const Rect& rc, double* out5, const double* w,
                                 const double* wd5, const double* nx, const double* ny, const double* vol, const double* volf,
                                 cudaStream_t st);

// Strip tangent, second version: one block = a tile of 32 x th (th <= 3) or tw x 32 (tw <= 3) cells of a boundary strip
// and ONE direction (rectangles thicker than a strip are cut into bands of three rows, blockIdx.y: the full colour loop of
// bcd_jacobian_coo uses the same kernel).  Every face of the tile is evaluated ONCE (33*th + 32*(th+1) <= 227 faces on 256 threads instead of
// 4 per cell), and only if a tangent input of its stencil is non-zero in this direction: the block first builds the
// activity map of its cell window from wd (seeds and linearised ghost fills), a face is skipped when none of the cells
// of its stencil is active and the tangents of its two sensor gradients are zero (with seeds 7 cells apart about half of
// the faces of a colour pass see no seed at all).  Faces are combined per cell through shared memory as rhs/balance.F does.
// ---------------------------------------------------------------------------------------------
// TL = cells of a tile along the strip (32: 227 faces on 256 threads; 16: 115 faces on 128 threads), MINB = CTAs per SM the
// register allocation aims at (the face flux in tangent arithmetic wants ~254 registers: 8 warps per SM; fewer registers =
// more warps to hide the dependent FP64 latency, at the price of spills)
template <int TL, int MINB>
__global__ void __launch_bounds__(TL * 8, MINB) k_strip_faces5(GridDesc g, SchemeConsts c, FieldPtrs f, bool wall, Rect rc, int wide,
                                                               double* __restrict__ out) {
  constexpr int SF_MAXF = (TL + 1) * 3 + TL * 4;
  __shared__ double sh[5][SF_MAXF];
  __shared__ unsigned char flag[(TL + 6) * 9];
  const int dir = blockIdx.z;
  // blockIdx.y: band of three rows (wide) / three columns (tall) of a rectangle thicker than a strip
  const int ti0 = wide ? rc.i0 + TL * blockIdx.x : rc.i0 + 3 * blockIdx.y;
  const int tj0 = wide ? rc.j0 + 3 * blockIdx.y : rc.j0 + TL * blockIdx.x;
  const int twc = min(wide ? TL : 3, rc.i1 - ti0 + 1);
  const int thc = min(wide ? 3 : TL, rc.j1 - tj0 + 1);
  const int fw = twc + 6, fh = thc + 6;   // activity window: cells ti0-3 .. ti0+twc+2, tj0-3 .. tj0+thc+2
  FieldPtrs f1 = f;
  f1.wd = f.wd + (long long)dir * 5 * g.sc;
  f1.primd = f.primd + (long long)dir * NPRIM * g.sc;
  f1.gradd = f.gradd + (long long)dir * NGRAD * g.sc;
  for (int idx = threadIdx.x; idx < fw * fh; idx += blockDim.x) {
    const int ci = ti0 - 3 + idx % fw, cj = tj0 - 3 + idx / fw;
    bool a = false;
    if (ci >= 1 - g.gh && ci <= g.im + g.gh && cj >= 1 - g.gh && cj <= g.jm + g.gh) {
      const long long k = g.cidx(ci, cj);
#pragma unroll
      for (int e = 0; e < 5; ++e) a = a || (f1.wd[e * g.sc + k] != 0.0);
    }
    flag[idx] = a ? 1 : 0;
  }
  __syncthreads();
  const int nI = (twc + 1) * thc, nJ = twc * (thc + 1);
  const int t = threadIdx.x;
  if (t < nI + nJ) {
    const bool isI = t < nI;
    const int u = isI ? t : t - nI;
    const int pw = isI ? twc + 1 : twc;
    const int fi = ti0 + u % pw, fj = tj0 + u / pw;
    // stencil of a regular face: along -3 .. 2 on its own row, along -2 .. 1 on the cross rows +-1, +-2; the off-centred
    // wall flux of face j = 2 reaches j = 5 (along +3)
    bool act = false;
    const int bi = fi - (ti0 - 3), bj = fj - (tj0 - 3);   // window coordinates of the face cell
    auto chk = [&](int di, int dj) -> bool {   // outside the window: assume active
      const int x = bi + di, y = bj + dj;
      return (x < 0 || x >= fw || y < 0 || y >= fh) ? true : flag[x + y * fw] != 0;
    };
    if (isI) {
      for (int s_ = -3; s_ <= 2; ++s_) act = act || chk(s_, 0);
      for (int t_ = -2; t_ <= 2; ++t_)
        for (int s_ = -2; s_ <= 1; ++s_) act = act || chk(s_, t_);
    } else {
      const int hi = (wall && fj == 2) ? 3 : 2;
      for (int s_ = -3; s_ <= hi; ++s_) act = act || chk(0, s_);
      for (int t_ = -2; t_ <= 2; ++t_)
        for (int s_ = -2; s_ <= 1; ++s_) act = act || chk(t_, s_);
    }
    if (!act) {   // sensor gradients of the two face cells (extrapolated ones in the first ghost layer included)
      const long long k0 = g.cidx(fi, fj), k1 = isI ? k0 - 1 : k0 - g.ldc;
#pragma unroll
      for (int q = 0; q < NGRAD; ++q) act = act || (f1.gradd[q * g.sc + k0] != 0.0) || (f1.gradd[q * g.sc + k1] != 0.0);
    }
    double hd[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (act) {
      Var<Tan<1>> hn[5];
      if (isI) face_dispatch<1, 0>(f1, g, c, wall, fi, fj, hn);
      else face_dispatch<1, 1>(f1, g, c, wall, fi, fj, hn);
#pragma unroll
      for (int e = 0; e < 5; ++e) hd[e] = hn[e].d.d[0];
    }
#pragma unroll
    for (int e = 0; e < 5; ++e) sh[e][t] = hd[e];
  }
  __syncthreads();
  const int ncell = twc * thc;
  for (int idx = threadIdx.x; idx < 5 * ncell; idx += blockDim.x) {
    const int e = idx / ncell, cell = idx % ncell;
    const int cc = cell % twc, r = cell / twc;
    const double* hI = sh[e];
    const double* hJ = sh[e] + nI;
    const double v = -(hI[r * (twc + 1) + cc + 1] - hI[r * (twc + 1) + cc]) - (hJ[(r + 1) * twc + cc] - hJ[r * twc + cc]);
    out[(long long)(dir * 5 + e) * g.sc + g.cidx(ti0 + cc, tj0 + r)] = v;
  }
}

cudaError_t tangent_strips_5(const GridDesc& g, const SchemeArgs& a, bool wall, const RectList& rows_in, double* out5, const double* w,
                             const double* wd5, const double* nx, const double* ny, const double* vol, const double* volf, cudaStream_t st) {
  // Row-shaped rectangles (the Jlo / Jhi strips, the whole grid of the colour loop) go to the fused dual-number tile kernel
  // (residual_tangent.cu: no global primitive / gradient arrays); thin column strips keep the chain below.
  static const bool use_tile = getenv("BROADCAST_B200_TANGENT_TILE") == nullptr || atoi(getenv("BROADCAST_B200_TANGENT_TILE")) != 0;
  RectList tiled, rest;
  tiled.n = rest.n = 0;
  for (int k = 0; k < 4; ++k) tiled.r[k] = rest.r[k] = Rect{1, 0, 1, 0};
  for (int k = 0; k < rows_in.n; ++k) {
    const Rect& q = rows_in.r[k];
    const int wi = q.i1 - q.i0 + 1, wj = q.j1 - q.j0 + 1;
    if (use_tile && (wi >= wj || wi > 3)) tiled.r[tiled.n++] = q;
    else rest.r[rest.n++] = q;
  }
  if (tiled.n > 0) {
    cudaError_t et = cudaSuccess;
    for_each_rect(tiled, st, [&](const RectList& r1, int, cudaStream_t s1) {
      const cudaError_t e1 = launch_tangent_tile5(g, a, wall, r1.r[0], out5, w, wd5, nx, ny, vol, volf, s1);
      if (e1 != cudaSuccess) et = e1;
    });
    if (et != cudaSuccess) return et;
  }
  if (rest.n == 0) return cudaGetLastError();
  const RectList& rows = rest;
  constexpr int N = 5;
  const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
  double* prim = scratch_doubles(0, (size_t)g.sc * NPRIM);
  double* grad = scratch_doubles(1, (size_t)g.sc * NGRAD);
  double* primd = scratch_doubles(2, (size_t)g.sc * NPRIM * N);
  double* gradd = scratch_doubles(3, (size_t)g.sc * NGRAD * N);
  if (!prim || !grad || !primd || !gradd) return cudaErrorMemoryAllocation;
  RectList rg = rows, rp = rows;
  for (int k = 0; k < rows.n; ++k) {
    const Rect rc = rows.r[k];
    rg.r[k] = Rect{max(g.glo(), rc.i0 - 1), min(g.ghi(), rc.i1 + 1), max(1, rc.j0 - 1), min(g.jm, rc.j1 + 1)};
    rp.r[k] = Rect{max(0, rc.i0 - 5 + g.gh), min(g.ni() - 1, rc.i1 + 3 + g.gh), max(0, rc.j0 - 5 + g.gh), min(g.nj() - 1, rc.j1 + 3 + g.gh)};
  }
  dim3 blk(32, 4);
  for_each_rect(rp, st, [&](const RectList& r1, int, cudaStream_t s1) { k_prims<N><<<grid_of(r1, 32, 4), blk, 0, s1>>>(g, c, r1, w, wd5, prim, primd); });
  FieldPtrs f{w, prim, grad, nx, ny, vol, volf, wd5, primd, gradd};
  for_each_rect(rg, st, [&](const RectList& r1, int, cudaStream_t s1) { k_grads<N, 5><<<grid_of(r1, 32, 4), blk, 0, s1>>>(g, f, r1, grad, gradd); });
  const int nt = g.im + g.jm;
  k_grad_ghost<<<dim3((nt + 127) / 128, NGRAD), 128, 0, st>>>(g, grad, NGRAD);
  k_grad_ghost<<<dim3((nt + 127) / 128, NGRAD * N), 128, 0, st>>>(g, gradd, NGRAD * N);
  for_each_rect(rows, st, [&](const RectList& r1, int, cudaStream_t s1) {
    const Rect& q = r1.r[0];
    const int wi = q.i1 - q.i0 + 1, wj = q.j1 - q.j0 + 1;
    static const bool v1 = getenv("BROADCAST_B200_STRIPS_V1") != nullptr;
    static const int cfg_env = getenv("BROADCAST_B200_STRIPS_CFG") ? atoi(getenv("BROADCAST_B200_STRIPS_CFG")) : -1;
    const int wide = wi >= wj || wi > 3 ? 1 : 0;   // thin column strips run along j, everything else in bands of three rows
    if (!v1) {
      const int len = wide ? wi : wj;
      const int nband = wide ? (wj + 2) / 3 : (wi + 2) / 3;
      // measured on B200 (profiles/r1_f_summary.md): 16-cell tiles at 168 registers (12 warps per SM, 552 B of spills) win on long
      // strips (4096x1024: 22.2 -> 19.8 ms), 32-cell tiles at 254 registers on short ones
      const int cfg = cfg_env >= 0 ? cfg_env : (len >= 1024 ? 2 : 0);
      if (cfg == 1)
        k_strip_faces5<32, 2><<<dim3((len + 31) / 32, nband, 5), 256, 0, s1>>>(g, c, f, wall, q, wide, out5);
      else if (cfg == 2)
        k_strip_faces5<16, 3><<<dim3((len + 15) / 16, nband, 5), 128, 0, s1>>>(g, c, f, wall, q, wide, out5);
      else if (cfg == 3)
        k_strip_faces5<16, 4><<<dim3((len + 15) / 16, nband, 5), 128, 0, s1>>>(g, c, f, wall, q, wide, out5);
      else
        k_strip_faces5<32, 1><<<dim3((len + 31) / 32, nband, 5), 256, 0, s1>>>(g, c, f, wall, q, wide, out5);
    } else {
      const int ncell = wi * wj;
      k_balance_faces5<<<dim3((ncell + 31) / 32, 1, 5), dim3(32, 1, 4), 0, s1>>>(g, c, f, wall, r1, out5);
    }
  });
  return cudaGetLastError();
}
#endif

#if BCAST_N == 0 && BCAST_ORD == 5
// passive prims + gradients into the scratch arena, for kernels that read them (jac_interior.cu)
cudaError_t prepare_prims_grads(const GridDesc& g, const SchemeArgs& a, const double* w, const double* nx, const double* ny,
                                const double* vol, const double* volf, FieldPtrs& f, cudaStream_t st) {
  const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
  double* prim = scratch_doubles(0, (size_t)g.sc * NPRIM);
  double* grad = scratch_doubles(1, (size_t)g.sc * NGRAD);
  if (!prim || !grad) return cudaErrorMemoryAllocation;
  dim3 blk(32, 4);
  dim3 gall((g.ni() + 31) / 32, (g.nj() + 3) / 4);
  k_prims<0><<<gall, blk, 0, st>>>(g, c, one_rect(Rect{0, g.ni() - 1, 0, g.nj() - 1}), w, nullptr, prim, nullptr);
  f = FieldPtrs{w, prim, grad, nx, ny, vol, volf, nullptr, nullptr, nullptr};
  dim3 gint((g.im + 2 + 31) / 32, (g.jm + 3) / 4);
  k_grads<0, 5><<<gint, blk, 0, st>>>(g, f, one_rect(Rect{g.glo(), g.ghi(), 1, g.jm}), grad, nullptr);
  k_grad_ghost<<<dim3((g.im + g.jm + 127) / 128, NGRAD), 128, 0, st>>>(g, grad, NGRAD);
  return cudaGetLastError();
}
#endif

#if BCAST_ORD == 5
#define BCAST_RESIDUAL_GENERIC BCAST_CAT(residual_generic_, BCAST_N)
#else
#define BCAST_RESIDUAL_GENERIC BCAST_CAT(BCAST_CAT(BCAST_CAT(residual_generic_, BCAST_N), _o), BCAST_ORD)
#endif
cudaError_t BCAST_RESIDUAL_GENERIC(const GridDesc& g, const SchemeArgs& a, bool wall, double* out, const double* w, const double* wd,
                                   const double* nx, const double* ny, const double* vol, const double* volf, const Rect* rect,
                                   cudaStream_t st) {
  return residual_generic_t<BCAST_N, BCAST_ORD>(g, a, wall, out, w, wd, nx, ny, vol, volf, rect, st);
}

}  // namespace bcast
