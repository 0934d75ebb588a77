// The fused primal residual of residual_fast.cuh as CUDA kernels for sm_100a.
//
//   k_residual_fast       DEFAULT (this file): one CTA of 320 threads per 32 x 9 tile, two CTAs per SM (5 warps on every SM
//                         sub-partition, what 96 registers allow); w loaded with LDG, all loads of a thread issued before its
//                         arithmetic; face metrics prefetched into registers one phase ahead of their use.
//   k_residual_fast_tma   variant RES_FAST_TMA / BROADCAST_B200_RESIDUAL_TMA=1 (residual_fast_tma.cu, 32 x 8 tile): PERSISTENT CTAs
//                         (2 per SM) walk the tiles; the five planes of w of the NEXT tile (38 x 14 x 5 box, halo included, zero
//                         fill outside the padded array) are delivered into a second shared-memory buffer by ONE TMA load
//                         (cp.async.bulk.tensor.3d + mbarrier) while the faces of the current tile are evaluated.  Falls back to
//                         k_residual_fast when TMA cannot describe the array (odd leading dimension: global strides must be
//                         multiples of 16 bytes).
// Measured on B200 at C5 (profiles/r1_e_summary.md): LDG kernel 2.51 ms (32 x 8) / 2.35 ms (32 x 9), TMA kernel 2.75 ms -- staging
// w by TMA removes the HBM wait of phase 0 but the metric loads of phases 1-3 stay exposed and the persistent loop costs more
// than it hides, so the LDG kernel stays the default.
//
// Phases and their barriers (both kernels):
//   0  (wait for the TMA) primitives         | 1  sensor cells + R_q of the i-faces | (1b ghost sensor cells, boundary tiles)
//   2  i-face fluxes -> exchange buffer      | 2b partial balance + R_q of the j-faces
//   3  j-face fluxes -> exchange buffer      | 3b balance, coalesced store of residu
// Reference: srcfv/rhs/flux_num_dnc5.F90:7-226.  Selected by launch_residual_tiled (residual_tile.cu keeps the first
// generation, BROADCAST_B200_RESIDUAL_V1=1, as a cross-check).
#include <cstdlib>
#include "kernels.cuh"
#include "residual_fast.cuh"

namespace bcast {

namespace {

__device__ __forceinline__ rf::TileCtx make_ctx(double* sm, const GridDesc& g, const SchemeConsts& c, double sqgr, bool wall, const double* w,
                                                const double* nx, const double* ny, const double* vol, const double* volf, double* res) {
  rf::TileCtx t(g, c);
  t.wsm = sm;
  t.sm = sm + rf::WBUF;
  t.sqgr = sqgr; t.wall = wall;
  t.w = w; t.nx = nx; t.ny = ny; t.vol = vol; t.volf = volf; t.res = res;
  t.i0 = 1; t.j0 = 1;
  return t;
}

// L2 prefetch of everything a LATER tile reads (w with halo, nx, ny, vol, volf rows): CTAs start in blockIdx order, so the CTA
// that will run `dist` tiles after this one finds its first loads in L2 instead of HBM (ncu r1_e: 18 % of all stall samples
// were the long-scoreboard wait at the top of each CTA).  One 128-byte line per thread and instruction, no register results.
__device__ __forceinline__ void prefetch_tile_l2(const GridDesc& g, const double* w, const double* nx, const double* ny, const double* vol,
                                                 const double* volf, int i0, int j0, int tid) {
  constexpr int LINES = (rf::PI * 8 + 127) / 128 + 1;   // lines per staged row (unaligned start)
  constexpr int ROWS = rf::PJ;
  const int jmax = g.jm + g.gh, imax = g.im + g.gh;
  for (int u = tid; u < 12 * ROWS * LINES; u += rf::NT) {
    const int plane = u / (ROWS * LINES), r = (u / LINES) % ROWS, l = u % LINES;
    const int gj = j0 - rf::H + r;
    const int gi = i0 - rf::H + l * 16;
    if (gj > jmax || gi > imax) continue;
    const double* p;
    if (plane < 5) p = w + plane * g.sc + g.cidx(gi, gj);
    else if (plane < 7) p = nx + (plane - 5) * g.sn + g.nidx(gi, gj);
    else if (plane < 9) p = ny + (plane - 7) * g.sn + g.nidx(gi, gj);
    else if (plane < 10) p = vol + g.cidx(gi, gj);
    else p = volf + (plane - 10) * g.sc + g.cidx(gi, gj);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  }
}

__global__ void __launch_bounds__(rf::NT, 2)
    k_residual_fast(int l2dist, int early, int ox, int oy, int ntx, int nty, GridDesc g, SchemeConsts c, double sqgr, bool wall, const double* __restrict__ w, const double* __restrict__ nx,
                    const double* __restrict__ ny, const double* __restrict__ vol, const double* __restrict__ volf,
                    double* __restrict__ res) {
  extern __shared__ __align__(128) double sm[];
  rf::TileCtx t = make_ctx(sm, g, c, sqgr, wall, w, nx, ny, vol, volf, res);
  // (ox, oy): first tile of this launch in the ntx x nty tile grid of the block (launches over a part of the tiles).
  // l2dist < 0: ring launch, a 1-D grid over the tiles outside the inner rectangle [1, ox] x [1, oy] (= bx1, by1):
  // first tile row, last tile rows, first tile column, last tile columns.
  int bx_ = blockIdx.x + ox, by_ = blockIdx.y + oy;
  const bool ring = l2dist < 0;
  // ring launch: tile of ring index u (first tile row, last tile rows, first tile column, last tile columns)
  auto ring_tile = [&](int u, int& bx, int& by) {
    const int bx1 = ox, by1 = oy;
    const int nA = ntx, nB = ntx * (nty - by1 - 1), nC = by1;
    if (u < nA) { bx = u; by = 0; }
    else if ((u -= nA) < nB) { bx = u % ntx; by = by1 + 1 + u / ntx; }
    else if ((u -= nB) < nC) { bx = 0; by = 1 + u; }
    else { u -= nC; const int wr = ntx - bx1 - 1; bx = bx1 + 1 + u % wr; by = 1 + u / wr; }
  };
  if (ring) ring_tile((int)blockIdx.x, bx_, by_);
  t.i0 = 1 + bx_ * rf::OI;
  t.j0 = 1 + by_ * rf::OJ;
  const int tid = threadIdx.x;
  if (ring) {
    // the ring follows the halo exchange and the boundary fills, which have just written what its tiles read: pull the tile one
    // wave of CTAs ahead (in ring order) into L2, as the whole-block launch does with its own distance
    const int u2 = (int)blockIdx.x - l2dist;   // l2dist = -(distance) in ring mode
    if (u2 < (int)gridDim.x) {
      int bx, by;
      ring_tile(u2, bx, by);
      prefetch_tile_l2(g, w, nx, ny, vol, volf, 1 + bx * rf::OI, 1 + by * rf::OJ, tid);
    }
  } else if (l2dist > 0) {
    const int L = by_ * ntx + bx_ + l2dist;
    const int bx = L % ntx, by = L / ntx;
    if (by < nty) prefetch_tile_l2(g, w, nx, ny, vol, volf, 1 + bx * rf::OI, 1 + by * rf::OJ, tid);
  }
  const rf::FaceGeom gi = rf::prefetch_iface(t, tid);   // metric loads in flight across phases 0 and 1
  // sensor-cell metrics: before phase 0 (early != 0) or at their use in phase 1 (with the L2 prefetch they are L2 hits there)
  rf::SensGeom sg0{}, sg1{};
  if (early) {
    sg0 = rf::prefetch_sensor(t, tid, 0);
    sg1 = rf::prefetch_sensor(t, tid, 1);
  }
  rf::phase0<false>(t, tid);
  __syncthreads();
  if (!early) {
    sg0 = rf::prefetch_sensor(t, tid, 0);
    sg1 = rf::prefetch_sensor(t, tid, 1);
  }
  rf::phase1(t, tid, sg0, sg1);
  __syncthreads();
  if (t.has_ghost_sensor()) {  // CTA-uniform
    rf::phase1b(t, tid);
    __syncthreads();
  }
  rf::phase2(t, tid, gi);
  const rf::FaceGeom gj = rf::prefetch_jface(t, tid);   // in flight across the balance / R_q phase
  __syncthreads();
  double r[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  rf::balance_i(t, tid, r);
  rf::phase_rj(t, tid);
  __syncthreads();
  rf::phase3(t, tid, gj);
  __syncthreads();
  rf::balance_j_store(t, tid, r);
}

template <class K>
cudaError_t prepare_kernel(K kernel, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // two CTAs per SM: ask for the largest shared-memory carve-out (the default heuristic picks 100 KB)
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

}  // namespace

cudaError_t launch_residual_fast_tma(const GridDesc& g, const SchemeConsts& c, double sqgr, bool wall, double* res, const double* w,
                                     const double* nx, const double* ny, const double* vol, const double* volf, cudaStream_t st, bool* done);

// part: 0 = every tile; 1 = the tiles that touch neither the first / last tile column nor the first / last tile row (they read no
// ghost cell and no slab halo column: they can run while the halo exchange and the boundary fills are still in flight);
// 2 = the remaining ring of tiles.  1 followed by 2 writes exactly what 0 writes.
cudaError_t launch_residual_fast(const GridDesc& g, const SchemeArgs& a, bool wall, double* res, const double* w, const double* nx,
                                 const double* ny, const double* vol, const double* volf, cudaStream_t st, bool tma, int part) {
  const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
  const double sqgr = ::sqrt(a.gam * a.rgaz);
  if (tma && part == 0) {
    bool done = false;
    cudaError_t e = launch_residual_fast_tma(g, c, sqgr, wall, res, w, nx, ny, vol, volf, st, &done);
    if (done || e != cudaSuccess) return e;
  }
  const int ntx = (g.im + rf::OI - 1) / rf::OI, nty = (g.jm + rf::OJ - 1) / rf::OJ;
  constexpr size_t SMEM = (size_t)rf::NSM * sizeof(double);
  static bool ready = false;
  if (!ready) {
    cudaError_t e = prepare_kernel(k_residual_fast, SMEM);
    if (e != cudaSuccess) return e;
    ready = true;
  }
  static const int l2dist = getenv("BROADCAST_B200_RESIDUAL_L2DIST") ? atoi(getenv("BROADCAST_B200_RESIDUAL_L2DIST")) : 592;
  static const int early = getenv("BROADCAST_B200_RESIDUAL_EARLY_SENSOR") ? atoi(getenv("BROADCAST_B200_RESIDUAL_EARLY_SENSOR")) : 1;
  auto go = [&](int ox, int oy, int cx, int cy, int dist) {
    if (cx > 0 && cy > 0) k_residual_fast<<<dim3(cx, cy), rf::NT, SMEM, st>>>(dist, early, ox, oy, ntx, nty, g, c, sqgr, wall, w, nx, ny, vol, volf, res);
  };
  // inner tiles: bx in [1, bx1], by in [1, by1] -- every cell they read (tile + gh halo) is an interior cell
  const int bx1 = (g.im - rf::OI - 3) / rf::OI, by1 = (g.jm - rf::OJ - 3) / rf::OJ;
  const bool has_inner = bx1 >= 1 && by1 >= 1;
  if (part == 0 || (part == 2 && !has_inner)) {
    go(0, 0, ntx, nty, l2dist);
  } else if (part == 1) {
    if (has_inner) go(1, 1, bx1, by1, l2dist);
  } else {   // the ring in ONE launch (1-D grid, tile found from the block index)
    const int nring = ntx + ntx * (nty - by1 - 1) + by1 + (ntx - bx1 - 1) * by1;
    k_residual_fast<<<nring, rf::NT, SMEM, st>>>(-296, early, bx1, by1, ntx, nty, g, c, sqgr, wall, w, nx, ny, vol, volf, res);   // -(prefetch distance): one wave
  }
  return cudaGetLastError();
}

}  // namespace bcast
