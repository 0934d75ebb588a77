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

__global__ void __launch_bounds__(rf::NT, 2)
    k_residual_fast(GridDesc g, SchemeConsts c, double sqgr, bool wall, const double* __restrict__ w, const double* __restrict__ nx,
                    const double* __restrict__ ny, const double* __restrict__ vol, const double* __restrict__ volf,
                    double* __restrict__ res) {
  extern __shared__ __align__(128) double sm[];
  rf::TileCtx t = make_ctx(sm, g, c, sqgr, wall, w, nx, ny, vol, volf, res);
  t.i0 = 1 + blockIdx.x * rf::OI;
  t.j0 = 1 + blockIdx.y * rf::OJ;
  const int tid = threadIdx.x;
  const rf::FaceGeom gi = rf::prefetch_iface(t, tid);   // metric loads in flight across phases 0 and 1
  rf::phase0<false>(t, tid);
  __syncthreads();
  rf::phase1(t, tid);
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

cudaError_t launch_residual_fast(const GridDesc& g, const SchemeArgs& a, bool wall, double* res, const double* w, const double* nx,
                                 const double* ny, const double* vol, const double* volf, cudaStream_t st, bool tma) {
  const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
  const double sqgr = ::sqrt(a.gam * a.rgaz);
  if (tma) {
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
  k_residual_fast<<<dim3(ntx, nty), rf::NT, SMEM, st>>>(g, c, sqgr, wall, w, nx, ny, vol, volf, res);
  return cudaGetLastError();
}

}  // namespace bcast
