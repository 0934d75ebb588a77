// k_residual_fast: the fused primal residual of residual_fast.cuh as one CTA per 32 x 8 tile (288 threads, 86 KB of
// shared memory, two CTAs per SM).  Phases and their barriers:
//   0  stage w + halo, primitives            | 1  sensor cells + R_q of the i-faces | (1b ghost sensor cells, boundary tiles)
//   2  i-face fluxes -> exchange buffer      | 2b partial balance + R_q of the j-faces
//   3  j-face fluxes -> exchange buffer      | 3b balance, coalesced store of residu
// Reference: srcfv/rhs/flux_num_dnc5.F90:7-226.  Selected by launch_residual_tiled (residual_tile.cu keeps the first
// version, BROADCAST_B200_RESIDUAL_V1=1, as a cross-check).
#include "kernels.cuh"
#include "residual_fast.cuh"

namespace bcast {

namespace {

__global__ void __launch_bounds__(rf::NT, 2)
    k_residual_fast(GridDesc g, SchemeConsts c, double sqgr, bool wall, const double* __restrict__ w, const double* __restrict__ nx,
                    const double* __restrict__ ny, const double* __restrict__ vol, const double* __restrict__ volf,
                    double* __restrict__ res) {
  extern __shared__ double sm[];
  rf::TileCtx t;
  t.sm = sm; t.g = g; t.c = c; t.sqgr = sqgr; t.wall = wall;
  t.w = w; t.nx = nx; t.ny = ny; t.vol = vol; t.volf = volf; t.res = res;
  t.i0 = 1 + blockIdx.x * rf::OI;
  t.j0 = 1 + blockIdx.y * rf::OJ;
  const int tid = threadIdx.x;
  const rf::FaceGeom gi = rf::prefetch_iface(t, tid);   // metric loads in flight across phases 0 and 1
  rf::phase0(t, tid);
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

}  // namespace

cudaError_t launch_residual_fast(const GridDesc& g, const SchemeArgs& a, bool wall, double* res, const double* w, const double* nx,
                                 const double* ny, const double* vol, const double* volf, cudaStream_t st) {
  const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
  constexpr size_t SMEM = (size_t)rf::NSM * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_residual_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return e;
    // two CTAs per SM need 2 x 87 KB: ask for the largest shared-memory carve-out (the default heuristic picks 100 KB)
    e = cudaFuncSetAttribute(k_residual_fast, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((g.im + rf::OI - 1) / rf::OI, (g.jm + rf::OJ - 1) / rf::OJ);
  k_residual_fast<<<grid, rf::NT, SMEM, st>>>(g, c, ::sqrt(a.gam * a.rgaz), wall, w, nx, ny, vol, volf, res);
  return cudaGetLastError();
}

}  // namespace bcast
