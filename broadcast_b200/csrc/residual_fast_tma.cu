// k_residual_fast_tma: the fused residual of residual_fast.cuh with PERSISTENT CTAs and TMA staging of w (see residual_fast.cu
// for the variants and the measurements).  Compiled with a 32 x 8 tile: two w buffers (2 x 21 KB) + the derived arrays fit twice
// per SM only at that height.
#define BCAST_RF_OJ 8
#define BCAST_RF_NS rf8
#include <cuda.h>
#include <cstdint>
#include <cstdlib>
#include "kernels.cuh"
#include "residual_fast.cuh"

namespace bcast {
namespace rf = rf8;

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// one 3-D tiled TMA load: box (PI, PJ, 5) of w at storage coordinates (x, y, 0) -> dst, completion on bar
__device__ __forceinline__ void tma_load_w(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(0), "r"(smem_u32(bar))
      : "memory");
}

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
    k_residual_fast_tma(const __grid_constant__ CUtensorMap wmap, GridDesc g, SchemeConsts c, double sqgr, bool wall,
                        const double* __restrict__ w, const double* __restrict__ nx, const double* __restrict__ ny,
                        const double* __restrict__ vol, const double* __restrict__ volf, double* __restrict__ res, int ntx, int ntiles) {
  extern __shared__ __align__(128) double sm[];
  // layout: [w buffer 0][w buffer 1][derived arrays, R buffer, exchange buffer][two mbarriers]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2 * rf::WBUF + rf::NSM_REST);
  rf::TileCtx t = make_ctx(sm, g, c, sqgr, wall, w, nx, ny, vol, volf, res);
  t.sm = sm + 2 * rf::WBUF;
  const int tid = threadIdx.x;
  constexpr uint32_t BYTES = 5u * rf::NC * sizeof(double);
  int tile = blockIdx.x;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tile >= ntiles) return;
  if (tid == 0) {
    mbar_expect_tx(&bar[0], BYTES);
    tma_load_w(sm, &wmap, (tile % ntx) * rf::OI, (tile / ntx) * rf::OJ, &bar[0]);   // storage coordinates of cell (i0-3, j0-3)
  }
  uint32_t parity = 0u;   // bit b: phase parity of mbarrier b
  int b = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    // buffer b^1 was read last by the faces of the previous tile; every thread has passed the barrier that follows them
    if (tid == 0 && next < ntiles) {
      mbar_expect_tx(&bar[b ^ 1], BYTES);
      tma_load_w(sm + (b ^ 1) * rf::WBUF, &wmap, (next % ntx) * rf::OI, (next / ntx) * rf::OJ, &bar[b ^ 1]);
    }
    t.i0 = 1 + (tile % ntx) * rf::OI;
    t.j0 = 1 + (tile / ntx) * rf::OJ;
    t.wsm = sm + b * rf::WBUF;
    const rf::FaceGeom gi = rf::prefetch_iface(t, tid);   // metric loads in flight across phases 0 and 1
    const rf::SensGeom sg0 = rf::prefetch_sensor(t, tid, 0), sg1 = rf::prefetch_sensor(t, tid, 1);
    mbar_wait(&bar[b], (parity >> b) & 1u);
    parity ^= 1u << b;
    rf::phase0<true>(t, tid);
    __syncthreads();
    rf::phase1(t, tid, sg0, sg1);
    __syncthreads();
    if (t.has_ghost_sensor()) {  // CTA-uniform
      rf::phase1b(t, tid);
      __syncthreads();
    }
    rf::phase2(t, tid, gi);
    const rf::FaceGeom gj = rf::prefetch_jface(t, tid);
    __syncthreads();
    double r[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    rf::balance_i(t, tid, r);
    rf::phase_rj(t, tid);
    __syncthreads();
    rf::phase3(t, tid, gj);
    __syncthreads();
    rf::balance_j_store(t, tid, r);
    b ^= 1;
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// tensor map of w seen as (ni, nj, 5) doubles, box (PI, PJ, 5); false if TMA cannot describe it
bool make_w_map(const GridDesc& g, const double* w, CUtensorMap* map) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(w) & 15) || (g.ldc & 1)) return false;   // base and strides: multiples of 16 bytes
  const cuuint64_t dims[3] = {(cuuint64_t)g.ni(), (cuuint64_t)g.nj(), 5};
  const cuuint64_t strides[2] = {(cuuint64_t)g.ldc * sizeof(double), (cuuint64_t)g.sc * sizeof(double)};
  const cuuint32_t box[3] = {rf::PI, rf::PJ, 5};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class K>
cudaError_t prepare_kernel(K kernel, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // two CTAs per SM: ask for the largest shared-memory carve-out (the default heuristic picks 100 KB)
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

}  // namespace

// *done = true when the TMA kernel was launched; false when TMA cannot describe w (the caller falls back to the LDG kernel)
cudaError_t launch_residual_fast_tma(const GridDesc& g, const SchemeConsts& c, double sqgr, bool wall, double* res, const double* w,
                                     const double* nx, const double* ny, const double* vol, const double* volf, cudaStream_t st, bool* done) {
  *done = false;
  CUtensorMap map;
  if (!make_w_map(g, w, &map)) return cudaSuccess;
  const int ntx = (g.im + rf::OI - 1) / rf::OI, nty = (g.jm + rf::OJ - 1) / rf::OJ;
  constexpr size_t SMEM = (size_t)rf::NSM_TMA * sizeof(double);
  static bool ready = false;
  static int nsm = 0;
  if (!ready) {
    cudaError_t e = prepare_kernel(k_residual_fast_tma, SMEM);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    ready = true;
  }
  const int ntiles = ntx * nty;
  const int grid = ntiles < 2 * nsm ? ntiles : 2 * nsm;
  k_residual_fast_tma<<<grid, rf::NT, SMEM, st>>>(map, g, c, sqgr, wall, w, nx, ny, vol, volf, res, ntx, ntiles);
  *done = true;
  return cudaGetLastError();
}

}  // namespace bcast
