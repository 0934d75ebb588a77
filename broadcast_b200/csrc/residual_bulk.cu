// k_residual_fast_bulk: the fused 32 x 9 tile residual of residual_fast.cuh with EVERY global input of the tile delivered by the
// bulk-copy engine (north_star: "i/j stencil tiles with their halo staged into shared memory by TMA").
//
// Why (ncu of k_residual_fast, profiles/r2_b_summary.md): the LDG kernel spends 750 of its 2 519 thread instructions per cell before
// its first barrier -- 615 of them integer / control work: 64-bit index arithmetic of ~34 metric loads per thread, the software L2
// prefetch loop with its divisions, the staging loop of w -- at 41 % issue utilisation (long scoreboard on the loads).  Here:
//   * warp 0 issues, lane-parallel, 3 TMA tensor loads (w box 38 x 15 x 5, vol box 34 x 11, volf box 34 x 10 x 2; zero fill outside
//     the arrays) and 48 one-dimensional bulk copies (one row of nx0 / nx1 / ny0 / ny1 each: node planes have an odd leading
//     dimension, no tensor map; a row is copied from the 16-byte aligned element at or below its first element and readers add
//     the parity `nshift`), all completing on ONE mbarrier; then the same list as L2 prefetches (cp.async.bulk.prefetch) for the
//     tile `l2dist` launches ahead;
//   * every thread waits on the mbarrier once; primitives, sensor metrics and the eight dual-cell normals of a face are then
//     computed from shared memory with compile-time offsets: no thread forms a global address except for the store of residu
//     (and the cold wall rows);
//   * phases, barriers and face formulas are those of k_residual_fast: results are bit-identical.
// Falls back to k_residual_fast when TMA cannot describe the arrays (odd cell leading dimension, unaligned base pointers).
// Reference: srcfv/rhs/flux_num_dnc5.F90:7-226.
#include <cuda.h>
#include <cstdint>
#include <cstdlib>
#include "kernels.cuh"
#include "residual_fast.cuh"

namespace bcast {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
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
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(0), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, int bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(0) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_1d(const void* src, int bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

struct BulkMaps {
  CUtensorMap w, vol, volf;
};

__global__ void __launch_bounds__(rf::NT, 2)
    k_residual_fast_bulk(const __grid_constant__ BulkMaps maps, int dbg, int l2dist, int ntx, int nty, GridDesc g, SchemeConsts c, double sqgr, bool wall,
                         const double* __restrict__ w, const double* __restrict__ nx, const double* __restrict__ ny,
                         const double* __restrict__ vol, const double* __restrict__ volf, double* __restrict__ res) {
  extern __shared__ __align__(128) double sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + rf::O_MET + rf::NMET);
  rf::TileCtx t(g, c);
  t.wsm = sm;
  t.sm = sm + rf::WBUF;
  t.met = sm + rf::O_MET;
  t.volbox = sm + rf::O_VOLBOX;
  t.sqgr = sqgr; t.wall = wall;
  t.w = w; t.nx = nx; t.ny = ny; t.vol = vol; t.volf = volf; t.res = res;
  t.i0 = 1 + blockIdx.x * rf::OI;
  t.j0 = 1 + blockIdx.y * rf::OJ;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(bar, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32) {
    // Warp 0 starts every copy: the three tensor boxes by lane 0 (their origins sit on even storage columns: a box whose first byte
    // is not 16-byte aligned raises an illegal-instruction fault), the 4 x MN_ROWS node rows dealt to the lanes in two rounds.
    // Every lane posts its own byte count.
    constexpr int NROW = 4 * rf::MN_ROWS, NRND = (NROW + 31) / 32;
    rf::BulkOp ops[NRND];
    uint32_t bytes = 0u;
    if (tid == 0) bytes = (uint32_t)(((dbg & 1) ? 0 : 5 * rf::NC) + ((dbg & 2) ? 0 : rf::MV_W * rf::MV_H) + ((dbg & 4) ? 0 : 2 * rf::MF_W * rf::MF_H)) * 8u;
#pragma unroll
    for (int k = 0; k < NRND; ++k) {
      const int op = 3 + tid + 32 * k;
      ops[k] = rf::bulk_op(g, nx, ny, t.i0, t.j0, op < rf::NBULK ? op : rf::NBULK);
      if (dbg & 8) ops[k].bytes = 0;
      bytes += (uint32_t)ops[k].bytes;
    }
    mbar_arrive_expect_tx(bar, bytes);
    if (tid == 0) {
      if (!(dbg & 1)) tma_load_3d(sm, &maps.w, t.i0 - 1, t.j0 - 1, bar);
      if (!(dbg & 2)) tma_load_2d(sm + rf::O_VOLBOX, &maps.vol, t.i0 + 1, t.j0 + 1, bar);
      if (!(dbg & 4)) tma_load_3d(sm + rf::O_MET + rf::M_VOLF, &maps.volf, t.i0 + 1, t.j0 + 2, bar);
    }
#pragma unroll
    for (int k = 0; k < NRND; ++k)
      if (ops[k].bytes > 0) bulk_load_1d(sm + ops[k].dst, ops[k].src, ops[k].bytes, bar);
    // the same list for the tile `l2dist` launches ahead, as L2 prefetches (CTAs start in blockIdx order)
    if (l2dist > 0) {
      const int L = blockIdx.y * ntx + blockIdx.x + l2dist;
      const int bx = L % ntx, by = L / ntx;
      if (by < nty) {
        const int pi0 = 1 + bx * rf::OI, pj0 = 1 + by * rf::OJ;
        if (tid == 0) {
          tma_prefetch_3d(&maps.w, pi0 - 1, pj0 - 1);
          tma_prefetch_2d(&maps.vol, pi0 + 1, pj0 + 1);
          tma_prefetch_3d(&maps.volf, pi0 + 1, pj0 + 2);
        }
#pragma unroll
        for (int k = 0; k < NRND; ++k) {
          const int op = 3 + tid + 32 * k;
          const rf::BulkOp o = rf::bulk_op(g, nx, ny, pi0, pj0, op < rf::NBULK ? op : rf::NBULK);
          if (o.bytes > 0) bulk_prefetch_1d(o.src, o.bytes);
        }
      }
    }
  }
  mbar_wait(bar, 0u);
  rf::phase0<true>(t, tid);
  __syncthreads();
  rf::phase1(t, tid, rf::sensor_geom_sm(t, tid, 0), rf::sensor_geom_sm(t, tid, 1));
  __syncthreads();
  if (t.has_ghost_sensor()) {  // CTA-uniform
    rf::phase1b(t, tid);
    __syncthreads();
  }
  rf::phase2(t, tid, rf::geom_iface_sm(t, tid));
  __syncthreads();
  double r[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  rf::balance_i(t, tid, r);
  rf::phase_rj(t, tid);
  __syncthreads();
  rf::phase3(t, tid, rf::geom_jface_sm(t, tid));
  __syncthreads();
  rf::balance_j_store(t, tid, r);
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

// tensor map of a cell-layout array seen as (ni, nj[, planes]) doubles with the given box; false if TMA cannot describe it
bool make_map(const GridDesc& g, const double* base, int planes, int bx, int by, CUtensorMap* map) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)g.ni(), (cuuint64_t)g.nj(), (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)g.ldc * sizeof(double), (cuuint64_t)g.sc * sizeof(double)};
  const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)planes};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, planes > 1 ? 3 : 2, const_cast<double*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// *done = true when the bulk kernel was launched; false when the bulk-copy engine cannot describe the arrays (the caller falls back to
// the LDG kernel)
cudaError_t launch_residual_fast_bulk(const GridDesc& g, const SchemeConsts& c, double sqgr, bool wall, double* res, const double* w,
                                      const double* nx, const double* ny, const double* vol, const double* volf, cudaStream_t st, bool* done) {
  *done = false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if ((g.ldc & 1) || !al16(w) || !al16(vol) || !al16(volf) || !al16(nx) || !al16(ny)) return cudaSuccess;
  // the maps depend on the base pointers and the grid only: cache the last set (a Newton loop calls with the same arrays)
  struct Key { const void *w, *vol, *volf; int im, jm; };
  static thread_local Key key{nullptr, nullptr, nullptr, 0, 0};
  static thread_local BulkMaps maps;
  if (key.w != w || key.vol != vol || key.volf != volf || key.im != g.im || key.jm != g.jm) {
    if (!make_map(g, w, 5, rf::PI, rf::PJ, &maps.w) || !make_map(g, vol, 1, rf::MV_W, rf::MV_H, &maps.vol) ||
        !make_map(g, volf, 2, rf::MF_W, rf::MF_H, &maps.volf))
      return cudaSuccess;
    key = Key{w, vol, volf, g.im, g.jm};
  }
  constexpr size_t SMEM = (size_t)rf::NSM_BULK * sizeof(double);
  static bool ready = false;
  if (!ready) {
    cudaError_t e = cudaFuncSetAttribute(k_residual_fast_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k_residual_fast_bulk, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    ready = true;
  }
  static const int l2dist = getenv("BROADCAST_B200_RESIDUAL_L2DIST") ? atoi(getenv("BROADCAST_B200_RESIDUAL_L2DIST")) : 592;
  const int ntx = (g.im + rf::OI - 1) / rf::OI, nty = (g.jm + rf::OJ - 1) / rf::OJ;
  static const int dbg = getenv("BROADCAST_B200_BULK_DEBUG") ? atoi(getenv("BROADCAST_B200_BULK_DEBUG")) : 0;
  k_residual_fast_bulk<<<dim3(ntx, nty), rf::NT, SMEM, st>>>(maps, dbg, l2dist, ntx, nty, g, c, sqgr, wall, w, nx, ny, vol, volf, res);
  *done = true;
  return cudaGetLastError();
}

}  // namespace bcast
