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
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(0) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
struct BulkMaps {
  CUtensorMap w, vol, volf, nx, ny;
};

__global__ void __launch_bounds__(rf::NT, 2)
    k_residual_fast_bulk(const __grid_constant__ BulkMaps maps, int l2dist, int ox, int oy, GridDesc g, SchemeConsts c, double sqgr, bool wall,
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
  // (ox, oy): first tile of this launch in the tile grid of the block (whole block: 0, 0; inner tiles of the overlapped step: 1, 1)
  t.i0 = 1 + (blockIdx.x + ox) * rf::OI;
  t.j0 = 1 + (blockIdx.y + oy) * rf::OJ;
  const int tid = threadIdx.x;
  t.nsh = rf::node_shift_mask(g, t.i0, t.j0);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // out-of-range parts of a box are zero-filled and still counted: the byte total of a tile is a compile-time constant
    mbar_arrive_expect_tx(bar, (uint32_t)rf::BULK_BYTES);
  }
  __syncthreads();
  // The eleven tensor boxes of the tile, one mbarrier; lane 0 of warp w starts boxes w, w + 10 (one thread issuing all of them kept
  // the other nine warps at the barrier for ~250 instructions: 7.5 % of all stall samples, ncu r2_13), then the same boxes of the
  // tile `l2dist` launches ahead as L2 prefetches (CTAs start in blockIdx order).
  if ((tid & 31) == 0) {
    const int L = blockIdx.y * gridDim.x + blockIdx.x + l2dist;
    const int pbx = L % gridDim.x, pby = L / gridDim.x;
    const bool pf = l2dist > 0 && pby < gridDim.y;
    const int pi0 = 1 + (pbx + ox) * rf::OI, pj0 = 1 + (pby + oy) * rf::OJ;
    for (int op = tid >> 5; op < rf::NBULK; op += rf::NT / 32) {
      const rf::BulkOp o = rf::bulk_op(g, t.i0, t.j0, op);
      const CUtensorMap* mp = op == 0 ? &maps.w : op == 1 ? &maps.vol : op == 2 ? &maps.volf : ((op - 3) >> 2) ? &maps.ny : &maps.nx;
      if (op == 0 || op == 2) tma_load_3d(sm + o.dst, mp, o.x, o.y, bar);
      else tma_load_2d(sm + o.dst, mp, o.x, o.y, bar);
      if (pf) {
        const rf::BulkOp q = rf::bulk_op(g, pi0, pj0, op);
        if (op == 0 || op == 2) tma_prefetch_3d(mp, q.x, q.y);
        else tma_prefetch_2d(mp, q.x, q.y);
      }
    }
  }
  // ONE thread polls the mbarrier, the others sleep in the hardware barrier behind it: ten polling warps woke up up to a microsecond
  // apart and their try_wait / branch pairs were 210 of the 2 540 thread instructions per cell of the first bulk version (ncu r2_11).
  // bar.sync orders the copies that thread 0 has observed complete for every thread of the CTA.
  if (tid == 0) mbar_wait(bar, 0u);
  __syncthreads();
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

// a node-layout array (2 planes of (nj + 1) rows of ldn doubles, ldn odd on even grids) seen as (nj + 1) rows of 2 ldn doubles
bool make_node_map(const GridDesc& g, const double* base, CUtensorMap* map) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)(2 * g.ldn), (cuuint64_t)(g.nj() + 1)};
  const cuuint64_t strides[1] = {(cuuint64_t)(2 * g.ldn) * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)rf::MN_SLOT, (cuuint32_t)rf::MN_HALF};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// *done = true when the bulk kernel was launched; false when the bulk-copy engine cannot describe the arrays (the caller falls back to
// the LDG kernel)
cudaError_t launch_residual_fast_bulk(const GridDesc& g, const SchemeConsts& c, double sqgr, bool wall, double* res, const double* w,
                                      const double* nx, const double* ny, const double* vol, const double* volf, cudaStream_t st, bool* done,
                                      int part) {
  *done = false;
  if (part == 2) return cudaSuccess;   // the ring of tiles stays with the LDG kernel's 1-D ring launch (bit-identical results)
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if ((g.ldc & 1) || !al16(w) || !al16(vol) || !al16(volf) || !al16(nx) || !al16(ny)) return cudaSuccess;
  // the maps depend on the base pointers and the grid only: cache the last set (a Newton loop calls with the same arrays)
  struct Key { const void *w, *vol, *volf, *nx, *ny; int im, jm; };
  static thread_local Key key{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
  static thread_local BulkMaps maps;
  if (key.w != w || key.vol != vol || key.volf != volf || key.nx != nx || key.ny != ny || key.im != g.im || key.jm != g.jm) {
    if (!make_map(g, w, 5, rf::PI, rf::PJ, &maps.w) || !make_map(g, vol, 1, rf::MV_W, rf::MV_H, &maps.vol) ||
        !make_map(g, volf, 2, rf::MF_W, rf::MF_H, &maps.volf) || !make_node_map(g, nx, &maps.nx) || !make_node_map(g, ny, &maps.ny))
      return cudaSuccess;
    key = Key{w, vol, volf, nx, ny, g.im, g.jm};
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
  if (part == 1) {
    // inner tiles (same rectangle as launch_residual_fast: every cell they read is an interior cell of the block)
    const int bx1 = (g.im - rf::OI - 3) / rf::OI, by1 = (g.jm - rf::OJ - 3) / rf::OJ;
    if (bx1 < 1 || by1 < 1) return cudaSuccess;   // no inner tile: the ring call (LDG kernel) does everything
    k_residual_fast_bulk<<<dim3(bx1, by1), rf::NT, SMEM, st>>>(maps, l2dist, 1, 1, g, c, sqgr, wall, w, nx, ny, vol, volf, res);
  } else {
    k_residual_fast_bulk<<<dim3(ntx, nty), rf::NT, SMEM, st>>>(maps, l2dist, 0, 0, g, c, sqgr, wall, w, nx, ny, vol, volf, res);
  }
  *done = true;
  return cudaGetLastError();
}

}  // namespace bcast
