// k_residual_march: the j-marching fused residual of residual_march.cuh as a persistent kernel for sm_100a (variant RES_MARCH /
// BROADCAST_B200_RESIDUAL_MARCH=1; measured 4 % behind the tile kernel at C5, see kernels.cuh and profiles/r2_a_summary.md).  Two CTAs of 320 threads (nine compute warps + one copy-issuing warp) per SM; a CTA walks work items (strip of 32 columns x segment of rows), keeps
// its rows in shared-memory rings and receives the rows of the next step while it evaluates the faces of the current one:
//   * w (5 planes), vol, volf: TMA, one cp.async.bulk.tensor box (38 x 1 x planes) per row, completion on two alternating mbarriers;
//   * nx, ny (node layout, odd leading dimension: no tensor map possible): LDGSTS (cp.async 8 bytes) by all threads.
// Falls back to k_residual_fast (residual_fast.cu) when TMA cannot describe the cell arrays (odd im: global strides must be
// multiples of 16 bytes).  Reference: srcfv/rhs/flux_num_dnc5.F90:7-226.
#include <cuda.h>
#include <cstdint>
#include <cstdlib>
#include "kernels.cuh"
#include "residual_march.cuh"

namespace bcast {

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
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct Maps {
  CUtensorMap w, vol, volf;
};

// One generation of copies (rows [cq0, cq0+cn) of the cell ring, [mq0, mq0+mn) of the metric ring), issued by ONE warp: every lane
// builds and starts its own operations; lane 0 first posts the byte total of the whole generation on the mbarrier.
__device__ __forceinline__ void start_copy(const rm::MCtx& t, const Maps& mp, uint64_t* bar, const rm::CopyOp& o) {
  double* dst = t.sm + o.dst;
  if (o.kind == 0) tma_load_3d(dst, &mp.w, o.x, o.y, bar);
  else if (o.kind == 1) tma_load_2d(dst, &mp.vol, o.x, o.y, bar);
  else if (o.kind == 2) tma_load_3d(dst, &mp.volf, o.x, o.y, bar);
  else if (o.kind == 3) bulk_load_1d(dst, o.src, (uint32_t)o.bytes, bar);
}
__device__ __forceinline__ void issue_rows(const rm::MCtx& t, const Maps& mp, uint64_t* bar, int lane, int cq0, int cn, int mq0, int mn) {
  const int nops = rm::copy_count(cn, mn);
  // a step has 28 operations: one per lane, built once; the prologue's 52 take a second round
  rm::CopyOp o0 = rm::copy_op(t, lane < nops ? lane : nops - 1, cq0, cn, mq0, mn);
  if (lane >= nops) { o0.kind = -1; o0.bytes = 0; }
  uint32_t bytes = (uint32_t)o0.bytes;
  for (int op = lane + 32; op < nops; op += 32) bytes += (uint32_t)rm::copy_op(t, op, cq0, cn, mq0, mn).bytes;
  bytes = __reduce_add_sync(0xffffffffu, bytes);
  if (lane == 0) mbar_expect_tx(bar, bytes);
  __syncwarp();
  start_copy(t, mp, bar, o0);
  for (int op = lane + 32; op < nops; op += 32) start_copy(t, mp, bar, rm::copy_op(t, op, cq0, cn, mq0, mn));
}

__global__ void __launch_bounds__(rm::NT_LAUNCH, 2)
    k_residual_march(const __grid_constant__ Maps mp, const __grid_constant__ GridDesc g, const __grid_constant__ SchemeConsts c, double sqgr,
                     bool wall, const double* __restrict__ w, const double* __restrict__ nx, const double* __restrict__ ny,
                     const double* __restrict__ vol, const double* __restrict__ volf, double* __restrict__ res, int nstrips, int nseg,
                     int seglen) {
  extern __shared__ __align__(128) double sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + rm::O_BAR);
  rm::MCtx t(g, c);
  t.sm = sm;
  t.sqgr = sqgr; t.wall = wall;
  t.w = w; t.nx = nx; t.ny = ny; t.vol = vol; t.volf = volf; t.res = res;
  const int tid = threadIdx.x, wp = tid >> 5, lane = tid & 31;
  constexpr int ISSUER = rm::NT / 32;   // warp 9 does nothing but start the copies: a compute warp that issued them was what every
                                        // other warp waited for at the next barrier (profiles/r2_a_summary.md)
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t gen = 0;   // copy generations issued so far: generation n completes on bar[n & 1] with parity (n >> 1) & 1
  const int nitems = nstrips * nseg;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int seg = item / nstrips, strip = item - seg * nstrips;
    t.it.i0 = 1 + strip * rm::W;
    t.it.j0 = 1 + seg * seglen;
    t.it.j1 = min(g.jm, (seg + 1) * seglen);
    const int nsteps = (t.it.j1 - t.it.j0 + rm::RB) / rm::RB;
    __syncthreads();   // the rings of the previous item are dead
    // ---- prologue: rows j0-3 .. j0+6 of the cells, j0-1 .. j0+5 of the metrics, primitives of all ten cell rows
    if (wp == ISSUER) issue_rows(t, mp, &bar[gen & 1], lane, 0, rm::PRO_CELL_ROWS, rm::PRO_MET_Q0, rm::PRO_MET_ROWS);
    mbar_wait(&bar[gen & 1], (gen >> 1) & 1u);
    ++gen;
    if (wp != ISSUER) rm::prims_rows(t, tid, rm::NT, 0, rm::PRO_CELL_ROWS);
    __syncthreads();
    // ---- steps; s = -1 evaluates the bottom j-faces of the segment (same code, rows below j0 are inactive) ----------------
    for (int s = -1; s < nsteps; ++s) {
      const int qJ = 3 + rm::RB * s;
      const bool more = s >= 0 && s + 1 < nsteps;   // step s+1 exists and its rows are not the prologue's
      if (more && wp == ISSUER) issue_rows(t, mp, &bar[gen & 1], lane, qJ + 7, rm::RB, qJ + 6, rm::RB);
      int sq0 = qJ + 1, sn = rm::RB, iq0 = qJ + 2, in = rm::RB, jq0 = qJ + 1, jn = rm::RB;
      if (s < 0) { sq0 = 2; sn = 3; iq0 = 1; in = 4; jq0 = 3; jn = 1; }
      if (wp != ISSUER) rm::phase_sens_r(t, tid, sq0, sn, iq0, in, jq0, jn);
      __syncthreads();
      if (rm::item_has_ghost_sensor(t, sq0, sn)) {
        if (wp != ISSUER) rm::phase_sens_ghost(t, tid, sq0, sn);
        __syncthreads();
      }
      if (wp != ISSUER) rm::phase_faces(t, tid, qJ);
      __syncthreads();
      if (tid < rm::NT_BAL) {
        rm::phase_balance(t, tid, qJ);
      } else if (more && wp != ISSUER) {   // the rows issued at the top of this step: primitives for step s+1
        mbar_wait(&bar[gen & 1], (gen >> 1) & 1u);
        rm::prims_rows(t, tid - rm::NT_BAL, rm::NT - rm::NT_BAL, qJ + 7, rm::RB);
      }
      if (more) ++gen;
      __syncthreads();
    }
  }
}

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

// tensor map of a cell-layout array seen as (ni, nj[, planes]) doubles, box (38, 1[, planes]); false if TMA cannot describe it
bool make_cell_map(const GridDesc& g, const double* base, int planes, CUtensorMap* map) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (g.ldc & 1) || (g.sc & 1)) return false;   // base and strides: multiples of 16 bytes
  const cuuint64_t dims[3] = {(cuuint64_t)g.ni(), (cuuint64_t)g.nj(), (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)g.ldc * sizeof(double), (cuuint64_t)g.sc * sizeof(double)};
  const cuuint32_t box[3] = {rm::PC, 1, (cuuint32_t)planes};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, planes > 1 ? 3 : 2, const_cast<double*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// segment length (multiple of RB) for `slots` resident CTAs: the candidate with the smallest modelled time
// (waves of items x (rows of a segment + the prologue's worth of extra rows))
int pick_seglen(int nstrips, int jm, int slots) {
  static const int forced = getenv("BROADCAST_B200_MARCH_SEG") ? atoi(getenv("BROADCAST_B200_MARCH_SEG")) : 0;
  if (forced > 0) return (forced + rm::RB - 1) / rm::RB * rm::RB;
  int best = rm::RB, best_cost = 1 << 30;
  for (int len = 32; len <= 512; len += rm::RB) {
    const int nseg = (jm + len - 1) / len;
    const long long items = (long long)nstrips * nseg;
    const int waves = (int)((items + slots - 1) / slots);
    const int last = jm - (nseg - 1) * len;   // rows of the last segment
    (void)last;
    const int cost = waves * (len + 10);
    if (cost < best_cost) { best_cost = cost; best = len; }
  }
  if (best > jm) best = (jm + rm::RB - 1) / rm::RB * rm::RB;
  return best;
}

}  // namespace

// *done = true when the marching kernel was launched; false when TMA cannot describe the cell arrays (the caller falls back)
cudaError_t launch_residual_march(const GridDesc& g, const SchemeConsts& c, double sqgr, bool wall, double* res, const double* w,
                                  const double* nx, const double* ny, const double* vol, const double* volf, cudaStream_t st, bool* done) {
  *done = false;
  Maps mp;
  if (!make_cell_map(g, w, 5, &mp.w) || !make_cell_map(g, vol, 1, &mp.vol) || !make_cell_map(g, volf, 2, &mp.volf)) return cudaSuccess;
  constexpr size_t SMEM = (size_t)rm::NSM * sizeof(double);
  static bool ready = false;
  static int nsm = 0;
  if (!ready) {
    cudaError_t e = cudaFuncSetAttribute(k_residual_march, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_residual_march, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    ready = true;
  }
  const int nstrips = (g.im + rm::W - 1) / rm::W;
  const int slots = 2 * nsm;
  const int seglen = pick_seglen(nstrips, g.jm, slots);
  const int nseg = (g.jm + seglen - 1) / seglen;
  const int nitems = nstrips * nseg;
  const int grid = nitems < slots ? nitems : slots;
  k_residual_march<<<grid, rm::NT_LAUNCH, SMEM, st>>>(mp, g, c, sqgr, wall, w, nx, ny, vol, volf, res, nstrips, nseg, seglen);
  *done = true;
  return cudaGetLastError();
}

}  // namespace bcast
