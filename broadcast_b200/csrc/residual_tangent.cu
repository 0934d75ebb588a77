// k_tangent_tile: the fused residual tile algorithm of residual_fast.cuh in forward-mode tangent arithmetic (value + one tangent
// direction per scalar), one CTA per 32 x 3 tile of rows and per direction.  This is the tangent of the boundary strips of the
// Jacobian assembly and of the device colour loops (bcd_jacobian_strips, bcd_jacobian_coo): it replaces, for row-shaped
// rectangles, the chain k_prims<5> -> k_grads<5> -> k_grad_ghost -> k_strip_faces5 (global primitive / gradient arrays with
// tangents, the reference-shaped face templates in Tan<1>) by ONE kernel that stages w and wd of its tile in shared memory,
// evaluates the re-associated face formulas in dual numbers and skips every face whose stencil carries no tangent.
// Checked on the host against the reference's Tapenade tangent (tests/test_residual_tangent_host_cpu.py) and on the GPU through
// the Jacobian parity tests.  Reference: srcfv/tangent/flux_num_dnc5_d.f90:15-3870.
#define BCAST_RF_DUAL 1
#define BCAST_RF_OJ 3
#define BCAST_RF_NS rfd
#include <cstdlib>
#include "kernels.cuh"
#include "residual_fast.cuh"

namespace bcast {

namespace {

constexpr size_t TT_SMEM = (size_t)rfd::NSM * sizeof(rfd::real) + rfd::NC;   // arrays + one activity byte per staged cell

__global__ void __launch_bounds__(rfd::NT, 2)
    k_tangent_tile(GridDesc g, SchemeConsts c, double sqgr, bool wall, Rect rc, const double* __restrict__ w, const double* __restrict__ wd5,
                   const double* __restrict__ nx, const double* __restrict__ ny, const double* __restrict__ vol,
                   const double* __restrict__ volf, double* __restrict__ out5) {
  extern __shared__ __align__(16) unsigned char smraw[];
  rfd::real* sm = reinterpret_cast<rfd::real*>(smraw);
  const int dir = blockIdx.z;
  rfd::TileCtx t(g, c);
  t.wsm = sm;
  t.sm = sm + rfd::WBUF;
  t.flags = smraw + (size_t)rfd::NSM * sizeof(rfd::real);
  t.sqgr = sqgr; t.wall = wall;
  t.w = w; t.wd = wd5 + (long long)dir * 5 * g.sc;
  t.nx = nx; t.ny = ny; t.vol = vol; t.volf = volf;
  t.res = out5 + (long long)dir * 5 * g.sc;
  t.i0 = rc.i0 + blockIdx.x * rfd::OI;
  t.j0 = rc.j0 + blockIdx.y * rfd::OJ;
  t.i1 = rc.i1; t.j1 = rc.j1;
  const int tid = threadIdx.x;
  rfd::phase0<false>(t, tid);
  // a tile whose staged cells carry no tangent at all in this direction: its rows are zero
  int any = 0;
  for (int idx = tid; idx < rfd::NC; idx += rfd::NT) any |= t.flags[idx];
  if (!__syncthreads_or(any)) {
    if (rfd::owns_cell(t, tid)) {
      const long long k = g.cidx(t.i0 + tid % rfd::OI, t.j0 + tid / rfd::OI);
#pragma unroll
      for (int e = 0; e < 5; ++e) t.res[e * g.sc + k] = 0.0;
    }
    return;
  }
  rfd::phase1(t, tid, rfd::prefetch_sensor(t, tid, 0), rfd::prefetch_sensor(t, tid, 1));
  __syncthreads();
  if (t.has_ghost_sensor()) {  // CTA-uniform
    rfd::phase1b(t, tid);
    __syncthreads();
  }
  rfd::phase2(t, tid, rfd::prefetch_iface(t, tid));
  __syncthreads();
  rfd::real r[5];
#pragma unroll
  for (int e = 0; e < 5; ++e) r[e] = rfd::Dual{0.0, 0.0};
  rfd::balance_i(t, tid, r);
  rfd::phase_rj(t, tid);
  __syncthreads();
  rfd::phase3(t, tid, rfd::prefetch_jface(t, tid));
  __syncthreads();
  rfd::balance_j_store(t, tid, r);
}

}  // namespace

// tangent (five directions, out5 = [dir][e] planes) of the rows of `rc` (any height: bands of three rows)
cudaError_t launch_tangent_tile5(const GridDesc& g, const SchemeArgs& a, bool wall, const Rect& rc, double* out5, const double* w,
                                 const double* wd5, const double* nx, const double* ny, const double* vol, const double* volf,
                                 cudaStream_t st) {
  const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
  static bool ready = false;
  if (!ready) {
    cudaError_t e = cudaFuncSetAttribute(k_tangent_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TT_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_tangent_tile, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    ready = true;
  }
  const int wi = rc.i1 - rc.i0 + 1, wj = rc.j1 - rc.j0 + 1;
  if (wi < 1 || wj < 1) return cudaSuccess;
  const dim3 grid((wi + rfd::OI - 1) / rfd::OI, (wj + rfd::OJ - 1) / rfd::OJ, 5);
  k_tangent_tile<<<grid, rfd::NT, TT_SMEM, st>>>(g, c, ::sqrt(a.gam * a.rgaz), wall, rc, w, wd5, nx, ny, vol, volf, out5);
  return cudaGetLastError();
}

}  // namespace bcast
