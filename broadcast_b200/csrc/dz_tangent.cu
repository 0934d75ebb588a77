// Tangent of the spanwise operator rows w.r.t. the base flow (f_lindz.coeffs_5p_dz_d / coeffs_5p_dz2_d): kernel + entry points.
// Reference: srcfv/tangentdz/coeffs_5p_dz_d.f90, coeffs_5p_dz2_d.f90; call sites BROADCAST_npz_sens.py:1768-1797, 2157-2185.
// The algorithm (hyper-dual arithmetic, one fused 32 x 8 tile pass for both operators) is in dz_tangent.cuh.
#include "../../include/broadcast_b200.h"
#include "kernels.cuh"
#include "dz_tangent.cuh"
#include <cstdlib>

namespace bcast {
void count_launches(int n);

// MINB = CTAs per SM the register budget is cut for: 2 (118 registers, no spills) or 3 (80 registers, 34 spilled doubles, 24 warps per SM)
template <int MINB>
__global__ void __launch_bounds__(dzt::NT, MINB) k_dz_tangent(dzt::Tile t, Rect rc) {
  extern __shared__ double dzt_sm[];
  t.sm = dzt_sm;
  t.i0 = rc.i0 + blockIdx.x * dzt::TI;
  t.j0 = rc.j0 + blockIdx.y * dzt::TJ;
  t.i1 = rc.i1;
  t.j1 = rc.j1;
  const int tid = threadIdx.x;
  long long ka, kb = 0;
  const int sa = dzt::own_cell(t, tid, &ka);
  const int sb = t.out1 ? dzt::halo_cell(t, tid, &kb) : -1;   // d2/dz2 rows only: cell-local, no halo (uniform over the grid)
  const dzt::Raw ra = dzt::load_raw(t, sa >= 0, ka), rb = dzt::load_raw(t, sb >= 0, kb);
  const dzt::Met m = dzt::load_metrics(t, tid);
  dzt::phase_b(t, sb, rb);
  const dzt::Carry c = dzt::carry_of(dzt::phase_a(t, tid, ra, ka));
  if (!t.out1) return;
  __syncthreads();
  dzt::phase_c(t, tid, c, m);
}

cudaError_t launch_dz_tangent(const GridDesc& g, const dzt::Consts& c, double* out1, double* out2, const double* w, const double* wd0,
                              const double* wd, const double* nx, const double* ny, const double* vol, const Rect& rc, cudaStream_t st) {
  if (rc.i1 < rc.i0 || rc.j1 < rc.j0 || (!out1 && !out2)) return cudaSuccess;
  // BCAST_DZT_MINB = 2 | 3 selects the register budget (A/B in profiles/r1_l_summary.md); the shared-memory opt-in is per device
  static int minb = 0;
  static bool attr_set[64] = {};
  constexpr int smem = dzt::NSM * (int)sizeof(double);
  if (!minb) {
    const char* env = getenv("BCAST_DZT_MINB");
    minb = env && env[0] == '3' ? 3 : 2;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_dz_tangent<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_dz_tangent<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  dzt::Tile t{};
  t.g = g;
  t.c = c;
  t.w = w;
  t.wa = wd;
  t.wb = wd0;
  t.nx = nx;
  t.ny = ny;
  t.vol = vol;
  t.out1 = out1;
  t.out2 = out2;
  const dim3 grid((rc.i1 - rc.i0 + dzt::TI) / dzt::TI, (rc.j1 - rc.j0 + dzt::TJ) / dzt::TJ);
  count_launches(1);
  if (minb == 3)
    k_dz_tangent<3><<<grid, dzt::NT, smem, st>>>(t, rc);
  else
    k_dz_tangent<2><<<grid, dzt::NT, smem, st>>>(t, rc);
  return cudaGetLastError();
}
// for the colour loop in jacobian.cu (keeps dz_tangent.cuh out of that translation unit)
cudaError_t launch_dz_tangent_raw(const GridDesc& g, double cp, double cv, double prandtl, double gam, double cs, double muref, double tref,
                                  double s_suth, double* out1, double* out2, const double* w, const double* wd0, const double* wd,
                                  const double* nx, const double* ny, const double* vol, const Rect& rc, cudaStream_t st) {
  return launch_dz_tangent(g, dzt::make_dz_consts(cp, cv, prandtl, gam, cs, muref, tref, s_suth), out1, out2, w, wd0, wd, nx, ny, vol, rc, st);
}
}  // namespace bcast

using namespace bcast;

extern "C" int bcd_dz_tangent(double* dz_outd, double* dz2_outd, const double* w, const double* wd0, const double* wd, const double* nx,
                              const double* ny, const double* vol, int gh, double cp, double cv, double prandtl, double gam, double rgaz,
                              double cs, double muref, double tref, double s_suth, int im, int jm, const int32_t* rect, void* stream) {
  (void)rgaz;
  if (im < 1 || jm < 1 || gh != 3 || (!dz_outd && !dz2_outd)) return BC_ERR_ARG;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  Rect rc{1, im, 1, jm};
  if (rect) rc = Rect{rect[0], rect[1], rect[2], rect[3]};
  if (rc.i0 < 1 || rc.i1 > im || rc.j0 < 1 || rc.j1 > jm) return BC_ERR_ARG;
  cudaError_t e = launch_dz_tangent(g, dzt::make_dz_consts(cp, cv, prandtl, gam, cs, muref, tref, s_suth), dz_outd, dz2_outd, w, wd0, wd,
                                    nx, ny, vol, rc, (cudaStream_t)stream);
  return e == cudaSuccess ? BC_OK : (int)e;
}

// f2py-compatible call on host arrays: dz_out is left untouched, the whole of dz_outd is written (ghost frame = 0), as the
// reference does ("dz_outd = 0.0_8" before the interior loop, tangentdz/coeffs_5p_dz_d.f90, coeffs_5p_dz2_d.f90:264)
static int dz_tangent_host(int which, double* dz_outd, const double* w, const double* wd0, const double* wd, const double* nx,
                           const double* ny, const double* vol, int gh, double cp, double cv, double prandtl, double gam, double rgaz,
                           double cs, double muref, double tref, double s_suth, int im, int jm) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return BC_ERR_NODEV;
  if (im < 1 || jm < 1) return BC_ERR_ARG;
  if (gh != 3) return BC_ERR_UNSUPPORTED;
  const GridDesc g = make_grid(im, jm, gh);
  double* dw = scratch_doubles(10, g.sc * 5);
  double* dwd = scratch_doubles(11, g.sc * 5);
  double* dout = scratch_doubles(12, g.sc * 5);
  double* dnx = scratch_doubles(13, g.sn * 2);
  double* dny = scratch_doubles(14, g.sn * 2);
  double* dvol = scratch_doubles(15, g.sc);
  double* dwd0 = scratch_doubles(16, g.sc * 5);
  if (!dw || !dwd || !dout || !dnx || !dny || !dvol || !dwd0) return BC_ERR_ALLOC;
#define CKD(call)                            \
  do {                                       \
    cudaError_t e__ = (call);                \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)
  CKD(cudaMemcpyAsync(dw, w, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  CKD(cudaMemcpyAsync(dwd, wd, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  CKD(cudaMemcpyAsync(dwd0, wd0, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  CKD(cudaMemcpyAsync(dnx, nx, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
  CKD(cudaMemcpyAsync(dny, ny, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
  CKD(cudaMemcpyAsync(dvol, vol, sizeof(double) * g.sc, cudaMemcpyHostToDevice, 0));
  CKD(cudaMemsetAsync(dout, 0, sizeof(double) * g.sc * 5, 0));
  int rc = bcd_dz_tangent(which == 1 ? dout : nullptr, which == 2 ? dout : nullptr, dw, dwd0, dwd, dnx, dny, dvol, gh, cp, cv, prandtl, gam,
                          rgaz, cs, muref, tref, s_suth, im, jm, nullptr, nullptr);
  if (rc) return rc;
  CKD(cudaMemcpyAsync(dz_outd, dout, sizeof(double) * g.sc * 5, cudaMemcpyDeviceToHost, 0));
  CKD(cudaStreamSynchronize(0));
#undef CKD
  return BC_OK;
}

extern "C" int bc_coeffs_5p_dz_d(double* dz_out, double* dz_outd, const double* w, const double* wd0, const double* wd, const double* x0,
                                 const double* y0, const double* nx, const double* ny, const double* xc, const double* yc, const double* vol,
                                 const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                                 double muref, double tref, double s_suth, int im, int jm) {
  (void)dz_out; (void)x0; (void)y0; (void)xc; (void)yc; (void)volf;
  return dz_tangent_host(1, dz_outd, w, wd0, wd, nx, ny, vol, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, im, jm);
}

extern "C" int bc_coeffs_5p_dz2_d(double* dz2_out, double* dz2_outd, const double* w, const double* wd0, const double* wd, const double* x0,
                                  const double* y0, const double* nx, const double* ny, const double* xc, const double* yc,
                                  const double* vol, const double* volf, int gh, double cp, double cv, double prandtl, double gam,
                                  double rgaz, double cs, double muref, double tref, double s_suth, int im, int jm) {
  (void)dz2_out; (void)x0; (void)y0; (void)xc; (void)yc; (void)volf;
  return dz_tangent_host(2, dz2_outd, w, wd0, wd, nx, ny, vol, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, im, jm);
}
