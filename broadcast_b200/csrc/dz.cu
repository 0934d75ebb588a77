// Spanwise (d/dz, d2/dz2) operator rows of the 3-D stability problems: entry points.
// Reference: srcfv/dz/coeffs_5p_dz.F90:10-174, coeffs_5p_dz2.F90 (f_dz.coeffs_5p_dz / coeffs_5p_dz2), called
// colour by colour from BROADCAST_npz.py:1231-1246.  Kernels: k_dz<N,WHICH> in generic_impl.cuh.
#include "../../include/broadcast_b200.h"
#include "kernels.cuh"

namespace bcast {
void count_launches(int n);
cudaError_t dz_generic_1(const GridDesc&, const SchemeArgs&, int, double*, const double*, const double*, const double*, const double*,
                         const double*, const double*, const Rect*, cudaStream_t);
cudaError_t dz_generic_5(const GridDesc&, const SchemeArgs&, int, double*, const double*, const double*, const double*, const double*,
                         const double*, const double*, const Rect*, cudaStream_t);

cudaError_t launch_dz(const GridDesc& g, const SchemeArgs& a, int which, int ndir, double* out, const double* w, const double* wd,
                      const double* nx, const double* ny, const double* vol, const double* volf, const Rect* rect, cudaStream_t st) {
  count_launches(2);
  if (ndir == 1) return dz_generic_1(g, a, which, out, w, wd, nx, ny, vol, volf, rect, st);
  if (ndir == 5) return dz_generic_5(g, a, which, out, w, wd, nx, ny, vol, volf, rect, st);
  return cudaErrorInvalidValue;
}
}  // namespace bcast

using namespace bcast;

extern "C" int bcd_dz(double* dz_out, const double* w, const double* wd, int ndir, int which, const double* nx, const double* ny,
                      const double* vol, const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz,
                      double cs, double muref, double tref, double s_suth, int im, int jm, const int32_t* rect, void* stream) {
  if (im < 1 || jm < 1 || gh != 3 || (which != 1 && which != 2) || (ndir != 1 && ndir != 5)) return BC_ERR_ARG;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const SchemeArgs a{cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, 0.0, 0.0};
  Rect rc{1, im, 1, jm};
  if (rect) rc = Rect{rect[0], rect[1], rect[2], rect[3]};
  cudaError_t e = launch_dz(g, a, which, ndir, dz_out, w, wd, nx, ny, vol, volf, &rc, (cudaStream_t)stream);
  return e == cudaSuccess ? BC_OK : (int)e;
}

static int dz_host(int which, double* dz_out, const double* w, const double* wd, const double* nx, const double* ny, const double* vol,
                   const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref,
                   double tref, double s_suth, int im, int jm) {
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
  if (!dw || !dwd || !dout || !dnx || !dny || !dvol) return BC_ERR_ALLOC;
#define CKD(call)                         \
  do {                                    \
    cudaError_t e__ = (call);             \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)
  CKD(cudaMemcpyAsync(dw, w, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  CKD(cudaMemcpyAsync(dwd, wd, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  CKD(cudaMemcpyAsync(dnx, nx, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
  CKD(cudaMemcpyAsync(dny, ny, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
  CKD(cudaMemcpyAsync(dvol, vol, sizeof(double) * g.sc, cudaMemcpyHostToDevice, 0));
  int rc = bcd_dz(dout, dw, dwd, 1, which, dnx, dny, dvol, nullptr, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, im, jm, nullptr,
                  nullptr);
  if (rc) return rc;
  // interior cells only: dz_out is intent(inout) and its ghost frame is never written by the reference
  for (int e = 0; e < 5; ++e) {
    const size_t off = (size_t)e * g.sc + g.cidx(1, 1);
    CKD(cudaMemcpy2DAsync(dz_out + off, sizeof(double) * g.ldc, dout + off, sizeof(double) * g.ldc, sizeof(double) * im, jm,
                          cudaMemcpyDeviceToHost, 0));
  }
  CKD(cudaStreamSynchronize(0));
#undef CKD
  (void)volf;
  return BC_OK;
}

extern "C" int bc_coeffs_5p_dz(double* dz_out, const double* w, const double* wd, const double* x0, const double* y0, const double* nx,
                               const double* ny, const double* xc, const double* yc, const double* vol, const double* volf, int gh,
                               double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref, double tref,
                               double s_suth, int im, int jm) {
  (void)x0; (void)y0; (void)xc; (void)yc;
  return dz_host(1, dz_out, w, wd, nx, ny, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, im, jm);
}

extern "C" int bc_coeffs_5p_dz2(double* dz_out, const double* w, const double* wd, const double* x0, const double* y0, const double* nx,
                                const double* ny, const double* xc, const double* yc, const double* vol, const double* volf, int gh,
                                double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref, double tref,
                                double s_suth, int im, int jm) {
  (void)x0; (void)y0; (void)xc; (void)yc;
  return dz_host(2, dz_out, w, wd, nx, ny, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, im, jm);
}
