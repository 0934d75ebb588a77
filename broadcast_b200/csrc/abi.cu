// extern "C" entry points of libbroadcast_b200 (see include/broadcast_b200.h for the contract and
// the reference file:line each one replaces).
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/broadcast_b200.h"
#include "kernels.cuh"

using namespace bcast;

namespace {
thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char* what) {
  g_err = what;
  return code;
}
int cuda_fail(cudaError_t e, const char* where) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  return (int)e;
}
#define CK(call)                                  \
  do {                                            \
    cudaError_t e__ = (call);                     \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

int check_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) return fail(BC_ERR_NODEV, "no CUDA device available (broadcast_b200 has no CPU fallback)");
  return BC_OK;
}
// gh = 2 / 3 / 4 / 5: the scheme family flux_num_dnc3 / 5 / 7 / 9 (gh = (order + 1) / 2, BROADCAST_npz.py:501-502).  The boundary fills,
// seeds, scatters, norms, the residual and the tangent exist for every member; the fused kernels and the block Jacobian are order 5.
int check_dims(int im, int jm, int gh) {
  if (im < 1 || jm < 1) return fail(BC_ERR_ARG, "im and jm must be positive");
  if (gh < 2 || gh > 5) return fail(BC_ERR_UNSUPPORTED, "gh must be 2, 3, 4 or 5 (schemes flux_num_dnc3 / 5 / 7 / 9)");
  return BC_OK;
}
int check_order5(int gh) {
  if (gh != 3) return fail(BC_ERR_UNSUPPORTED, "this entry point exists for the order-5 scheme (gh = 3) only");
  return BC_OK;
}
// the scheme routine named `order` on a block with `gh` ghost layers: the off-centred wall rows read the cell rows 1 .. NP
int check_scheme(int order, int gh, int jm, bool wall) {
  if (gh != (order + 1) / 2) return fail(BC_ERR_UNSUPPORTED, "the dnc schemes run with gh = (order + 1) / 2 ghost layers");
  const int np = order == 3 ? 2 : order == 5 ? 5 : order == 7 ? 9 : 11;
  if (wall && jm + gh < np) return fail(BC_ERR_ARG, "jm too small for the wall rows of this order");
  return BC_OK;
}

// host-API device buffers (scratch slots >= 100: disjoint from the slots the device loops use, 0-3 / 20-22 / 30 / 40)
enum { S_W = 100, S_WD, S_RES, S_NX, S_NY, S_VOL, S_VOLF, S_AUX, S_AUX2, S_SEG, S_IA, S_JA, S_OUT10, S_GEOM };

template <class T>
T* dbuf(int slot, size_t count) {
  return reinterpret_cast<T*>(scratch_doubles(slot, (count * sizeof(T) + 7) / 8));
}

SchemeArgs sargs(double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth,
                 double k2, double k4) {
  return SchemeArgs{cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4};
}

struct Geo {
  const double *nx, *ny, *vol, *volf;
};
int upload_geo(const GridDesc& g, const double* nx, const double* ny, const double* vol, const double* volf, Geo& d) {
  double* dnx = dbuf<double>(S_NX, g.sn * 2);
  double* dny = dbuf<double>(S_NY, g.sn * 2);
  double* dvol = dbuf<double>(S_VOL, g.sc);
  double* dvolf = dbuf<double>(S_VOLF, g.sc * 2);
  if (!dnx || !dny || !dvol || !dvolf) return fail(BC_ERR_ALLOC, "device allocation failed");
  CK(cudaMemcpyAsync(dnx, nx, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(dny, ny, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(dvol, vol, sizeof(double) * g.sc, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(dvolf, volf, sizeof(double) * g.sc * 2, cudaMemcpyHostToDevice, 0));
  d = Geo{dnx, dny, dvol, dvolf};
  return BC_OK;
}

int residual_host(bool wall, double* residu, const double* w, const double* nx, const double* ny, const double* vol, const double* volf,
                  int gh, const SchemeArgs& a, int im, int jm) {
  if (int rc = check_device()) return rc;
  if (int rc = check_dims(im, jm, gh)) return rc;
  const GridDesc g = make_grid(im, jm, gh);
  Geo geo;
  if (int rc = upload_geo(g, nx, ny, vol, volf, geo)) return rc;
  double* dw = dbuf<double>(S_W, g.sc * 5);
  double* dres = dbuf<double>(S_RES, g.sc * 5);
  if (!dw || !dres) return fail(BC_ERR_ALLOC, "device allocation failed");
  CK(cudaMemcpyAsync(dw, w, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  const bool generic = getenv("BROADCAST_B200_GENERIC") != nullptr;
  int rc = bcd_residual(dres, dw, geo.nx, geo.ny, geo.vol, geo.volf, gh, a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref,
                        a.s_suth, a.k2, a.k4, im, jm, wall ? 1 : 0, generic ? 1 : 0, nullptr);
  if (rc) return rc;
  // interior cells only: ghosts of residu are left untouched (intent(inout) array never written there)
  for (int e = 0; e < 5; ++e) {
    const size_t off = (size_t)e * g.sc + g.cidx(1, 1);
    CK(cudaMemcpy2DAsync(residu + off, sizeof(double) * g.ldc, dres + off, sizeof(double) * g.ldc, sizeof(double) * im, jm,
                         cudaMemcpyDeviceToHost, 0));
  }
  CK(cudaStreamSynchronize(0));
  return BC_OK;
}

int tangent_host(bool wall, double* residud, const double* w, const double* wd, const double* nx, const double* ny, const double* vol,
                 const double* volf, int gh, const SchemeArgs& a, int im, int jm) {
  if (int rc = check_device()) return rc;
  if (int rc = check_dims(im, jm, gh)) return rc;
  const GridDesc g = make_grid(im, jm, gh);
  Geo geo;
  if (int rc = upload_geo(g, nx, ny, vol, volf, geo)) return rc;
  double* dw = dbuf<double>(S_W, g.sc * 5);
  double* dwd = dbuf<double>(S_WD, g.sc * 5);
  double* dres = dbuf<double>(S_RES, g.sc * 5);
  if (!dw || !dwd || !dres) return fail(BC_ERR_ALLOC, "device allocation failed");
  CK(cudaMemcpyAsync(dw, w, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(dwd, wd, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  CK(cudaMemsetAsync(dres, 0, sizeof(double) * g.sc * 5, 0));
  int rc = bcd_tangent(dres, dw, dwd, 1, geo.nx, geo.ny, geo.vol, geo.volf, gh, a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref,
                       a.tref, a.s_suth, a.k2, a.k4, im, jm, wall ? 1 : 0, nullptr, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(residud, dres, sizeof(double) * g.sc * 5, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return BC_OK;
}

// boundary fill on host arrays: upload w (and wd), run, download w (and wd)
template <class F>
int bc_host(double* w, double* wd, int gh, int im, int jm, F&& run) {
  if (int rc = check_device()) return rc;
  if (int rc = check_dims(im, jm, gh)) return rc;
  const GridDesc g = make_grid(im, jm, gh);
  double* dw = dbuf<double>(S_W, g.sc * 5);
  double* dwd = wd ? dbuf<double>(S_WD, g.sc * 5) : nullptr;
  if (!dw || (wd && !dwd)) return fail(BC_ERR_ALLOC, "device allocation failed");
  CK(cudaMemcpyAsync(dw, w, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  if (wd) CK(cudaMemcpyAsync(dwd, wd, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  if (int rc = run(g, dw, dwd)) return rc;
  CK(cudaMemcpyAsync(w, dw, sizeof(double) * g.sc * 5, cudaMemcpyDeviceToHost, 0));
  if (wd) CK(cudaMemcpyAsync(wd, dwd, sizeof(double) * g.sc * 5, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return BC_OK;
}

int scatter_host(int kind, double* jac, int32_t* ia, int32_t* ja, const double* resd, int m, int l, int k, int gh, int im, int jm,
                 int64_t nbentry, const double* coefdiag, const double* vol) {
  if (int rc = check_device()) return rc;
  if (int rc = check_dims(im, jm, gh)) return rc;
  const int s = 2 * gh + 1;
  if (m < 0 || m > 4 || l < 0 || l >= s || k < 0 || k >= s) return fail(BC_ERR_ARG, "colour indices out of range");
  const GridDesc g = make_grid(im, jm, gh);
  const long long n = 5LL * im * jm;
  const long long base = (long long)k * n + (long long)l * n * s + (long long)m * n * s * s;
  if (base + n > nbentry) return fail(BC_ERR_ARG, "jac/ia/ja too short for this colour");
  double* dres = dbuf<double>(S_RES, g.sc * 5);
  double* dseg = dbuf<double>(S_SEG, n);
  int* dia = dbuf<int>(S_IA, n);
  int* dja = dbuf<int>(S_JA, n);
  double* dcoef = coefdiag ? dbuf<double>(S_AUX, (size_t)im * jm) : nullptr;
  double* dvol = vol ? dbuf<double>(S_VOL, g.sc) : nullptr;
  if (!dres || !dseg || !dia || !dja) return fail(BC_ERR_ALLOC, "device allocation failed");
  CK(cudaMemcpyAsync(dres, resd, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  if (coefdiag) CK(cudaMemcpyAsync(dcoef, coefdiag, sizeof(double) * im * jm, cudaMemcpyHostToDevice, 0));
  if (vol) CK(cudaMemcpyAsync(dvol, vol, sizeof(double) * g.sc, cudaMemcpyHostToDevice, 0));
  if (int rc = bcd_scatter(kind, dseg, dia, dja, dres, m, l, k, gh, im, jm, dcoef, dvol, nullptr)) return rc;
  CK(cudaMemcpyAsync(jac + base, dseg, sizeof(double) * n, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(ia + base, dia, sizeof(int) * n, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(ja + base, dja, sizeof(int) * n, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return BC_OK;
}

int norms_host(double* s2, double* s10, const double* rhs, int im, int jm, int gh) {
  if (int rc = check_device()) return rc;
  if (int rc = check_dims(im, jm, gh)) return rc;
  const GridDesc g = make_grid(im, jm, gh);
  double* dres = dbuf<double>(S_RES, g.sc * 5);
  double* dout = dbuf<double>(S_OUT10, 16);
  if (!dres || !dout) return fail(BC_ERR_ALLOC, "device allocation failed");
  CK(cudaMemcpyAsync(dres, rhs, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  if (int rc = bcd_norm_sums(dout, dres, im, jm, gh, nullptr)) return rc;
  double h[10];
  CK(cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost));
  for (int e = 0; e < 5; ++e) {
    s2[e] = h[e];
    s10[e] = h[5 + e];
  }
  return BC_OK;
}
}  // namespace

namespace bcast {
void count_launches(int n) { g_launches += n; }
}  // namespace bcast

extern "C" {

int bcd_slab_begin(int ioff, int im_global, int edges) {
  if (ioff < 0 || im_global < 1 || edges < 0 || edges > 3) return fail(BC_ERR_ARG, "bad slab descriptor");
  current_slab() = SlabInfo{ioff, im_global, edges};
  return BC_OK;
}
// strided copy between a HOST array and a device array (columns of a Fortran-ordered block = pitched rows):
// width bytes per row, `height` rows; kind 1 = host -> device, 2 = device -> host; asynchronous on `stream`
int bcd_memcpy2d(void* dst, long long dpitch, const void* src, long long spitch, long long width, long long height, int kind,
                 void* stream) {
  if (kind != 1 && kind != 2) return fail(BC_ERR_ARG, "kind must be 1 (H2D) or 2 (D2H)");
  cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height,
                                    kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_memcpy2d");
}
// colour sharding: the colour loops of this thread (bcd_jacobian_coo, bcd_jacobian_strips, bcd_dz_coo) visit only the passes
// c0 <= l * (2gh+1) + k < c1; c1 <= c0 restores all colours.  Slots of the other colours keep their content.
int bcd_colour_range(int c0, int c1) {
  if (c0 < 0 || c1 < 0) return fail(BC_ERR_ARG, "bad colour range");
  current_colours() = ColourRange{c0, c1};
  return BC_OK;
}
int bcd_slab_end(void) {
  current_slab() = SlabInfo{0, 0, 0};
  return BC_OK;
}

int bc_set_bndbl_2d(const double* w, double* field, double* wbd, int im, int jm, int gh) {
  if (im < 1 || jm < 1 || gh < 1) return fail(BC_ERR_ARG, "im, jm, gh must be positive");
  const GridDesc g = make_grid(im, jm, gh);
  for (int e = 0; e < 5; ++e) {
    for (int depth = 1; depth <= gh; ++depth)
      for (int j = 1; j <= jm; ++j) field[(j - 1) + (size_t)(depth - 1) * jm + (size_t)e * jm * gh] = w[e * g.sc + g.cidx(1 - depth, j)];
    for (int i = 1; i <= im + gh; ++i) wbd[(i - 1) + (size_t)e * (im + gh)] = w[e * g.sc + g.cidx(i - gh, jm)];
  }
  return BC_OK;
}

int bc_version(void) { return 100; }
int bc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
const char* bc_last_error(void) { return g_err.c_str(); }
long long bc_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------ device API
int bcd_residual(double* residu, const double* w, const double* nx, const double* ny, const double* vol, const double* volf, int gh,
                 double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth,
                 double k2, double k4, int im, int jm, int wall, int use_generic, void* stream) {
  if (int rc = check_dims(im, jm, gh)) return rc;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const SchemeArgs a = sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4);
  if (int rc = check_scheme(2 * gh - 1, gh, jm, wall != 0)) return rc;
  cudaError_t e;
  if (use_generic == RES_GENERIC || gh != 3) {   // orders 3 / 7 / 9: the reference-shaped pipeline on the order's stencil tables
    e = launch_residual_generic(g, a, wall != 0, 0, residu, w, nullptr, nx, ny, vol, volf, nullptr, (cudaStream_t)stream);
    g_launches += 4;
  } else {
    e = launch_residual_tiled(g, a, wall != 0, residu, w, nx, ny, vol, volf, (cudaStream_t)stream, use_generic);
    g_launches += 1;
  }
  if (e != cudaSuccess) return cuda_fail(e, "bcd_residual");
  return BC_OK;
}

int bcd_residual_part(double* residu, const double* w, const double* nx, const double* ny, const double* vol, const double* volf, int gh,
                      double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth,
                      double k2, double k4, int im, int jm, int wall, int part, void* stream) {
  if (int rc = check_dims(im, jm, gh)) return rc;
  if (int rc = check_order5(gh)) return rc;
  if (part < 0 || part > 2) return fail(BC_ERR_ARG, "part must be 0 (all), 1 (inner tiles) or 2 (ring of tiles)");
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const SchemeArgs a = sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4);
  cudaError_t e = launch_residual_tiled(g, a, wall != 0, residu, w, nx, ny, vol, volf, (cudaStream_t)stream, RES_DEFAULT, part);
  g_launches += 1;
  if (e != cudaSuccess) return cuda_fail(e, "bcd_residual_part");
  return BC_OK;
}

int bcd_tangent(double* residud, const double* w, const double* wd, int ndir, const double* nx, const double* ny, const double* vol,
                const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref,
                double tref, double s_suth, double k2, double k4, int im, int jm, int wall, const int32_t* rect, void* stream) {
  if (int rc = check_dims(im, jm, gh)) return rc;
  if (ndir != 1 && ndir != 5) return fail(BC_ERR_ARG, "ndir must be 1 or 5");
  if (int rc = check_scheme(2 * gh - 1, gh, jm, wall != 0)) return rc;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const SchemeArgs a = sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4);
  Rect rc{1, im, 1, jm};
  if (rect) rc = Rect{rect[0], rect[1], rect[2], rect[3]};
  cudaError_t e = launch_residual_generic(g, a, wall != 0, ndir, residud, w, wd, nx, ny, vol, volf, &rc, (cudaStream_t)stream);
  g_launches += 5;
  if (e != cudaSuccess) return cuda_fail(e, "bcd_tangent");
  return BC_OK;
}

#define BCD_PROLOGUE()                                                       \
  if (int rc = check_dims(im, jm, gh)) return rc;                            \
  if (ndir != 0 && ndir != 1 && ndir != 5) return fail(BC_ERR_ARG, "ndir must be 0, 1 or 5"); \
  if (ndir && !wd) return fail(BC_ERR_ARG, "wd is null");                    \
  const GridDesc g = make_grid(im, jm, gh);                                  \
  BcLine b;                                                                  \
  if (!loc || !decode_interface(loc, interf, b)) return fail(BC_ERR_ARG, "loc must be Ilo, Ihi, Jlo or Jhi")

int bcd_bc_wall_viscous_adia(double* w, double* wd, int ndir, const char* loc, double gam, const int32_t* interf, int gh, int im, int jm,
                             void* stream) {
  BCD_PROLOGUE();
  cudaError_t e = launch_bc_wall(g, b, gam, ndir, w, wd, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_bc_wall_viscous_adia");
}
int bcd_bc_wall_viscous_iso(double* w, double* wd, int ndir, double twall, const char* loc, double gam, double rgaz, const int32_t* interf,
                            int gh, int im, int jm, void* stream) {
  BCD_PROLOGUE();
  cudaError_t e = launch_bc_wall_iso(g, b, twall, gam, rgaz, ndir, w, wd, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_bc_wall_viscous_iso");
}
int bcd_bc_symmetry(double* w, double* wd, int ndir, const char* loc, const int32_t* interf, const double* nx, const double* ny, int gh,
                    int im, int jm, int anti, void* stream) {
  BCD_PROLOGUE();
  if (!nx || !ny) return fail(BC_ERR_ARG, "nx / ny is null");
  cudaError_t e = launch_bc_symmetry(g, b, ndir, w, wd, nx, ny, anti != 0, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_bc_symmetry");
}
int bcd_bc_wall_profile(double* w, double* wd, int ndir, int blow, const double* prof, const double* profd, const char* loc, double gam,
                        double gamd, double rgaz, double rgazd, const int32_t* interf, int gh, int im, int jm, int lm, void* stream) {
  BCD_PROLOGUE();
  if (!prof) return fail(BC_ERR_ARG, "profile is null");
  if (lm < b.lmax) return fail(BC_ERR_ARG, "profile shorter than the boundary line");
  cudaError_t e = launch_bc_wall_profile(g, b, blow != 0, prof, profd, gam, gamd, rgaz, rgazd, ndir, w, wd, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_bc_wall_profile");
}
int bcd_bc_pressure(double* w, double* wd, int ndir, const char* loc, const int32_t* interf, double pext, int noref, double gam,
                    const double* nx, const double* ny, int im, int jm, int gh, void* stream) {
  BCD_PROLOGUE();
  if (!nx || !ny) return fail(BC_ERR_ARG, "nx / ny is null");
  cudaError_t e = launch_bc_pressure(g, b, pext, noref != 0, gam, ndir, w, wd, nx, ny, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_bc_pressure");
}
int bcd_bc_no_reflexion(double* w, double* wd, int ndir, const double* wbd, const char* loc, const int32_t* interf, const double* nx,
                        const double* ny, double gam, int gh, int im, int jm, int lm, void* stream) {
  BCD_PROLOGUE();
  if (lm < b.lmax) return fail(BC_ERR_ARG, "wbd has fewer rows than the interface");
  cudaError_t e = launch_bc_noref(g, b, gam, ndir, w, wd, wbd, lm, nx, ny, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_bc_no_reflexion");
}
int bcd_bc_supandsubinlet(double* w, double* wd, int ndir, const char* loc, const int32_t* interf, const double* field, const double* nx,
                          const double* ny, double gam, int im, int jm, int lm, int gh, void* stream) {
  BCD_PROLOGUE();
  if (lm < b.lmax) return fail(BC_ERR_ARG, "field has fewer rows than the interface");
  cudaError_t e = launch_bc_inlet(g, b, gam, ndir, w, wd, field, lm, nx, ny, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_bc_supandsubinlet");
}
int bcd_bc_extrapolate_o2(double* w, double* wd, int ndir, const char* loc, const int32_t* interf, int im, int jm, int gh, void* stream) {
  BCD_PROLOGUE();
  cudaError_t e = launch_bc_extrap(g, b, ndir, w, wd, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_bc_extrapolate_o2");
}

int bcd_bc_general(double* w, double* wd, int ndir, const char* loc, const int32_t* interf, const double* field, int gh, int im, int jm,
                   int lm, void* stream) {
  BCD_PROLOGUE();
  if (!field || lm < b.lmax) return fail(BC_ERR_ARG, "field is null or has fewer rows than the interface");
  cudaError_t e = launch_bc_general(g, b, ndir, w, wd, field, lm, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_bc_general");
}

int bcd_jn_match(double* wr, const int32_t* prr, int gh1r, int gh2r, int gh3r, int gh4r, int imr, int jmr, const double* wd,
                 const int32_t* prd, int gh1d, int gh2d, int gh3d, int gh4d, int imd, int jmd, const int32_t* tr, int em, void* stream) {
  if (!prr || !prd || !tr || em < 1) return fail(BC_ERR_ARG, "jn_match: null window or em < 1");
  Window r{imr + gh1r + gh2r, (long long)(imr + gh1r + gh2r) * (jmr + gh3r + gh4r), 1 - gh1r, 1 - gh3r};
  Window d{imd + gh1d + gh2d, (long long)(imd + gh1d + gh2d) * (jmd + gh3d + gh4d), 1 - gh1d, 1 - gh3d};
  cudaError_t e = launch_jn_match(wr, r, prr, wd, d, prd, tr, em, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_jn_match");
}

int bcd_testvector(double* wd, int ndir, int m, int l, int k, int gh, int im, int jm, const int32_t* zone, void* stream) {
  if (int rc = check_dims(im, jm, gh)) return rc;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  cudaError_t e = launch_testvector(g, wd, ndir, m, l, k, zone, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_testvector");
}

int bcd_scatter(int kind, double* seg_jac, int32_t* seg_ia, int32_t* seg_ja, const double* resd, int m, int l, int k, int gh, int im,
                int jm, const double* coefdiag, const double* vol, void* stream) {
  if (int rc = check_dims(im, jm, gh)) return rc;
  if (kind < 0 || kind > 6) return fail(BC_ERR_ARG, "unknown scatter kind");
  if ((kind == SCATTER_JV_RELAXED || kind == SCATTER_JV_RELAXED_JN || kind == SCATTER_JV_RELAXED_DBYVOL) && !coefdiag)
    return fail(BC_ERR_ARG, "coefdiag is null");
  if ((kind == SCATTER_JV_DBYVOL || kind == SCATTER_JV_RELAXED_DBYVOL) && !vol) return fail(BC_ERR_ARG, "vol is null");
  const GridDesc g = make_grid_ctx(im, jm, gh);
  cudaError_t e = launch_scatter(g, kind, seg_jac, seg_ia, seg_ja, resd, m, l, k, coefdiag, vol, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_scatter");
}

int bcd_norm_sums(double* out10, const double* rhs, int im, int jm, int gh, void* stream) {
  if (int rc = check_dims(im, jm, gh)) return rc;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  cudaError_t e = launch_norms(g, rhs, out10, (cudaStream_t)stream);
  g_launches += 1;
  return e == cudaSuccess ? BC_OK : cuda_fail(e, "bcd_norm_sums");
}

// ------------------------------------------------------------------------------------------ host API
int bc_flux_num_dnc5_2d(double* residu, const double* w, const double*, const double*, const double* nx, const double* ny, const double*,
                        const double*, const double* vol, const double* volf, int gh, double cp, double cv, double prandtl, double gam,
                        double rgaz, double cs, double muref, double tref, double s_suth, double k2, double k4, int im, int jm) {
  if (int rc = check_scheme(5, gh, jm, true)) return rc;
  return residual_host(true, residu, w, nx, ny, vol, volf, gh, sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4), im, jm);
}
int bc_flux_num_dnc5_nowall_2d(double* residu, const double* w, const double*, const double*, const double* nx, const double* ny,
                               const double*, const double*, const double* vol, const double* volf, int gh, double cp, double cv,
                               double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth, double k2,
                               double k4, int im, int jm) {
  if (int rc = check_scheme(5, gh, jm, false)) return rc;
  return residual_host(false, residu, w, nx, ny, vol, volf, gh, sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4), im, jm);
}
int bc_flux_num_dnc5_2d_d(double*, double* residud, const double* w, const double* wd, const double*, const double*, const double* nx,
                          const double* ny, const double*, const double*, const double* vol, const double* volf, int gh, double cp,
                          double cv, double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth,
                          double k2, double k4, int im, int jm) {
  if (int rc = check_scheme(5, gh, jm, true)) return rc;
  return tangent_host(true, residud, w, wd, nx, ny, vol, volf, gh, sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4), im,
                      jm);
}
int bc_flux_num_dnc5_nowall_2d_d(double*, double* residud, const double* w, const double* wd, const double*, const double*,
                                 const double* nx, const double* ny, const double*, const double*, const double* vol, const double* volf,
                                 int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref,
                                 double tref, double s_suth, double k2, double k4, int im, int jm) {
  if (int rc = check_scheme(5, gh, jm, false)) return rc;
  return tangent_host(false, residud, w, wd, nx, ny, vol, volf, gh, sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4), im,
                      jm);
}

// the other orders of the family (srcfv/rhs/flux_num_dnc{3,7,9}.F90, _nowall variants, srcfv/tangent/flux_num_dnc{3,7,9}_d.f90): same
// argument lists; the routine's order must agree with the ghost depth of the arrays it is handed
#define BCAST_SCHEME_ENTRIES(ORD)                                                                                                          \
  int bc_flux_num_dnc##ORD##_2d(double* residu, const double* w, const double*, const double*, const double* nx, const double* ny,         \
                                const double*, const double*, const double* vol, const double* volf, int gh, double cp, double cv,       \
                                double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth, double k2,   \
                                double k4, int im, int jm) {                                                                              \
    if (int rc = check_scheme(ORD, gh, jm, true)) return rc;                                                                              \
    return residual_host(true, residu, w, nx, ny, vol, volf, gh, sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4), im,  \
                         jm);                                                                                                             \
  }                                                                                                                                       \
  int bc_flux_num_dnc##ORD##_nowall_2d(double* residu, const double* w, const double*, const double*, const double* nx, const double* ny,  \
                                       const double*, const double*, const double* vol, const double* volf, int gh, double cp, double cv, \
                                       double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth,      \
                                       double k2, double k4, int im, int jm) {                                                            \
    if (int rc = check_scheme(ORD, gh, jm, false)) return rc;                                                                             \
    return residual_host(false, residu, w, nx, ny, vol, volf, gh, sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4), im, \
                         jm);                                                                                                             \
  }                                                                                                                                       \
  int bc_flux_num_dnc##ORD##_2d_d(double*, double* residud, const double* w, const double* wd, const double*, const double*,              \
                                  const double* nx, const double* ny, const double*, const double*, const double* vol,                    \
                                  const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,   \
                                  double muref, double tref, double s_suth, double k2, double k4, int im, int jm) {                       \
    if (int rc = check_scheme(ORD, gh, jm, true)) return rc;                                                                              \
    return tangent_host(true, residud, w, wd, nx, ny, vol, volf, gh, sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4),  \
                        im, jm);                                                                                                          \
  }                                                                                                                                       \
  int bc_flux_num_dnc##ORD##_nowall_2d_d(double*, double* residud, const double* w, const double* wd, const double*, const double*,       \
                                         const double* nx, const double* ny, const double*, const double*, const double* vol,             \
                                         const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz,       \
                                         double cs, double muref, double tref, double s_suth, double k2, double k4, int im, int jm) {     \
    if (int rc = check_scheme(ORD, gh, jm, false)) return rc;                                                                             \
    return tangent_host(false, residud, w, wd, nx, ny, vol, volf, gh, sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4), \
                        im, jm);                                                                                                          \
  }
BCAST_SCHEME_ENTRIES(3)
BCAST_SCHEME_ENTRIES(7)
BCAST_SCHEME_ENTRIES(9)
#undef BCAST_SCHEME_ENTRIES

// isothermal-wall variant of the scheme (flux_num_dnc5_iso.F90, tangent/flux_num_dnc5_iso_d.f90): the same kernels with the wall
// flux of rhs/fluxwall_iso.F, selected through the calling thread's wall context for the duration of the call
struct WallIsoScope {
  WallIso saved;
  WallIsoScope(double twall) : saved(current_wall_iso()) { current_wall_iso() = WallIso{1, twall}; }
  ~WallIsoScope() { current_wall_iso() = saved; }
};
int bc_flux_num_dnc5_iso_2d(double* residu, const double* w, double twall, const double*, const double*, const double* nx,
                            const double* ny, const double*, const double*, const double* vol, const double* volf, int gh, double cp,
                            double cv, double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth,
                            double k2, double k4, int im, int jm) {
  if (int rc = check_scheme(5, gh, jm, true)) return rc;
  WallIsoScope scope(twall);
  return residual_host(true, residu, w, nx, ny, vol, volf, gh, sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4), im, jm);
}
int bc_flux_num_dnc5_iso_2d_d(double*, double* residud, const double* w, const double* wd, double twall, const double*, const double*,
                              const double* nx, const double* ny, const double*, const double*, const double* vol, const double* volf,
                              int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref,
                              double tref, double s_suth, double k2, double k4, int im, int jm) {
  if (int rc = check_scheme(5, gh, jm, true)) return rc;
  WallIsoScope scope(twall);
  return tangent_host(true, residud, w, wd, nx, ny, vol, volf, gh, sargs(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4), im,
                      jm);
}
// resident mode: the wall flux of the calling thread's next bcd_* launches (on != 0: rhs/fluxwall_iso.F with this twall)
int bcd_wall_iso(int on, double twall) {
  current_wall_iso() = WallIso{on ? 1 : 0, on ? twall : 0.0};
  return BC_OK;
}

int bc_bc_wall_viscous_adia_2d(double* w, const char* loc, double gam, const int32_t* interf, int gh, int im, int jm) {
  return bc_host(w, nullptr, gh, im, jm,
                 [&](const GridDesc&, double* dw, double*) { return bcd_bc_wall_viscous_adia(dw, nullptr, 0, loc, gam, interf, gh, im, jm, nullptr); });
}
int bc_bc_wall_viscous_adia_2d_d(double* w, double* wd, const char* loc, double gam, const int32_t* interf, int gh, int im, int jm) {
  if (!wd) return fail(BC_ERR_ARG, "wd is null");
  return bc_host(w, wd, gh, im, jm,
                 [&](const GridDesc&, double* dw, double* dwd) { return bcd_bc_wall_viscous_adia(dw, dwd, 1, loc, gam, interf, gh, im, jm, nullptr); });
}

int bc_bc_wall_viscous_iso_2d(double* w, double twall, const char* loc, double gam, double rgaz, const int32_t* interf, int gh, int im,
                              int jm) {
  return bc_host(w, nullptr, gh, im, jm, [&](const GridDesc&, double* dw, double*) {
    return bcd_bc_wall_viscous_iso(dw, nullptr, 0, twall, loc, gam, rgaz, interf, gh, im, jm, nullptr);
  });
}
int bc_bc_wall_viscous_iso_2d_d(double* w, double* wd, double twall, const char* loc, double gam, double rgaz, const int32_t* interf, int gh,
                                int im, int jm) {
  if (!wd) return fail(BC_ERR_ARG, "wd is null");
  return bc_host(w, wd, gh, im, jm, [&](const GridDesc&, double* dw, double* dwd) {
    return bcd_bc_wall_viscous_iso(dw, dwd, 1, twall, loc, gam, rgaz, interf, gh, im, jm, nullptr);
  });
}

static int symmetry_host(double* w, double* wd, const char* loc, const int32_t* interf, const double* nx, const double* ny, int gh, int im,
                         int jm, int anti = 0) {
  return bc_host(w, wd, gh, im, jm, [&](const GridDesc& g, double* dw, double* dwd) -> int {
    double* dnx = dbuf<double>(S_NX, g.sn * 2);
    double* dny = dbuf<double>(S_NY, g.sn * 2);
    if (!dnx || !dny) return fail(BC_ERR_ALLOC, "device allocation failed");
    CK(cudaMemcpyAsync(dnx, nx, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(dny, ny, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
    return bcd_bc_symmetry(dw, dwd, wd ? 1 : 0, loc, interf, dnx, dny, gh, im, jm, anti, nullptr);
  });
}
int bc_bc_antisymmetry_2d(double* w, const char* loc, const int32_t* interf, const double* nx, const double* ny, int gh, int im, int jm) {
  return symmetry_host(w, nullptr, loc, interf, nx, ny, gh, im, jm, 1);
}
int bc_bc_antisymmetry_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, const double* nx, const double* ny, int gh,
                            int im, int jm) {
  if (!wd) return fail(BC_ERR_ARG, "wd is null");
  return symmetry_host(w, wd, loc, interf, nx, ny, gh, im, jm, 1);
}
static int wall_profile_host(int blow, double* w, double* wd, const double* prof, const double* profd, const char* loc, double gam,
                             double gamd, double rgaz, double rgazd, const int32_t* interf, int gh, int im, int jm, int lm) {
  if (lm < 1 || !prof) return fail(BC_ERR_ARG, "profile / lm");
  return bc_host(w, wd, gh, im, jm, [&](const GridDesc&, double* dw, double* dwd) -> int {
    double* dprof = dbuf<double>(S_AUX, (size_t)lm * 2);
    if (!dprof) return fail(BC_ERR_ALLOC, "device allocation failed");
    CK(cudaMemcpyAsync(dprof, prof, sizeof(double) * lm, cudaMemcpyHostToDevice, 0));
    if (profd) CK(cudaMemcpyAsync(dprof + lm, profd, sizeof(double) * lm, cudaMemcpyHostToDevice, 0));
    return bcd_bc_wall_profile(dw, dwd, wd ? 1 : 0, blow, dprof, profd ? dprof + lm : nullptr, loc, gam, gamd, rgaz, rgazd, interf, gh, im, jm,
                               lm, nullptr);
  });
}
int bc_bc_wall_blow_profile_2d(double* w, const double* velprof, const char* loc, double gam, const int32_t* interf, int gh, int im, int jm,
                               int lm) {
  return wall_profile_host(1, w, nullptr, velprof, nullptr, loc, gam, 0.0, 1.0, 0.0, interf, gh, im, jm, lm);
}
int bc_bc_wall_blow_profile_2d_d(double* w, double* wd, const double* velprof, const double* velprofd, const char* loc, double gam,
                                 double gamd, const int32_t* interf, int gh, int im, int jm, int lm) {
  if (!wd || !velprofd) return fail(BC_ERR_ARG, "wd / velprofd is null");
  return wall_profile_host(1, w, wd, velprof, velprofd, loc, gam, gamd, 1.0, 0.0, interf, gh, im, jm, lm);
}
int bc_bc_wall_viscous_iso_profile_2d(double* w, const double* twallprof, const char* loc, double gam, double rgaz, const int32_t* interf,
                                      int gh, int im, int jm, int lm) {
  return wall_profile_host(0, w, nullptr, twallprof, nullptr, loc, gam, 0.0, rgaz, 0.0, interf, gh, im, jm, lm);
}
int bc_bc_wall_viscous_iso_profile_2d_d(double* w, double* wd, const double* twallprof, const double* twallprofd, const char* loc,
                                        double gam, double gamd, double rgaz, double rgazd, const int32_t* interf, int gh, int im, int jm,
                                        int lm) {
  if (!wd || !twallprofd) return fail(BC_ERR_ARG, "wd / twallprofd is null");
  return wall_profile_host(0, w, wd, twallprof, twallprofd, loc, gam, gamd, rgaz, rgazd, interf, gh, im, jm, lm);
}
static int pressure_host(double* w, double* wd, const char* loc, const int32_t* interf, double pext, int noref, double gam, const double* nx,
                         const double* ny, int im, int jm, int gh, int em) {
  if (em != 5) return fail(BC_ERR_UNSUPPORTED, "bc_pressure_2d: em must be 5 (no transported scalars on this path)");
  return bc_host(w, wd, gh, im, jm, [&](const GridDesc& g, double* dw, double* dwd) -> int {
    double* dnx = dbuf<double>(S_NX, g.sn * 2);
    double* dny = dbuf<double>(S_NY, g.sn * 2);
    if (!dnx || !dny) return fail(BC_ERR_ALLOC, "device allocation failed");
    CK(cudaMemcpyAsync(dnx, nx, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(dny, ny, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
    return bcd_bc_pressure(dw, dwd, wd ? 1 : 0, loc, interf, pext, noref, gam, dnx, dny, im, jm, gh, nullptr);
  });
}
int bc_bc_pressure_2d(double* w, const char* loc, const int32_t* interf, double pext, int noref, double gam, const double* nx,
                      const double* ny, int im, int jm, int gh, int em) {
  return pressure_host(w, nullptr, loc, interf, pext, noref, gam, nx, ny, im, jm, gh, em);
}
int bc_bc_pressure_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, double pext, int noref, double gam, const double* nx,
                        const double* ny, int im, int jm, int gh, int em) {
  if (!wd) return fail(BC_ERR_ARG, "wd is null");
  return pressure_host(w, wd, loc, interf, pext, noref, gam, nx, ny, im, jm, gh, em);
}
int bc_bc_symmetry_2d(double* w, const char* loc, const int32_t* interf, const double* nx, const double* ny, int gh, int im, int jm) {
  return symmetry_host(w, nullptr, loc, interf, nx, ny, gh, im, jm);
}
int bc_bc_symmetry_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, const double* nx, const double* ny, int gh, int im,
                        int jm) {
  if (!wd) return fail(BC_ERR_ARG, "wd is null");
  return symmetry_host(w, wd, loc, interf, nx, ny, gh, im, jm);
}

static int noref_host(double* w, double* wd, const double* wbd, const char* loc, const int32_t* interf, const double* nx, const double* ny,
                      double gam, int gh, int im, int jm, int lm) {
  if (lm < 1) return fail(BC_ERR_ARG, "lm must be positive");
  return bc_host(w, wd, gh, im, jm, [&](const GridDesc& g, double* dw, double* dwd) -> int {
    double* dnx = dbuf<double>(S_NX, g.sn * 2);
    double* dny = dbuf<double>(S_NY, g.sn * 2);
    double* dwbd = dbuf<double>(S_AUX, (size_t)lm * 5);
    if (!dnx || !dny || !dwbd) return fail(BC_ERR_ALLOC, "device allocation failed");
    CK(cudaMemcpyAsync(dnx, nx, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(dny, ny, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(dwbd, wbd, sizeof(double) * lm * 5, cudaMemcpyHostToDevice, 0));
    return bcd_bc_no_reflexion(dw, dwd, wd ? 1 : 0, dwbd, loc, interf, dnx, dny, gam, gh, im, jm, lm, nullptr);
  });
}
int bc_bc_no_reflexion_2d(double* w, const double* wbd, const char* loc, const int32_t* interf, const double* nx, const double* ny,
                          double gam, int gh, int im, int jm, int lm) {
  return noref_host(w, nullptr, wbd, loc, interf, nx, ny, gam, gh, im, jm, lm);
}
int bc_bc_no_reflexion_2d_d(double* w, double* wd, const double* wbd, const char* loc, const int32_t* interf, const double* nx,
                            const double* ny, double gam, int gh, int im, int jm, int lm) {
  if (!wd) return fail(BC_ERR_ARG, "wd is null");
  return noref_host(w, wd, wbd, loc, interf, nx, ny, gam, gh, im, jm, lm);
}

static int inlet_host(double* w, double* wd, const char* loc, const int32_t* interf, const double* field, const double* nx,
                      const double* ny, double gam, int im, int jm, int lm, int gh) {
  if (lm < 1) return fail(BC_ERR_ARG, "lm must be positive");
  return bc_host(w, wd, gh, im, jm, [&](const GridDesc& g, double* dw, double* dwd) -> int {
    double* dnx = dbuf<double>(S_NX, g.sn * 2);
    double* dny = dbuf<double>(S_NY, g.sn * 2);
    double* dfield = dbuf<double>(S_AUX, (size_t)lm * gh * 5);
    if (!dnx || !dny || !dfield) return fail(BC_ERR_ALLOC, "device allocation failed");
    CK(cudaMemcpyAsync(dnx, nx, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(dny, ny, sizeof(double) * g.sn * 2, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(dfield, field, sizeof(double) * lm * gh * 5, cudaMemcpyHostToDevice, 0));
    return bcd_bc_supandsubinlet(dw, dwd, wd ? 1 : 0, loc, interf, dfield, dnx, dny, gam, im, jm, lm, gh, nullptr);
  });
}
int bc_bc_supandsubinlet_2d(double* w, const char* loc, const int32_t* interf, const double* field, const double* nx, const double* ny,
                            double gam, int im, int jm, int lm, int gh) {
  return inlet_host(w, nullptr, loc, interf, field, nx, ny, gam, im, jm, lm, gh);
}
int bc_bc_supandsubinlet_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, const double* field, const double* nx,
                              const double* ny, double gam, int im, int jm, int lm, int gh) {
  if (!wd) return fail(BC_ERR_ARG, "wd is null");
  return inlet_host(w, wd, loc, interf, field, nx, ny, gam, im, jm, lm, gh);
}

int bc_bc_extrapolate_o2_2d(double* w, const char* loc, const int32_t* interf, int im, int jm, int gh, int em) {
  if (em != 5) return fail(BC_ERR_UNSUPPORTED, "bc_extrapolate_o2_2d: em must be 5");
  return bc_host(w, nullptr, gh, im, jm,
                 [&](const GridDesc&, double* dw, double*) { return bcd_bc_extrapolate_o2(dw, nullptr, 0, loc, interf, im, jm, gh, nullptr); });
}
int bc_bc_extrapolate_o2_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, int im, int jm, int gh, int em) {
  if (em != 5) return fail(BC_ERR_UNSUPPORTED, "bc_extrapolate_o2_2d_d: em must be 5");
  if (!wd) return fail(BC_ERR_ARG, "wd is null");
  return bc_host(w, wd, gh, im, jm,
                 [&](const GridDesc&, double* dw, double* dwd) { return bcd_bc_extrapolate_o2(dw, dwd, 1, loc, interf, im, jm, gh, nullptr); });
}

static int general_host(double* w, double* wd, const char* loc, const int32_t* interf, const double* field, int gh, int im, int jm, int lm) {
  if (lm < 1) return fail(BC_ERR_ARG, "lm must be positive");
  return bc_host(w, wd, gh, im, jm, [&](const GridDesc&, double* dw, double* dwd) -> int {
    double* dfield = dbuf<double>(S_AUX, (size_t)lm * gh * 5);
    if (!dfield) return fail(BC_ERR_ALLOC, "device allocation failed");
    CK(cudaMemcpyAsync(dfield, field, sizeof(double) * lm * gh * 5, cudaMemcpyHostToDevice, 0));
    return bcd_bc_general(dw, dwd, wd ? 1 : 0, loc, interf, dfield, gh, im, jm, lm, nullptr);
  });
}
int bc_bc_general_2d(double* w, const char* loc, const int32_t* interf, const double* field, int gh, int im, int jm, int lm, int em,
                     int gh1) {
  if (em != 5 || gh1 != gh) return fail(BC_ERR_UNSUPPORTED, "bc_general_2d: field must be (lm, gh, 5)");
  return general_host(w, nullptr, loc, interf, field, gh, im, jm, lm);
}
int bc_bc_general_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, const double* field, int gh, int im, int jm, int lm) {
  if (!wd) return fail(BC_ERR_ARG, "wd is null");
  return general_host(w, wd, loc, interf, field, gh, im, jm, lm);
}

int bc_jn_match_2d(double* wr, const int32_t* prr, int gh1r, int gh2r, int gh3r, int gh4r, int imr, int jmr, const double* wd,
                   const int32_t* prd, int gh1d, int gh2d, int gh3d, int gh4d, int imd, int jmd, const int32_t* tr, int em) {
  if (int rc = check_device()) return rc;
  if (em < 1) return fail(BC_ERR_ARG, "em must be positive");
  const size_t nr = (size_t)(imr + gh1r + gh2r) * (jmr + gh3r + gh4r) * em;
  const size_t nd = (size_t)(imd + gh1d + gh2d) * (jmd + gh3d + gh4d) * em;
  double* dr = dbuf<double>(S_W, nr);
  if (!dr) return fail(BC_ERR_ALLOC, "device allocation failed");
  CK(cudaMemcpyAsync(dr, wr, sizeof(double) * nr, cudaMemcpyHostToDevice, 0));
  const double* dd = dr;
  if (wd != wr) {
    double* t = dbuf<double>(S_WD, nd);
    if (!t) return fail(BC_ERR_ALLOC, "device allocation failed");
    CK(cudaMemcpyAsync(t, wd, sizeof(double) * nd, cudaMemcpyHostToDevice, 0));
    dd = t;
  }
  if (int rc = bcd_jn_match(dr, prr, gh1r, gh2r, gh3r, gh4r, imr, jmr, dd, prd, gh1d, gh2d, gh3d, gh4d, imd, jmd, tr, em, nullptr)) return rc;
  CK(cudaMemcpyAsync(wr, dr, sizeof(double) * nr, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return BC_OK;
}
int bc_jn_match_geom_2d(double* wr, const int32_t* prr, int gh1r, int gh2r, int gh3r, int gh4r, int imr, int jmr, const double* wd,
                        const int32_t* prd, int gh1d, int gh2d, int gh3d, int gh4d, int imd, int jmd, const int32_t* tr) {
  return bc_jn_match_2d(wr, prr, gh1r, gh2r, gh3r, gh4r, imr, jmr, wd, prd, gh1d, gh2d, gh3d, gh4d, imd, jmd, tr, 1);
}

int bc_testvector(double* wd, int m, int l, int k, int gh, int im, int jm) {
  if (int rc = check_device()) return rc;
  if (int rc = check_dims(im, jm, gh)) return rc;
  const GridDesc g = make_grid(im, jm, gh);
  double* d = dbuf<double>(S_WD, g.sc * 5);
  if (!d) return fail(BC_ERR_ALLOC, "device allocation failed");
  if (int rc = bcd_testvector(d, 1, m, l, k, gh, im, jm, nullptr, nullptr)) return rc;
  CK(cudaMemcpy(wd, d, sizeof(double) * g.sc * 5, cudaMemcpyDeviceToHost));
  return BC_OK;
}
int bc_testvector_partial(double* wd, int m, int l, int k, int gh, int im, int jm, int istart, int iend, int jstart, int jend) {
  if (int rc = check_device()) return rc;
  if (int rc = check_dims(im, jm, gh)) return rc;
  const GridDesc g = make_grid(im, jm, gh);
  double* d = dbuf<double>(S_WD, g.sc * 5);
  if (!d) return fail(BC_ERR_ALLOC, "device allocation failed");
  const int32_t zone[4] = {istart, iend, jstart, jend};
  if (int rc = bcd_testvector(d, 1, m, l, k, gh, im, jm, zone, nullptr)) return rc;
  CK(cudaMemcpy(wd, d, sizeof(double) * g.sc * 5, cudaMemcpyDeviceToHost));
  return BC_OK;
}

int bc_computejacobianfromjv(double* jac, int32_t* ia, int32_t* ja, const double* resd, int m, int l, int k, int gh, int im, int jm,
                             int64_t nbentry) {
  return scatter_host(SCATTER_JV, jac, ia, ja, resd, m, l, k, gh, im, jm, nbentry, nullptr, nullptr);
}
int bc_computejacobianfromjv_relaxed(double* jac, int32_t* ia, int32_t* ja, const double* resd, int m, int l, int k, int gh, int im,
                                     int jm, int64_t nbentry, const double* coefdiag) {
  if (!coefdiag) return fail(BC_ERR_ARG, "coefdiag is null");
  return scatter_host(SCATTER_JV_RELAXED, jac, ia, ja, resd, m, l, k, gh, im, jm, nbentry, coefdiag, nullptr);
}
int bc_computejacobianfromjv_relaxed_withjn(double* jac, int32_t* ia, int32_t* ja, const double* resd, int m, int l, int k, int gh,
                                            int im, int jm, int64_t nbentry, const double* coefdiag) {
  if (!coefdiag) return fail(BC_ERR_ARG, "coefdiag is null");
  return scatter_host(SCATTER_JV_RELAXED_JN, jac, ia, ja, resd, m, l, k, gh, im, jm, nbentry, coefdiag, nullptr);
}
int bc_computejacobianfromjv_relaxed_withjnandcheck(double* jac, int32_t* ia, int32_t* ja, const double* resd, int m, int l, int k,
                                                    int gh, int im, int jm, int64_t nbentry, const double* coefdiag, double mini,
                                                    int n) {
  if (!coefdiag) return fail(BC_ERR_ARG, "coefdiag is null");
  if (int rc = check_device()) return rc;
  if (int rc = check_dims(im, jm, gh)) return rc;
  const int s = 2 * gh + 1;
  if (m < 0 || m > 4 || l < 0 || l >= s || k < 0 || k >= s || n < 0 || n > 1) return fail(BC_ERR_ARG, "colour / zone indices out of range");
  const GridDesc g = make_grid(im, jm, gh);
  const long long nn = 5LL * im * jm;
  const long long base = (long long)k * nn + (long long)l * nn * s + (long long)m * nn * s * s + (long long)n * nn * 5 * s * s;
  if (base + nn > nbentry) return fail(BC_ERR_ARG, "jac/ia/ja too short for this colour and zone");
  double* dres = dbuf<double>(S_RES, g.sc * 5);
  double* dseg = dbuf<double>(S_SEG, nn);
  int* dia = dbuf<int>(S_IA, nn);
  int* dja = dbuf<int>(S_JA, nn);
  double* dcoef = dbuf<double>(S_AUX, (size_t)im * jm);
  if (!dres || !dseg || !dia || !dja || !dcoef) return fail(BC_ERR_ALLOC, "device allocation failed");
  CK(cudaMemcpyAsync(dres, resd, sizeof(double) * g.sc * 5, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(dcoef, coefdiag, sizeof(double) * im * jm, cudaMemcpyHostToDevice, 0));
  // the routine is read-modify-write on its slot segment
  CK(cudaMemcpyAsync(dseg, jac + base, sizeof(double) * nn, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(dia, ia + base, sizeof(int) * nn, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(dja, ja + base, sizeof(int) * nn, cudaMemcpyHostToDevice, 0));
  cudaError_t e = launch_scatter_check(g, dseg, dia, dja, dres, m, l, k, dcoef, mini, n, 0);
  g_launches += 1;
  if (e != cudaSuccess) return cuda_fail(e, "computejacobianfromjv_relaxed_withjnandcheck");
  CK(cudaMemcpyAsync(jac + base, dseg, sizeof(double) * nn, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(ia + base, dia, sizeof(int) * nn, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(ja + base, dja, sizeof(int) * nn, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return BC_OK;
}
int bc_computejacobianfromjv_withjn(double* jac, int32_t* ia, int32_t* ja, const double* resd, int m, int l, int k, int gh, int im,
                                    int jm, int64_t nbentry) {
  return scatter_host(SCATTER_JV_JN, jac, ia, ja, resd, m, l, k, gh, im, jm, nbentry, nullptr, nullptr);
}
int bc_computejacobianfromdz(double* jac, int32_t* ia, int32_t* ja, const double* dz, int m, int l, int k, int gh, int im, int jm,
                             int64_t nbentry) {
  return scatter_host(SCATTER_DZ, jac, ia, ja, dz, m, l, k, gh, im, jm, nbentry, nullptr, nullptr);
}

int bc_compute_norml2(double* norm, double* nmoy, const double* rhs, int im, int jm, int gh) {
  // norm.F90:2-32: norm = sqrt(sum r^2), nmoy = norm / (im*jm)   [see the file for the exact second output]
  double s2[5], s10[5];
  if (int rc = norms_host(s2, s10, rhs, im, jm, gh)) return rc;
  for (int e = 0; e < 5; ++e) {
    norm[e] = ::sqrt(s2[e]);
    nmoy[e] = norm[e] / ((double)im * jm);
  }
  return BC_OK;
}
int bc_compute_norml2inf(double* norm, double* ninf, const double* rhs, int im, int jm, int gh) {
  double s2[5], s10[5];
  if (int rc = norms_host(s2, s10, rhs, im, jm, gh)) return rc;
  for (int e = 0; e < 5; ++e) {
    norm[e] = ::sqrt(s2[e]);
    ninf[e] = ::pow(s10[e], 0.1);
  }
  return BC_OK;
}

}  // extern "C"
