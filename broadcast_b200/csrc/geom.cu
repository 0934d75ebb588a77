// Mesh metrics on the device: ghost-node extension, cell centres, volumes, area-scaled face normals,
// ghost extension of the metrics and the face volume factors.
// Reference (restated): srcfv/geom/computegeom.F90:3-104 with centers.F, volumes.F:1-14,
// normals_idir.F, normals_jdir.F.  The reference's whole-array statements are executed in the same
// order (each one is a data-parallel line update), so corner values come out identical.
#include "../../include/broadcast_b200.h"
#include "kernels.cuh"

namespace bcast {

// a(dst, t) = 2 a(s1, t) - a(s2, t) for t in [lo, hi]   (dim == 0: first index fixed, line along j)
// a(t, dst) = 2 a(t, s1) - a(t, s2)                       (dim == 1: second index fixed, line along i)
// indices are 0-based storage indices; `planes` consecutive planes of stride `ps` are updated.
__global__ void k_line_extrap(double* a, int ld, long long ps, int planes, int dim, int dst, int s1, int s2, int lo, int hi) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x + lo;
  const int p = blockIdx.y;
  if (t > hi || p >= planes) return;
  double* q = a + p * ps;
  if (dim == 0)
    q[dst + (long long)t * ld] = 2.0 * q[s1 + (long long)t * ld] - q[s2 + (long long)t * ld];
  else
    q[t + (long long)dst * ld] = 2.0 * q[t + (long long)s1 * ld] - q[t + (long long)s2 * ld];
}

__global__ void k_metrics(GridDesc g, const double* __restrict__ x0, const double* __restrict__ y0, double* __restrict__ nx,
                          double* __restrict__ ny, double* __restrict__ xc, double* __restrict__ yc, double* __restrict__ vol) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > g.im + 1 || j > g.jm + 1) return;
  const long long n00 = g.nidx(i, j), n10 = g.nidx(i + 1, j), n01 = g.nidx(i, j + 1), n11 = g.nidx(i + 1, j + 1);
  const double xa = x0[n00], ya = y0[n00], xb = x0[n10], yb = y0[n10], xc1 = x0[n01], yc1 = y0[n01], xd = x0[n11], yd = y0[n11];
  const long long c = g.cidx(i, j);
  xc[c] = 0.25 * (xa + xb + xc1 + xd);
  yc[c] = 0.25 * (ya + yb + yc1 + yd);
  const double abx = xb - xa, aby = yb - ya, acx = xc1 - xa, acy = yc1 - ya;
  const double q1 = 0.5 * ::fabs(abx * acy - acx * aby);
  const double dcx = xc1 - xd, dcy = yc1 - yd, dbx = xb - xd, dby = yb - yd;
  const double q2 = 0.5 * ::fabs(dcx * dby - dbx * dcy);
  vol[c] = q1 + q2;
  nx[n00] = yc1 - ya;
  ny[n00] = xa - xc1;
  nx[g.sn + n00] = ya - yb;
  ny[g.sn + n00] = xb - xa;
}

__global__ void k_volf(GridDesc g, const double* __restrict__ vol, double* __restrict__ volf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > g.im + 1 || j > g.jm + 1) return;
  const long long c = g.cidx(i, j);
  volf[c] = 2.0 / (vol[c] + vol[g.cidx(i - 1, j)]);
  volf[g.sc + c] = 2.0 / (vol[c] + vol[g.cidx(i, j - 1)]);
}

static int computegeom_device(const GridDesc& g, double* x0, double* y0, double* nx, double* ny, double* xc, double* yc, double* vol,
                              double* volf, cudaStream_t st) {
  const int im = g.im, jm = g.jm, gh = g.gh;
  // storage index of Fortran index f (lower bound 1-gh)
  auto S = [&](int f) { return f - 1 + gh; };
  auto ext = [&](double* a, int ld, long long ps, int planes, int dim, int dst, int s1, int s2, int lo, int hi) {
    const int n = hi - lo + 1;
    k_line_extrap<<<dim3((n + 127) / 128, planes), 128, 0, st>>>(a, ld, ps, planes, dim, S(dst), S(s1), S(s2), S(lo), S(hi));
  };
  const int nlo = 1 - gh, nihi = im + gh + 1, njhi = jm + gh + 1;  // node array bounds
  const int clo = 1 - gh, cihi = im + gh, cjhi = jm + gh;          // cell array bounds
  for (int gg = 1; gg <= gh; ++gg)
    for (int dummy = 0; dummy < 2; ++dummy) {
      ext(x0, g.ldn, g.sn, 1, 0, 1 - gg, 2 - gg, 3 - gg, nlo, njhi);
      ext(x0, g.ldn, g.sn, 1, 0, im + 1 + gg, im + gg, im - 1 + gg, nlo, njhi);
      ext(x0, g.ldn, g.sn, 1, 1, 1 - gg, 2 - gg, 3 - gg, nlo, nihi);
      ext(x0, g.ldn, g.sn, 1, 1, jm + 1 + gg, jm + gg, jm - 1 + gg, nlo, nihi);
      ext(y0, g.ldn, g.sn, 1, 1, 1 - gg, 2 - gg, 3 - gg, nlo, nihi);
      ext(y0, g.ldn, g.sn, 1, 1, jm + 1 + gg, jm + gg, jm - 1 + gg, nlo, nihi);
      ext(y0, g.ldn, g.sn, 1, 0, 1 - gg, 2 - gg, 3 - gg, nlo, njhi);
      ext(y0, g.ldn, g.sn, 1, 0, im + 1 + gg, im + gg, im - 1 + gg, nlo, njhi);
    }
  dim3 blk(32, 4), grd((im + 1 + 31) / 32, (jm + 1 + 3) / 4);
  k_metrics<<<grd, blk, 0, st>>>(g, x0, y0, nx, ny, xc, yc, vol);
  for (int gg = 1; gg <= gh; ++gg)
    for (int dummy = 0; dummy < 2; ++dummy) {
      for (double* a : {xc, yc}) {
        ext(a, g.ldc, g.sc, 1, 0, 1 - gg, 2 - gg, 3 - gg, clo, cjhi);
        ext(a, g.ldc, g.sc, 1, 0, im + gg, im - 1 + gg, im - 2 + gg, clo, cjhi);
        ext(a, g.ldc, g.sc, 1, 1, 1 - gg, 2 - gg, 3 - gg, clo, cihi);
        ext(a, g.ldc, g.sc, 1, 1, jm + gg, jm - 1 + gg, jm - 2 + gg, clo, cihi);
      }
      ext(vol, g.ldc, g.sc, 1, 0, 1 - gg, 2 - gg, 3 - gg, 1, jm);
      ext(vol, g.ldc, g.sc, 1, 0, im + gg, im - 1 + gg, im - 2 + gg, 1, jm);
      ext(vol, g.ldc, g.sc, 1, 1, 1 - gg, 2 - gg, 3 - gg, clo, cihi);
      ext(vol, g.ldc, g.sc, 1, 1, jm + gg, jm - 1 + gg, jm - 2 + gg, clo, cihi);
      ext(nx, g.ldn, g.sn, 2, 0, 1 - gg, 2 - gg, 3 - gg, nlo, njhi);
      ext(ny, g.ldn, g.sn, 2, 0, 1 - gg, 2 - gg, 3 - gg, nlo, njhi);
      ext(nx, g.ldn, g.sn, 2, 0, im + 1 + gg, im + gg, im - 1 + gg, nlo, njhi);
      ext(ny, g.ldn, g.sn, 2, 0, im + 1 + gg, im + gg, im - 1 + gg, nlo, njhi);
      ext(nx, g.ldn, g.sn, 2, 1, 1 - gg, 2 - gg, 3 - gg, nlo, nihi);
      ext(ny, g.ldn, g.sn, 2, 1, 1 - gg, 2 - gg, 3 - gg, nlo, nihi);
      ext(nx, g.ldn, g.sn, 2, 1, 1 + jm + gg, jm + gg, gg + jm - 1, nlo, nihi);
      ext(ny, g.ldn, g.sn, 2, 1, 1 + jm + gg, jm + gg, gg + jm - 1, nlo, nihi);
    }
  k_volf<<<grd, blk, 0, st>>>(g, vol, volf);
  return (int)cudaGetLastError();
}

void count_launches(int n);

}  // namespace bcast

using namespace bcast;

extern "C" int bc_computegeom_2d(double* x0, double* y0, double* nx, double* ny, double* xc, double* yc, double* vol, double* volf,
                                 int im, int jm, int gh) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return BC_ERR_NODEV;
  if (im < 1 || jm < 1 || gh < 1) return BC_ERR_ARG;
  const GridDesc g = make_grid(im, jm, gh);
  const size_t nn = (size_t)g.sn, nc = (size_t)g.sc;
  // one arena: x0 y0 nx(2) ny(2) | xc yc vol volf(2)
  double* base = scratch_doubles(40, 6 * nn + 5 * nc);
  if (!base) return BC_ERR_ALLOC;
  double* dx0 = base;
  double* dy0 = dx0 + nn;
  double* dnx = dy0 + nn;
  double* dny = dnx + 2 * nn;
  double* dxc = dny + 2 * nn;
  double* dyc = dxc + nc;
  double* dvol = dyc + nc;
  double* dvolf = dvol + nc;
  struct Cp {
    double* d;
    double* h;
    size_t n;
  } cps[] = {{dx0, x0, nn}, {dy0, y0, nn}, {dnx, nx, 2 * nn}, {dny, ny, 2 * nn}, {dxc, xc, nc}, {dyc, yc, nc}, {dvol, vol, nc}, {dvolf, volf, 2 * nc}};
  for (auto& c : cps)
    if (cudaMemcpyAsync(c.d, c.h, c.n * sizeof(double), cudaMemcpyHostToDevice, 0) != cudaSuccess) return (int)cudaGetLastError();
  int rc = computegeom_device(g, dx0, dy0, dnx, dny, dxc, dyc, dvol, dvolf, 0);
  if (rc) return rc;
  count_launches(2 + gh * 2 * 32);
  for (auto& c : cps)
    if (cudaMemcpyAsync(c.h, c.d, c.n * sizeof(double), cudaMemcpyDeviceToHost, 0) != cudaSuccess) return (int)cudaGetLastError();
  if (cudaStreamSynchronize(0) != cudaSuccess) return (int)cudaGetLastError();
  return BC_OK;
}
