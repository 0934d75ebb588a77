// The step after the assembly (row f4 of SURVEY.md section 8): the Newton correction  A dw = res  on the device, with A the CSR row
// block the assembly path produces (csr.cu).  The reference solves it with PETSc / MUMPS LU on the host (misc/PETSc_func.py:137-152
// kspLUPetsc, :247-263 iterNewton; BROADCAST_npz.py:1043-1047 leaves `lasolver == 'gmres'` as "not yet implemented"); here the matrix
// never leaves HBM: restarted GMRES with right preconditioning by the inverse 5 x 5 diagonal blocks (one per cell: the pseudo-time
// term coefdiag = vol / dt of the relaxed Jacobian sits on exactly those blocks).  The adjoint systems of the sensitivity drivers
// (cylinder.py:1090-1177) use the same solver on the transposed CSR (bcd_csr_transpose_*).
//
//   k_spmv_csr          one warp per row (<= 145 entries), lanes stride over the row, shuffle reduction
//   k_bj_setup          one thread per cell: picks the diagonal block out of its five rows, inverts it (partial pivoting)
//   k_bj_apply          z = D^-1 r, one thread per cell, planes of D^-1 read coalesced
//   k_multi_dot / k_multi_dot_reduce   h_i = <V_i, w> for all basis vectors in ONE pass over w (fixed grid: deterministic sums)
//   k_multi_axpy        w -= sum_i h_i V_i with h read from device memory (no host round trip between the two)
// The host loop (bcd_gmres) only handles the (m + 1) x m Hessenberg matrix and its Givens rotations.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../include/broadcast_b200.h"
#include "kernels.cuh"

namespace bcast {
void count_launches(int n);

namespace {

constexpr int DOT_BLOCKS = 592;    // 4 x 148 SMs
constexpr int DOT_THREADS = 256;
constexpr int MAXK = 64;           // basis vectors handled by one multi-dot / multi-axpy launch

__global__ void __launch_bounds__(256) k_spmv_csr(const long long* __restrict__ indptr, const int* __restrict__ indices,
                                                  const double* __restrict__ data, const double* __restrict__ x, long long nrows,
                                                  double* __restrict__ y) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const long long p0 = indptr[r], p1 = indptr[r + 1];
  double s = 0.0;
  for (long long p = p0 + lane; p < p1; p += 32) s += data[p] * __ldg(x + indices[p]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) y[r] = s;
}

// dinv[q][cell], q = e * 5 + m: inverse of the diagonal block of cell `cell` (rows / columns 5 cell .. 5 cell + 4, local rows, global
// columns col0 + 5 cell + m).  A singular block (pivot 0) is replaced by the identity and counted in *bad.
__global__ void __launch_bounds__(128) k_bj_setup(const long long* __restrict__ indptr, const int* __restrict__ indices,
                                                  const double* __restrict__ data, long long ncell, long long col0, double* __restrict__ dinv,
                                                  int* __restrict__ bad) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  double a[5][10];
#pragma unroll
  for (int e = 0; e < 5; ++e) {
#pragma unroll
    for (int m = 0; m < 10; ++m) a[e][m] = (m - 5 == e) ? 1.0 : 0.0;
    const long long p0 = indptr[5 * c + e], p1 = indptr[5 * c + e + 1];
    const long long cb = col0 + 5 * c;
    for (long long p = p0; p < p1; ++p) {
      const long long col = indices[p];
      if (col >= cb && col < cb + 5) {
        const int m = (int)(col - cb);
#pragma unroll
        for (int q = 0; q < 5; ++q)
          if (q == m) a[e][q] = data[p];
      }
    }
  }
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    int piv = k;
    double best = ::fabs(a[k][k]);
#pragma unroll
    for (int r = 0; r < 5; ++r)
      if (r > k && ::fabs(a[r][k]) > best) { best = ::fabs(a[r][k]); piv = r; }
    if (best == 0.0) { ok = false; break; }
#pragma unroll
    for (int r = 0; r < 5; ++r)
      if (r == piv && piv != k) {
#pragma unroll
        for (int m = 0; m < 10; ++m) { const double t = a[k][m]; a[k][m] = a[r][m]; a[r][m] = t; }
      }
    const double inv = 1.0 / a[k][k];
#pragma unroll
    for (int m = 0; m < 10; ++m) a[k][m] *= inv;
#pragma unroll
    for (int r = 0; r < 5; ++r)
      if (r != k) {
        const double f = a[r][k];
#pragma unroll
        for (int m = 0; m < 10; ++m) a[r][m] -= f * a[k][m];
      }
  }
  if (!ok) atomicAdd(bad, 1);
#pragma unroll
  for (int e = 0; e < 5; ++e)
#pragma unroll
    for (int m = 0; m < 5; ++m) dinv[(long long)(e * 5 + m) * ncell + c] = ok ? a[e][5 + m] : (e == m ? 1.0 : 0.0);
}

__global__ void __launch_bounds__(128) k_bj_apply(const double* __restrict__ dinv, const double* __restrict__ r, long long ncell,
                                                  double* __restrict__ z) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  double v[5];
#pragma unroll
  for (int m = 0; m < 5; ++m) v[m] = r[5 * c + m];
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < 5; ++m) s += dinv[(long long)(e * 5 + m) * ncell + c] * v[m];
    z[5 * c + e] = s;
  }
}

// partial[b][i] = sum over the elements of block b of V_i * w, i < k (V_i = V + i * ldv)
__global__ void __launch_bounds__(DOT_THREADS) k_multi_dot(const double* __restrict__ V, long long ldv, int k, const double* __restrict__ w,
                                                           long long n, double* __restrict__ partial) {
  __shared__ double red[DOT_THREADS / 32][MAXK];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i0 = 0; i0 < k; i0 += 8) {
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int kk = min(8, k - i0);
    for (long long t = blockIdx.x * (long long)DOT_THREADS + threadIdx.x; t < n; t += (long long)gridDim.x * DOT_THREADS) {
      const double wv = w[t];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (u < kk) s[u] += V[(long long)(i0 + u) * ldv + t] * wv;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      double v = s[u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0 && u < kk) red[wid][i0 + u] = v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += DOT_THREADS) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < DOT_THREADS / 32; ++q) v += red[q][i];
    partial[(long long)blockIdx.x * MAXK + i] = v;
  }
}
// out[i] (+)= sum_b partial[b][i]; one warp per i, fixed order
__global__ void k_multi_dot_reduce(const double* __restrict__ partial, int nb, int k, double* __restrict__ out, int accumulate) {
  const int i = blockIdx.x;
  double v = 0.0;
  for (int b = threadIdx.x; b < nb; b += 32) v += partial[(long long)b * MAXK + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (threadIdx.x == 0 && i < k) out[i] = accumulate ? out[i] + v : v;
}
// w -= sum_i h[i] V_i
__global__ void __launch_bounds__(256) k_multi_axpy(const double* __restrict__ V, long long ldv, int k, const double* __restrict__ h,
                                                    long long n, double* __restrict__ w) {
  __shared__ double hs[MAXK];
  if (threadIdx.x < k) hs[threadIdx.x] = h[threadIdx.x];
  __syncthreads();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    double v = w[t];
    for (int i = 0; i < k; ++i) v -= hs[i] * V[(long long)i * ldv + t];
    w[t] = v;
  }
}
// y = a * x + b * y (b == 0: y = a * x); a taken as *pa (device) scaled by `as`, or `as` alone when pa is null; inv: use 1 / *pa
__global__ void __launch_bounds__(256) k_axpby(const double* __restrict__ x, const double* __restrict__ pa, double as, int inv, double b,
                                               long long n, double* __restrict__ y) {
  double a = as;
  if (pa) a *= inv ? 1.0 / *pa : *pa;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    y[t] = b == 0.0 ? a * x[t] : a * x[t] + b * y[t];
}
__global__ void k_sqrt1(double* v) { *v = ::sqrt(*v); }
// x += sum_i yv[i] Z_i  (yv on the device)
__global__ void __launch_bounds__(256) k_update(const double* __restrict__ Z, long long ldz, int k, const double* __restrict__ yv, long long n,
                                                double* __restrict__ x) {
  __shared__ double ys[MAXK];
  if (threadIdx.x < k) ys[threadIdx.x] = yv[threadIdx.x];
  __syncthreads();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    double v = x[t];
    for (int i = 0; i < k; ++i) v += ys[i] * Z[(long long)i * ldz + t];
    x[t] = v;
  }
}

struct Csr {
  const long long* indptr;
  const int* indices;
  const double* data;
  long long n;
};
void spmv(const Csr& A, const double* x, double* y, cudaStream_t st) {
  k_spmv_csr<<<(unsigned)((A.n * 32 + 255) / 256), 256, 0, st>>>(A.indptr, A.indices, A.data, x, A.n, y);
}
void multi_dot(const double* V, long long ldv, int k, const double* w, long long n, double* partial, double* out, int accumulate,
               cudaStream_t st) {
  k_multi_dot<<<DOT_BLOCKS, DOT_THREADS, 0, st>>>(V, ldv, k, w, n, partial);
  k_multi_dot_reduce<<<k, 32, 0, st>>>(partial, DOT_BLOCKS, k, out, accumulate);
}

}  // namespace
}  // namespace bcast

using namespace bcast;

extern "C" int bcd_csr_spmv(double* y, const long long* indptr, const int32_t* indices, const double* data, const double* x, long long nrows,
                            void* stream) {
  if (nrows < 0 || !y || !indptr || !x) return BC_ERR_ARG;
  if (nrows == 0) return BC_OK;
  spmv(Csr{indptr, indices, data, nrows}, x, y, (cudaStream_t)stream);
  count_launches(1);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? BC_OK : (int)e;
}

extern "C" int bcd_block_jacobi_setup(double* dinv, int32_t* nbad, const long long* indptr, const int32_t* indices, const double* data,
                                      long long ncell, long long col0, void* stream) {
  if (ncell < 1 || !dinv || !nbad) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(nbad, 0, sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  k_bj_setup<<<(unsigned)((ncell + 127) / 128), 128, 0, st>>>(indptr, indices, data, ncell, col0, dinv, nbad);
  count_launches(1);
  e = cudaGetLastError();
  return e == cudaSuccess ? BC_OK : (int)e;
}

extern "C" int bcd_block_jacobi_apply(double* z, const double* dinv, const double* r, long long ncell, void* stream) {
  if (ncell < 1 || !z || !dinv || !r) return BC_ERR_ARG;
  k_bj_apply<<<(unsigned)((ncell + 127) / 128), 128, 0, (cudaStream_t)stream>>>(dinv, r, ncell, z);
  count_launches(1);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? BC_OK : (int)e;
}

extern "C" long long bcd_gmres_work_doubles(long long n, int restart) {
  if (n < 1 || restart < 1 || restart >= MAXK) return -1;
  // V (restart + 1 vectors), w, z, partial sums, h / scalars
  return (long long)(restart + 3) * n + (long long)DOT_BLOCKS * MAXK + 4 * MAXK;
}

// Restarted GMRES for the square system A x = b (n = 5 ncell rows, whole matrix on this device), preconditioned by the inverse
// diagonal blocks `dinv` (bcd_block_jacobi_setup; null: none) on the LEFT (side = 1: M^-1 A x = M^-1 b, the default of the Python
// layer -- the rows of the finite-volume Jacobian scale with the cell sizes, which vary by orders of magnitude on the stretched
// boundary-layer meshes, and the left-preconditioned operator is free of that scaling) or on the RIGHT (side = 0: A M^-1 y = b).
// x: start vector in, solution out.  Stops when the (preconditioned, for side = 1) residual is below rtol times the (preconditioned)
// right-hand side -- estimated by the Givens recurrence, re-evaluated at every restart -- or after maxit matrix-vector products.
// info[0] = products used, info[1] = 1 if converged; relres[0] = final TRUE relative residual ||b - A x|| / ||b||, relres[1] = the
// relative residual the iteration controls.
extern "C" int bcd_gmres(double* x, const double* b, const long long* indptr, const int32_t* indices, const double* data, const double* dinv,
                         long long n, int restart, int maxit, double rtol, int side, double* work, long long work_len, int32_t* info,
                         double* relres, void* stream) {
  if (n < 5 || n % 5 || restart < 1 || restart >= MAXK || maxit < 1 || !x || !b || !work || (side != 0 && side != 1)) return BC_ERR_ARG;
  if (work_len < bcd_gmres_work_doubles(n, restart)) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const Csr A{indptr, indices, data, n};
  const int m = restart;
  const bool left = side == 1 && dinv != nullptr;
  double* V = work;                              // (m + 1) x n
  double* w = V + (long long)(m + 1) * n;
  double* z = w + n;
  double* partial = z + n;
  double* hd = partial + (long long)DOT_BLOCKS * MAXK;   // [MAXK] first pass, [MAXK] second pass, [MAXK] scalars / y, [MAXK] spare
  double* h2 = hd + MAXK;
  double* sc = h2 + MAXK;
  const int GB = 592 * 2;
  auto precond = [&](const double* in, double* out) {
    if (dinv) k_bj_apply<<<(unsigned)((n / 5 + 127) / 128), 128, 0, st>>>(dinv, in, n / 5, out);
    else cudaMemcpyAsync(out, in, sizeof(double) * n, cudaMemcpyDeviceToDevice, st);
  };
  auto fetch = [&](double* host, const double* dev, int cnt) -> cudaError_t {
    cudaError_t e = cudaMemcpyAsync(host, dev, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e;
  };
  auto norm_of = [&](const double* v, double* out) -> cudaError_t {   // *out = ||v|| (host), sc[0] = ||v|| (device)
    multi_dot(v, n, 1, v, n, partial, sc, 0, st);
    k_sqrt1<<<1, 1, 0, st>>>(sc);
    return fetch(out, sc, 1);
  };
  // norms of the right-hand side: true, and the one the iteration controls
  double bnorm_true = 0.0, bnorm = 0.0;
  cudaError_t e = norm_of(b, &bnorm_true);
  if (e != cudaSuccess) return (int)e;
  bnorm = bnorm_true;
  if (left) {
    precond(b, z);
    e = norm_of(z, &bnorm);
    if (e != cudaSuccess) return (int)e;
  }
  int its = 0, conv = 0;
  double rel = 0.0;
  if (bnorm_true == 0.0 || bnorm == 0.0) {
    cudaMemsetAsync(x, 0, sizeof(double) * n, st);
    if (info) { info[0] = 0; info[1] = 1; }
    if (relres) { relres[0] = 0.0; relres[1] = 0.0; }
    return BC_OK;
  }
  std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), hcol(2 * MAXK + 1), y(m);
  while (true) {
    // r = [M^-1] (b - A x) -> V_0 = r / ||r||
    spmv(A, x, w, st);
    k_axpby<<<GB, 256, 0, st>>>(b, nullptr, 1.0, 0, -1.0, n, w);       // w = b - w
    if (left) {
      precond(w, z);
      cudaMemcpyAsync(w, z, sizeof(double) * n, cudaMemcpyDeviceToDevice, st);
    }
    double beta = 0.0;
    e = norm_of(w, &beta);
    if (e != cudaSuccess) return (int)e;
    rel = beta / bnorm;
    if (rel <= rtol) { conv = 1; break; }
    if (its >= maxit) break;
    k_axpby<<<GB, 256, 0, st>>>(w, sc, 1.0, 1, 0.0, n, V);              // V_0 = w / beta
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int j = 0;
    for (; j < m && its < maxit; ++j) {
      if (left) {
        spmv(A, V + (long long)j * n, z, st);
        precond(z, w);
      } else {
        precond(V + (long long)j * n, z);
        spmv(A, z, w, st);
      }
      ++its;
      // classical Gram-Schmidt, twice (as stable as the modified variant, two passes over the basis instead of j + 1)
      multi_dot(V, n, j + 1, w, n, partial, hd, 0, st);
      k_multi_axpy<<<GB, 256, 0, st>>>(V, n, j + 1, hd, n, w);
      multi_dot(V, n, j + 1, w, n, partial, h2, 0, st);
      k_multi_axpy<<<GB, 256, 0, st>>>(V, n, j + 1, h2, n, w);
      multi_dot(w, n, 1, w, n, partial, sc, 0, st);
      k_sqrt1<<<1, 1, 0, st>>>(sc);
      k_axpby<<<GB, 256, 0, st>>>(w, sc, 1.0, 1, 0.0, n, V + (long long)(j + 1) * n);   // V_{j+1} = w / ||w||
      e = fetch(hcol.data(), hd, 2 * MAXK + 1);                         // hd, h2, sc[0] are contiguous
      if (e != cudaSuccess) return (int)e;
      for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = hcol[i] + hcol[MAXK + i];
      const double hn = hcol[2 * MAXK];
      H[(size_t)(j + 1) * m + j] = hn;
      for (int i = 0; i < j; ++i) {
        const double t = cs[i] * H[(size_t)i * m + j] + sn[i] * H[(size_t)(i + 1) * m + j];
        H[(size_t)(i + 1) * m + j] = -sn[i] * H[(size_t)i * m + j] + cs[i] * H[(size_t)(i + 1) * m + j];
        H[(size_t)i * m + j] = t;
      }
      const double a0 = H[(size_t)j * m + j], a1 = H[(size_t)(j + 1) * m + j];
      const double d = std::hypot(a0, a1);
      cs[j] = d == 0.0 ? 1.0 : a0 / d;
      sn[j] = d == 0.0 ? 0.0 : a1 / d;
      H[(size_t)j * m + j] = d;
      H[(size_t)(j + 1) * m + j] = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      rel = std::fabs(g[j + 1]) / bnorm;
      if (rel <= rtol || hn == 0.0) { ++j; break; }
    }
    // y = H^-1 g (upper triangular, j columns); x += V y (left) or M^-1 (V y) (right)
    for (int i = j - 1; i >= 0; --i) {
      double s = g[i];
      for (int q = i + 1; q < j; ++q) s -= H[(size_t)i * m + q] * y[q];
      y[i] = s / H[(size_t)i * m + i];
    }
    e = cudaMemcpyAsync(sc + 1, y.data(), sizeof(double) * j, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return (int)e;
    if (left) {
      k_update<<<GB, 256, 0, st>>>(V, n, j, sc + 1, n, x);              // x += V y
    } else {
      cudaMemsetAsync(w, 0, sizeof(double) * n, st);
      k_update<<<GB, 256, 0, st>>>(V, n, j, sc + 1, n, w);              // w = V y
      precond(w, z);
      k_axpby<<<GB, 256, 0, st>>>(z, nullptr, 1.0, 0, 1.0, n, x);       // x += z
    }
    e = cudaStreamSynchronize(st);                                      // y (host vector) is reused by the next cycle
    if (e != cudaSuccess) return (int)e;
    count_launches(12 * j + 8);
  }
  // the true relative residual of what is returned
  spmv(A, x, w, st);
  k_axpby<<<GB, 256, 0, st>>>(b, nullptr, 1.0, 0, -1.0, n, w);
  double rt = 0.0;
  e = norm_of(w, &rt);
  if (e != cudaSuccess) return (int)e;
  if (info) { info[0] = its; info[1] = conv; }
  if (relres) { relres[0] = rt / bnorm_true; relres[1] = rel; }
  e = cudaGetLastError();
  return e == cudaSuccess ? BC_OK : (int)e;
}
