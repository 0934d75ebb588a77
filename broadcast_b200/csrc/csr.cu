// Hybrid block Jacobian -> CSR on the device: the last step of the assembly path, what the reference does on the host with
// remove_zero_jac (BROADCAST_npz.py:129-135: keep |v| > 2e-16), scipy's csr_matrix((Jac,(IA,JA))) (misc/PETSc_func.py:85) and the
// interpreted "divide by the cell volume" loop (BROADCAST_npz.py:1206-1209).
//
// Input: the 29 fixed 5x5 blocks per regular row cell (values[slot][e*5+m][cell], bcd_jacobian_interior) and the compact COO
// lists of the boundary strips (bcd_jacobian_strips).  Output: indptr (int64), indices (int32, ascending inside a row), data of
// the rows of this block (local row = e + 5 (j-1) + 5 jm (i-1), the reference's numbering; columns are global).
//   1. k_csr_count_interior / k_csr_count_strip   kept entries per row (the sparsity pattern is value dependent, as in the reference)
//   2. k_scan_*                                   exclusive scan of the counts: warp shuffles inside a warp, shared memory across the
//                                                 warps of a block, a second level over the block sums
//   3. k_csr_fill_interior                        32 cells per block: the 145 candidate planes of an equation row staged in shared
//                                                 memory by coalesced loads, then one WARP per cell row: candidates in column
//                                                 order, kept ones compacted with ballot + popc and written as contiguous runs
//                                                 (slot order is column order: no sort)
//      k_csr_fill_strip + k_csr_sort_strip_rows   strip entries are appended with one atomic per entry, then every strip row
//                                                 (<= 245 entries) is put in column order by a warp-wide rank sort in shared memory
#include <cstdint>
#include "../../include/broadcast_b200.h"
#include "facejac.cuh"
#include "kernels.cuh"

namespace bcast {
void count_launches(int n);

namespace {

__constant__ int kSlotDi[JAC_NSLOT] = {
#define X(a, b) a,
    BCAST_JAC_OFFSETS(X)
#undef X
};
__constant__ int kSlotDj[JAC_NSLOT] = {
#define X(a, b) b,
    BCAST_JAC_OFFSETS(X)
#undef X
};

constexpr int NCAND = JAC_NSLOT * 5;  // candidates of one row: 29 column cells x 5 variables

// ---- 1. counts ---------------------------------------------------------------------------------------------------------
// thread per cell of the region (i fastest: every plane is read coalesced), five row counters in registers
__global__ void __launch_bounds__(128) k_csr_count_interior(GridDesc g, Rect rc, const double* __restrict__ V, double thresh,
                                                            int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + rc.i0;
  const int j = blockIdx.y + rc.j0;
  if (i > rc.i1) return;
  const long long ncell = (long long)g.im * g.jm;
  const long long cell = (long long)(i - 1) + (long long)(j - 1) * g.im;
  int cnt[5] = {0, 0, 0, 0, 0};
  for (int s = 0; s < JAC_NSLOT; ++s) {
#pragma unroll
    for (int q = 0; q < 25; ++q) cnt[q / 5] += ::fabs(__ldg(V + ((long long)s * 25 + q) * ncell + cell)) > thresh ? 1 : 0;
  }
  const long long row = 5LL * (j - 1) + 5LL * g.jm * (i - 1);
#pragma unroll
  for (int e = 0; e < 5; ++e) counts[row + e] = cnt[e];
}

// (a strip list may hold rows of other row blocks too -- the banded assembly hands the lists of the whole block to every band:
// entries outside the nrows rows of this block are skipped)
__global__ void k_csr_count_strip(const double* __restrict__ jac, const int* __restrict__ ia, long long n, double thresh, long long row0,
                                  long long nrows, int* __restrict__ counts) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  if (::fabs(jac[t]) > thresh) {
    const long long r = ia[t] - row0;
    if (r >= 0 && r < nrows) atomicAdd(&counts[r], 1);
  }
}

// ---- 2. exclusive scan int32 -> int64 ---------------------------------------------------------------------------------
constexpr int SCAN_T = 256, SCAN_ITEMS = 8, SCAN_CHUNK = SCAN_T * SCAN_ITEMS;

__device__ __forceinline__ long long warp_incl_scan(long long v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long u = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += u;
  }
  return v;
}
// inclusive scan over the threads of a block (blockDim.x <= 1024); *total = block sum
__device__ __forceinline__ long long block_incl_scan(long long v, long long* total) {
  __shared__ long long wsum[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_incl_scan(v);
  if (lane == 31) wsum[wid] = v;
  __syncthreads();
  if (wid == 0) {
    long long s = lane < nw ? wsum[lane] : 0;
    s = warp_incl_scan(s);
    wsum[lane] = s;
  }
  __syncthreads();
  if (wid > 0) v += wsum[wid - 1];
  *total = wsum[nw - 1];
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_reduce(const int* __restrict__ counts, long long n, long long* __restrict__ bsum) {
  const long long base = (long long)blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) s += counts[base + k];
  long long tot;
  block_incl_scan(s, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
// one block: exclusive scan of the block sums in place (tiles of blockDim.x with a running carry)
__global__ void __launch_bounds__(1024) k_scan_blocks(long long* __restrict__ bsum, int nb) {
  long long carry = 0;
  for (int t0 = 0; t0 < nb; t0 += blockDim.x) {
    const int idx = t0 + threadIdx.x;
    const long long v = idx < nb ? bsum[idx] : 0;
    long long tot;
    const long long inc = block_incl_scan(v, &tot);
    if (idx < nb) bsum[idx] = carry + inc - v;
    carry += tot;
  }
}
__global__ void __launch_bounds__(SCAN_T) k_scan_final(const int* __restrict__ counts, long long n, const long long* __restrict__ bsum,
                                                       long long* __restrict__ indptr) {
  const long long base = (long long)blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  int c[SCAN_ITEMS];
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    c[k] = base + k < n ? counts[base + k] : 0;
    s += c[k];
  }
  long long tot;
  const long long inc = block_incl_scan(s, &tot);
  long long run = bsum[blockIdx.x] + inc - s;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) indptr[base + k] = run;
    run += c[k];
  }
  if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) indptr[n] = run;   // the thread that holds the last count
}

// ---- 3. fill --------------------------------------------------------------------------------------------------------
// one warp per cell of the region: for each of its five rows the 145 candidates are visited in column order, 32 at a time;
// kept entries get their position from ballot / popc and leave as one contiguous run per 32 candidates
// Block = 32 consecutive cells in i (one row j), 8 warps.  Per equation row e the 145 candidate planes are first staged into shared
// memory with COALESCED loads (one plane = 32 consecutive cells = one 256-byte warp load), then every warp compacts the rows of
// four cells: lane = candidate (column order), ballot + popc give the position, the kept run leaves contiguously.
constexpr int FILL_CELLS = 32, FILL_PITCH = FILL_CELLS + 1;   // +1: lanes walk a column of the staged tile, conflict free
__global__ void __launch_bounds__(256) k_csr_fill_interior(GridDesc g, Rect rc, const double* __restrict__ V, double thresh,
                                                           const double* __restrict__ vol, const long long* __restrict__ indptr,
                                                           int* __restrict__ indices, double* __restrict__ data) {
  __shared__ double sv[NCAND * FILL_PITCH];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ib = rc.i0 + blockIdx.x * FILL_CELLS, j = rc.j0 + blockIdx.y;
  const int nc = min(FILL_CELLS, rc.i1 - ib + 1);
  const long long ncell = (long long)g.im * g.jm;
  const long long cell0 = (long long)(ib - 1) + (long long)(j - 1) * g.im;
#pragma unroll 1
  for (int e = 0; e < 5; ++e) {
    __syncthreads();
    // asynchronous copies straight into shared memory (LDGSTS): all ~18 loads of a thread are in flight at once, no registers
    for (int c = w; c < NCAND; c += 8) {   // candidate c = (slot c / 5, variable c % 5): plane slot * 25 + e * 5 + m
      const long long plane = (long long)(c / 5) * 25 + e * 5 + c % 5;
      if (lane < nc) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&sv[c * FILL_PITCH + lane]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(V + plane * ncell + cell0 + lane) : "memory");
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    for (int q = w; q < nc; q += 8) {
      const int i = ib + q;
      const double volc = vol ? vol[g.cidx(i, j)] : 1.0;
      const int ig = i + g.ioff;
      long long pos = indptr[e + 5LL * (j - 1) + 5LL * g.jm * (i - 1)];
#pragma unroll 1
      for (int c0 = 0; c0 < NCAND; c0 += 32) {
        const int c = c0 + lane;
        double v = 0.0;
        int col = 0;
        if (c < NCAND) {
          const int s = c / 5, m = c % 5;
          v = sv[c * FILL_PITCH + q];
          col = m + 5 * (j + kSlotDj[s] - 1) + 5 * g.jm * (ig + kSlotDi[s] - 1);
        }
        const bool keep = c < NCAND && ::fabs(v) > thresh;
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const long long p = pos + __popc(mask & ((1u << lane) - 1u));
          data[p] = vol ? v / volc : v;   // true division, as the reference's Jacvol loop
          indices[p] = col;
        }
        pos += __popc(mask);
      }
    }
  }
}

__global__ void k_csr_fill_strip(const double* __restrict__ jac, const int* __restrict__ ia, const int* __restrict__ ja, long long n,
                                 double thresh, long long row0, long long nrows, const long long* __restrict__ indptr,
                                 int* __restrict__ cursor, int* __restrict__ indices, double* __restrict__ data) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double v = jac[t];
  if (::fabs(v) > thresh) {
    const long long r = ia[t] - row0;
    if (r < 0 || r >= nrows) return;
    const long long p = indptr[r] + atomicAdd(&cursor[r], 1);
    data[p] = v;
    indices[p] = ja[t];
  }
}

// one warp per strip row: rank sort by column in shared memory (columns of a row are distinct: each (row, column) pair has
// exactly one slot in the reference's colouring), then the optional division by the row cell's volume
constexpr int SORT_MAX = 256;
__global__ void __launch_bounds__(128) k_csr_sort_strip_rows(GridDesc g, RectList rl, const double* __restrict__ vol,
                                                             const long long* __restrict__ indptr, int* __restrict__ indices,
                                                             double* __restrict__ data, int* __restrict__ overflow) {
  __shared__ int scol[4][SORT_MAX];
  __shared__ double sval[4][SORT_MAX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long wrow = (long long)blockIdx.x * 4 + w;   // index among the strip rows: rect by rect, cells i fastest, 5 rows per cell
  int q = 0;
  for (; q < rl.n; ++q) {
    const long long nr = 5LL * (rl.r[q].i1 - rl.r[q].i0 + 1) * (rl.r[q].j1 - rl.r[q].j0 + 1);
    if (wrow < nr) break;
    wrow -= nr;
  }
  if (q == rl.n) return;
  const Rect rc = rl.r[q];
  const int e = (int)(wrow % 5);
  const long long cellq = wrow / 5;
  const int wi = rc.i1 - rc.i0 + 1;
  const int i = rc.i0 + (int)(cellq % wi), j = rc.j0 + (int)(cellq / wi);
  const long long row = e + 5LL * (j - 1) + 5LL * g.jm * (i - 1);
  const long long p0 = indptr[row];
  const int n = (int)(indptr[row + 1] - p0);
  if (n > SORT_MAX) {
    if (lane == 0) atomicAdd(overflow, 1);
    return;
  }
  for (int k = lane; k < n; k += 32) {
    scol[w][k] = indices[p0 + k];
    sval[w][k] = data[p0 + k];
  }
  __syncwarp();
  const double volc = vol ? vol[g.cidx(i, j)] : 1.0;
  for (int k = lane; k < n; k += 32) {
    const int ck = scol[w][k];
    int rank = 0;
    for (int u = 0; u < n; ++u) rank += (scol[w][u] < ck || (scol[w][u] == ck && u < k)) ? 1 : 0;
    indices[p0 + rank] = ck;
    data[p0 + rank] = vol ? sval[w][k] / volc : sval[w][k];
  }
}


// ---- transpose (adjoint operator): CSR of A^T from the CSR of A ------------------------------------------------------------
// The adjoint baseflow / optimal-forcing drivers take the transpose of the assembled Jacobian (cylinder.py:1090-1177,
// misc/PETSc_func.py: createTranspose): column counts by atomics, the int32 -> int64 scan above, a fill with one atomic cursor
// per transposed row, then a warp-wide rank sort of every transposed row (a row of A^T holds <= 245 entries: the columns of A a
// cell's variable appears in), so the result has ascending column indices like scipy's csr(A.T) and is deterministic.
__global__ void k_tr_count(const int* __restrict__ indices, long long nnz, int* __restrict__ counts) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < nnz) atomicAdd(&counts[indices[t]], 1);
}
// one warp per row of A: lanes stride over the row's entries
__global__ void __launch_bounds__(256) k_tr_fill(const long long* __restrict__ indptr, const int* __restrict__ indices,
                                                 const double* __restrict__ data, long long nrows, long long row0,
                                                 const long long* __restrict__ tptr, int* __restrict__ cursor, int* __restrict__ tind,
                                                 double* __restrict__ tdat) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const long long p0 = indptr[r], p1 = indptr[r + 1];
  for (long long p = p0 + lane; p < p1; p += 32) {
    const int c = indices[p];
    const int k = atomicAdd(&cursor[c], 1);
    tind[tptr[c] + k] = (int)(r + row0);
    tdat[tptr[c] + k] = data[p];
  }
}
__global__ void __launch_bounds__(128) k_tr_sort_rows(long long n, const long long* __restrict__ tptr, int* __restrict__ tind,
                                                      double* __restrict__ tdat, int* __restrict__ overflow) {
  __shared__ int scol[4][SORT_MAX];
  __shared__ double sval[4][SORT_MAX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * 4 + w;
  if (row >= n) return;
  const long long p0 = tptr[row];
  const int m = (int)(tptr[row + 1] - p0);
  if (m > SORT_MAX) {
    if (lane == 0) atomicAdd(overflow, 1);
    return;
  }
  for (int k = lane; k < m; k += 32) {
    scol[w][k] = tind[p0 + k];
    sval[w][k] = tdat[p0 + k];
  }
  __syncwarp();
  for (int k = lane; k < m; k += 32) {
    const int ck = scol[w][k];
    int rank = 0;
    for (int u = 0; u < m; ++u) rank += scol[w][u] < ck ? 1 : 0;   // (row, column) pairs of a CSR matrix are distinct
    tind[p0 + rank] = ck;
    tdat[p0 + rank] = sval[w][k];
  }
}
}  // namespace
}  // namespace bcast

using namespace bcast;

// Step 1 + 2: indptr[0 .. 5 im jm] (int64, device) of the CSR row block; counts / bsum are device work arrays of
// 5 im jm + 1 ints and ceil(5 im jm / 2048) + 1 int64.  The caller reads indptr[5 im jm] (= nnz) and allocates indices / data.
static int hybrid_csr_indptr(long long* indptr, int32_t* counts, long long* bsum, const double* values, const int32_t* region, int nstrip,
                             const double* const* sjac, const int32_t* const* sia, const long long* slen, double thresh, int gh, int im,
                             int jm, int counted, void* stream);
extern "C" int bcd_hybrid_csr_indptr(long long* indptr, int32_t* counts, long long* bsum, const double* values, const int32_t* region,
                                     int nstrip, const double* const* sjac, const int32_t* const* sia, const long long* slen, double thresh,
                                     int gh, int im, int jm, void* stream) {
  return hybrid_csr_indptr(indptr, counts, bsum, values, region, nstrip, sjac, sia, slen, thresh, gh, im, jm, 0, stream);
}
// the same with the counts of the regular rows already in `counts` (bcd_jacobian_interior_counted): only the strip rows are counted
extern "C" int bcd_hybrid_csr_indptr_counted(long long* indptr, int32_t* counts, long long* bsum, const int32_t* region, int nstrip,
                                             const double* const* sjac, const int32_t* const* sia, const long long* slen, double thresh,
                                             int gh, int im, int jm, void* stream) {
  return hybrid_csr_indptr(indptr, counts, bsum, nullptr, region, nstrip, sjac, sia, slen, thresh, gh, im, jm, 1, stream);
}
static int hybrid_csr_indptr(long long* indptr, int32_t* counts, long long* bsum, const double* values, const int32_t* region, int nstrip,
                             const double* const* sjac, const int32_t* const* sia, const long long* slen, double thresh, int gh, int im,
                             int jm, int counted, void* stream) {
  if (im < 1 || jm < 1 || gh != 3 || nstrip < 0 || nstrip > 4) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  // limits of this layout: one grid row per j (65 535), 32-bit column indices (5 * im_global * jm < 2^31)
  if (jm > 65535 || 5LL * g.img * jm >= (1LL << 31)) return BC_ERR_UNSUPPORTED;
  const long long n = 5LL * im * jm;
  const long long row0 = 5LL * jm * g.ioff;
  cudaError_t e = counted ? cudaSuccess : cudaMemsetAsync(counts, 0, sizeof(int) * (n + 1), st);
  if (e != cudaSuccess) return (int)e;
  const Rect rc{region[0], region[1], region[2], region[3]};
  if (!counted && rc.i1 >= rc.i0 && rc.j1 >= rc.j0 && values)
    k_csr_count_interior<<<dim3((rc.i1 - rc.i0 + 128) / 128, rc.j1 - rc.j0 + 1), 128, 0, st>>>(g, rc, values, thresh, counts);
  for (int q = 0; q < nstrip; ++q)
    if (slen[q] > 0) k_csr_count_strip<<<(unsigned)((slen[q] + 255) / 256), 256, 0, st>>>(sjac[q], sia[q], slen[q], thresh, row0, n, counts);
  const int nb = (int)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
  k_scan_reduce<<<nb, SCAN_T, 0, st>>>(counts, n, bsum);
  k_scan_blocks<<<1, 1024, 0, st>>>(bsum, nb);
  k_scan_final<<<nb, SCAN_T, 0, st>>>(counts, n, bsum, indptr);
  count_launches(4 + nstrip);
  e = cudaGetLastError();
  return e == cudaSuccess ? BC_OK : (int)e;
}

// Step 3: indices (int32) and data of the rows, columns ascending inside a row; vol (cell layout, or null) divides every row by
// the volume of its cell; cursor = 5 im jm ints of work space.  Returns BC_ERR_UNSUPPORTED if a strip row holds more than 256
// entries (cannot happen with the reference's colouring: 245 slots per row).
extern "C" int bcd_hybrid_csr_fill(int32_t* indices, double* data, int32_t* cursor, const long long* indptr, const double* values,
                                   const int32_t* region, int nstrip, const int32_t* srect /* [nstrip][4] */, const double* const* sjac,
                                   const int32_t* const* sia, const int32_t* const* sja, const long long* slen, double thresh,
                                   const double* vol, int gh, int im, int jm, void* stream) {
  if (im < 1 || jm < 1 || gh != 3 || nstrip < 0 || nstrip > 4) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const long long n = 5LL * im * jm;
  const long long row0 = 5LL * jm * g.ioff;
  const Rect rc{region[0], region[1], region[2], region[3]};
  if (rc.i1 >= rc.i0 && rc.j1 >= rc.j0 && values) {
    k_csr_fill_interior<<<dim3((rc.i1 - rc.i0 + FILL_CELLS) / FILL_CELLS, rc.j1 - rc.j0 + 1), 256, 0, st>>>(g, rc, values, thresh, vol, indptr,
                                                                                                      indices, data);
  }
  if (nstrip > 0) {
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(int) * (n + 1), st);
    if (e != cudaSuccess) return (int)e;
    RectList rl;
    rl.n = nstrip;
    long long nrows = 0;
    for (int q = 0; q < 4; ++q) {
      rl.r[q] = q < nstrip ? Rect{srect[4 * q], srect[4 * q + 1], srect[4 * q + 2], srect[4 * q + 3]} : Rect{1, 0, 1, 0};
      if (q < nstrip) nrows += 5LL * (rl.r[q].i1 - rl.r[q].i0 + 1) * (rl.r[q].j1 - rl.r[q].j0 + 1);
    }
    for (int q = 0; q < nstrip; ++q)
      if (slen[q] > 0)
        k_csr_fill_strip<<<(unsigned)((slen[q] + 255) / 256), 256, 0, st>>>(sjac[q], sia[q], sja[q], slen[q], thresh, row0, n, indptr,
                                                                           cursor, indices, data);
    // cursor[n] doubles as the overflow flag of the sort (zeroed by the memset above)
    k_csr_sort_strip_rows<<<(unsigned)((nrows + 3) / 4), 128, 0, st>>>(g, rl, vol, indptr, indices, data, cursor + n);
    int ovf = 0;
    e = cudaMemcpyAsync(&ovf, cursor + n, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    if (ovf) return BC_ERR_UNSUPPORTED;
  }
  count_launches(2 + nstrip);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? BC_OK : (int)e;
}

// Transpose of a CSR row block (rows row0 .. row0 + nrows - 1 of a matrix with ncols columns; row0 = 0 and nrows = ncols for a whole
// matrix): step 1 fills tptr[0 .. ncols] (int64), the caller reads tptr[ncols] (= nnz) and allocates tind / tdat; step 2 fills
// them, columns (= rows of A) ascending.  counts / cursor: ncols + 1 ints of work space each, bsum: ceil(ncols / 2048) + 1 int64.
extern "C" int bcd_csr_transpose_indptr(long long* tptr, int32_t* counts, long long* bsum, const int32_t* indices, long long nnz,
                                        long long ncols, void* stream) {
  if (ncols < 1 || nnz < 0) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * (ncols + 1), st);
  if (e != cudaSuccess) return (int)e;
  if (nnz > 0) k_tr_count<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(indices, nnz, counts);
  const int nb = (int)((ncols + SCAN_CHUNK - 1) / SCAN_CHUNK);
  k_scan_reduce<<<nb, SCAN_T, 0, st>>>(counts, ncols, bsum);
  k_scan_blocks<<<1, 1024, 0, st>>>(bsum, nb);
  k_scan_final<<<nb, SCAN_T, 0, st>>>(counts, ncols, bsum, tptr);
  count_launches(4);
  e = cudaGetLastError();
  return e == cudaSuccess ? BC_OK : (int)e;
}
extern "C" int bcd_csr_transpose_fill(int32_t* tind, double* tdat, int32_t* cursor, const long long* tptr, const long long* indptr,
                                      const int32_t* indices, const double* data, long long nrows, long long row0, long long ncols,
                                      void* stream) {
  if (ncols < 1 || nrows < 0 || row0 < 0 || row0 + nrows >= (1LL << 31)) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(int) * (ncols + 1), st);
  if (e != cudaSuccess) return (int)e;
  if (nrows > 0) k_tr_fill<<<(unsigned)((nrows * 32 + 255) / 256), 256, 0, st>>>(indptr, indices, data, nrows, row0, tptr, cursor, tind, tdat);
  k_tr_sort_rows<<<(unsigned)((ncols + 3) / 4), 128, 0, st>>>(ncols, tptr, tind, tdat, cursor + ncols);
  int ovf = 0;
  e = cudaMemcpyAsync(&ovf, cursor + ncols, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return (int)e;
  count_launches(2);
  return ovf ? BC_ERR_UNSUPPORTED : BC_OK;
}
