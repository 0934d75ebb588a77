// Block-Jacobian assembly of the regular interior rows by face linearisation (facejac.cuh):
//   k_face_packages  one thread per cell: linearisation packages of its i-face and j-face -> HBM (SoA planes)
//   k_jac_assemble   one thread per row cell: the 29 structural 5x5 blocks from the packages of its four faces,
//                    written as values[slot][e*5+m][cell] (coalesced 8-byte stores, cell = (i-1) + (j-1)*im)
// Replaces the 29 direct AD kernels of jac_blocks.cuh (kept as a cross-check: bcd_jacobian_interior(..., method=1)).
#include "../../include/broadcast_b200.h"
#include "facejac.cuh"
#include "kernels.cuh"
#include <cstdlib>

namespace bcast {
void count_launches(int n);
cudaError_t prepare_prims_grads(const GridDesc& g, const SchemeArgs& a, const double* w, const double* nx, const double* ny,
                                const double* vol, const double* volf, FieldPtrs& f, cudaStream_t st);

// blockIdx.z = face direction: one face package per thread (half the work per thread of the both-faces version, twice the threads)
__global__ void __launch_bounds__(128) k_face_packages(GridDesc g, SchemeConsts c, FieldPtrs f, Rect rc, double* __restrict__ pkg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + rc.i0;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + rc.j0;
  if (i > rc.i1 || j > rc.j1) return;
  const GlobalAcc<0> a(f, g, i, j);
  const long long k = g.cidx(i, j);
  const long long sc = g.sc;
  if (blockIdx.z == 0) {
    double* p0 = pkg + k;
    face_package<0>(a, c, [&](int fld, double v) { p0[fld * sc] = v; });
  } else {
    double* p1 = pkg + (long long)FPK_N * g.sc + k;
    face_package<1>(a, c, [&](int fld, double v) { p1[fld * sc] = v; });
  }
}

template <int DIR>
__device__ __forceinline__ FaceCtx make_ctx(const FieldPtrs& f, const GridDesc& g, const double* pkg, int i, int j) {
  FaceCtx x;
  x.pk = pkg + (long long)DIR * FPK_N * g.sc + g.cidx(i, j);
  x.stride = g.sc;
  x.vs = x.pk + (long long)FPK_VS * g.sc;
  x.vstride = (int)g.sc;
  return x;
}

__global__ void __launch_bounds__(128) k_jac_assemble(GridDesc g, SchemeConsts c, FieldPtrs f, Rect rc, const double* __restrict__ pkg,
                                                      double* __restrict__ V, const double* __restrict__ coefdiag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + rc.i0;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + rc.j0;
  if (i > rc.i1 || j > rc.j1) return;
  const FaceCtx fi0 = make_ctx<0>(f, g, pkg, i, j), fi1 = make_ctx<0>(f, g, pkg, i + 1, j);
  const FaceCtx fj0 = make_ctx<1>(f, g, pkg, i, j), fj1 = make_ctx<1>(f, g, pkg, i, j + 1);
  const long long ncell = (long long)g.im * g.jm;
  const long long cell = (long long)(i - 1) + (long long)(j - 1) * g.im;
  const double cd = coefdiag ? coefdiag[cell] : 0.0;
  int slot = 0;
#define X(DI, DJ)                                                                       \
  {                                                                                     \
    double wc[5], B[25];                                                                \
    const long long kc = g.cidx(i + (DI), j + (DJ));                                    \
    _Pragma("unroll") for (int e = 0; e < 5; ++e) wc[e] = __ldg(f.w + e * g.sc + kc);   \
    block_of<DI, DJ>(fi0, fi1, fj0, fj1, wc, c, B);                                     \
    if (DI == 0 && DJ == 0) {                                                           \
      _Pragma("unroll") for (int e = 0; e < 5; ++e) B[e * 6] += cd;                     \
    }                                                                                   \
    double* out = V + ((long long)slot * 25) * ncell + cell;                            \
    _Pragma("unroll") for (int q = 0; q < 25; ++q) out[q * ncell] = B[q];               \
    ++slot;                                                                             \
  }
  BCAST_JAC_OFFSETS(X)
#undef X
}

__constant__ JacTab kJacTabDev = fj::make_jac_tab();

// table-driven assembly: one runtime loop over the 29 column slots (see facejac.cuh).  CTA = 32 x 4 row cells; the
// staged subset (27 of the 54 fields: everything that more than two column cells of a face read) of the packages of its
// 33x4 i-faces and 32x5 j-faces lives in shared memory (62.8 KB); ncu of the 14-field version: 1 160 package LDGs per cell
// at 14 % L1 hit rate, long_scoreboard 6.6 per issue (profiles/r1_d_summary.md).
// ST = first staged field (27: DEG + SE + viscous = 27 fields, 62.8 KB; 35: SE + viscous = 19 fields, 44.2 KB; 40: viscous only),
// MINB = CTAs per SM the register allocation aims at.
constexpr int JT_I = 32, JT_J = 4;
// COUNT: the thread also counts, per equation row of its cell, the entries the CSR conversion will keep (|v| > thresh, the
// reference's remove_zero_jac) while the block values are still in registers: the conversion's counting pass over the 5.8 KB of
// block values per cell (csr.cu: k_csr_count_interior) is then not needed.
template <int ST, int MINB, bool COUNT = false>
__global__ void __launch_bounds__(JT_I* JT_J, MINB)
    k_jac_assemble_rt(GridDesc g, SchemeConsts c, FieldPtrs f, Rect rc, const double* __restrict__ pkg, double* __restrict__ V,
                      const double* __restrict__ coefdiag, int* __restrict__ counts = nullptr, double thresh = 0.0) {
  constexpr int FPK_NVS = FPK_N - ST;
  constexpr int FPK_VS = ST;
  using FaceCtx = FaceCtxT<ST>;
  extern __shared__ double jsm[];
  double* sI = jsm;
  double* sJ = jsm + FPK_NVS * JT_J * (JT_I + 1);
  const int tid = threadIdx.y * JT_I + threadIdx.x;
  const int bi0 = blockIdx.x * JT_I + rc.i0, bj0 = blockIdx.y * JT_J + rc.j0;
  const double* pI = pkg + (long long)FPK_VS * g.sc;
  const double* pJ = pkg + (long long)(FPK_N + FPK_VS) * g.sc;
  // staging by asynchronous copies straight into shared memory (LDGSTS): every copy of a thread is in flight at once and no
  // register holds staged data (r2_c; the register-staged loop -- eight loads, then eight stores -- held 8 % of the kernel's stall
  // samples on its first store, ncu r2_30)
  constexpr int NTH = JT_I * JT_J;
  {
    constexpr int NI = FPK_NVS * JT_J * (JT_I + 1);
    for (int idx = tid; idx < NI; idx += NTH) {
      const int k = idx / (JT_J * (JT_I + 1)), r = idx % (JT_J * (JT_I + 1));
      const int fi = bi0 + r % (JT_I + 1), fj = bj0 + r / (JT_I + 1);
      if (fi <= rc.i1 + 1 && fj <= rc.j1) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(sI + idx);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(pI + k * g.sc + g.cidx(fi, fj)) : "memory");
      } else {
        sI[idx] = 0.0;
      }
    }
  }
  {
    constexpr int NJ = FPK_NVS * (JT_J + 1) * JT_I;
    for (int idx = tid; idx < NJ; idx += NTH) {
      const int k = idx / ((JT_J + 1) * JT_I), r = idx % ((JT_J + 1) * JT_I);
      const int fi = bi0 + r % JT_I, fj = bj0 + r / JT_I;
      if (fi <= rc.i1 && fj <= rc.j1 + 1) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(sJ + idx);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(pJ + k * g.sc + g.cidx(fi, fj)) : "memory");
      } else {
        sJ[idx] = 0.0;
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = bi0 + tx, j = bj0 + ty;
  if (i > rc.i1 || j > rc.j1) return;
  const double* pk = pkg + g.cidx(i, j);
  const long long dj = (long long)FPK_N * g.sc;
  const FaceCtx fi0{pk, g.sc, sI + ty * (JT_I + 1) + tx, JT_J * (JT_I + 1)};
  const FaceCtx fi1{pk + 1, g.sc, sI + ty * (JT_I + 1) + tx + 1, JT_J * (JT_I + 1)};
  const FaceCtx fj0{pk + dj, g.sc, sJ + ty * JT_I + tx, (JT_J + 1) * JT_I};
  const FaceCtx fj1{pk + dj + g.ldc, g.sc, sJ + (ty + 1) * JT_I + tx, (JT_J + 1) * JT_I};
  const long long ncell = (long long)g.im * g.jm;
  const long long cell = (long long)(i - 1) + (long long)(j - 1) * g.im;
  const double cd = coefdiag ? coefdiag[cell] : 0.0;
  int cnt[5] = {0, 0, 0, 0, 0};
  double wn[5];   // state of the next slot's column cell: loaded one slot ahead
  {
    const long long kc = g.cidx(i + kJacTabDev.di[0], j + kJacTabDev.dj[0]);
#pragma unroll
    for (int e = 0; e < 5; ++e) wn[e] = __ldg(f.w + e * g.sc + kc);
  }
#pragma unroll 1
  for (int s = 0; s < JAC_NSLOT; ++s) {
    double wc[5], B[25];
#pragma unroll
    for (int e = 0; e < 5; ++e) wc[e] = wn[e];
    if (s + 1 < JAC_NSLOT) {
      const long long kc = g.cidx(i + kJacTabDev.di[s + 1], j + kJacTabDev.dj[s + 1]);
#pragma unroll
      for (int e = 0; e < 5; ++e) wn[e] = __ldg(f.w + e * g.sc + kc);
    }
    block_of_rt(kJacTabDev, s, fi0, fi1, fj0, fj1, wc, c, B);
    if (kJacTabDev.di[s] == 0 && kJacTabDev.dj[s] == 0) {
#pragma unroll
      for (int e = 0; e < 5; ++e) B[e * 6] += cd;
    }
    double* out = V + ((long long)s * 25) * ncell + cell;
#pragma unroll
    for (int q = 0; q < 25; ++q) out[q * ncell] = B[q];
    if constexpr (COUNT) {
#pragma unroll
      for (int q = 0; q < 25; ++q) cnt[q / 5] += ::fabs(B[q]) > thresh ? 1 : 0;
    }
  }
  if constexpr (COUNT) {
    const long long row = 5LL * (j - 1) + 5LL * g.jm * (i - 1);
#pragma unroll
    for (int e = 0; e < 5; ++e) counts[row + e] = cnt[e];
  }
}

cudaError_t launch_jacobian_faces(const GridDesc& g, const SchemeArgs& a, const double* w, const double* nx, const double* ny,
                                  const double* vol, const double* volf, const Rect& rc, double* values, const double* coefdiag,
                                  cudaStream_t st, int* counts, double thresh) {
  const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
  FieldPtrs f;
  cudaError_t e = prepare_prims_grads(g, a, w, nx, ny, vol, volf, f, st);
  if (e != cudaSuccess) return e;
  double* pkg = scratch_doubles(30, (size_t)2 * FPK_N * g.sc);
  if (!pkg) return cudaErrorMemoryAllocation;
  const Rect rf{rc.i0, rc.i1 + 1, rc.j0, rc.j1 + 1};
  dim3 blk(32, 4);
  dim3 gf((rf.i1 - rf.i0 + 32) / 32, (rf.j1 - rf.j0 + 4) / 4, 2);
  k_face_packages<<<gf, blk, 0, st>>>(g, c, f, rf, pkg);
  dim3 gr((rc.i1 - rc.i0 + 32) / 32, (rc.j1 - rc.j0 + 4) / 4);
  static const bool unrolled = getenv("BROADCAST_B200_JAC_UNROLLED") != nullptr;   // template-unrolled variant (cross-check)
  if (unrolled)
    k_jac_assemble<<<gr, blk, 0, st>>>(g, c, f, rc, pkg, values, coefdiag);
  else
  {
    static const int cfg = getenv("BROADCAST_B200_JAC_CFG") ? atoi(getenv("BROADCAST_B200_JAC_CFG")) : 5;
    const dim3 grd((rc.i1 - rc.i0 + JT_I) / JT_I, (rc.j1 - rc.j0 + JT_J) / JT_J), blk2(JT_I, JT_J);
    auto go = [&](auto kern, int st_field, bool counted = false) -> cudaError_t {
      const size_t smem = (size_t)(FPK_N - st_field) * (JT_J * (JT_I + 1) + (JT_J + 1) * JT_I) * sizeof(double);
      cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e2 != cudaSuccess) return e2;
      e2 = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
      if (e2 != cudaSuccess) return e2;
      if (counted) kern<<<grd, blk2, smem, st>>>(g, c, f, rc, pkg, values, coefdiag, counts, thresh);
      else kern<<<grd, blk2, smem, st>>>(g, c, f, rc, pkg, values, coefdiag, nullptr, 0.0);
      return cudaGetLastError();
    };
    // default since r2_c: 33 staged fields (FPK_LS: the along-line / face-value fields too, 77 KB, two CTAs per SM), 255 registers
    static const bool st27 = cfg == 6;
    if (counts) e = st27 ? go(k_jac_assemble_rt<27, 2, true>, 27, true) : go(k_jac_assemble_rt<FPK_LS, 2, true>, FPK_LS, true);
    else if (cfg == 5) e = go(k_jac_assemble_rt<FPK_LS, 2>, FPK_LS);
    else
    if (cfg == 1) e = go(k_jac_assemble_rt<35, 4>, 35);        // 19 staged fields, 128 registers, 16 warps per SM
    else if (cfg == 2) e = go(k_jac_assemble_rt<40, 4>, 40);   // 14 staged fields, 128 registers
    else if (cfg == 3) e = go(k_jac_assemble_rt<40, 5>, 40);   // 14 staged fields, 96 registers, 20 warps per SM
    else if (cfg == 4) e = go(k_jac_assemble_rt<35, 3>, 35);   // 19 staged fields, 168 registers
    else if (cfg == 0) e = go(k_jac_assemble_rt<27, 3>, 27);   // 27 staged fields, 168 registers, 12 warps per SM
    else e = go(k_jac_assemble_rt<27, 2>, 27);                 // cfg 6 (the r1 / r2_b default): 27 staged fields, 255 registers, 8 warps per SM
    // measured at 4096x1024 (profiles/r1_g_summary.md, r1_h): 12.17 ms (168 registers) vs 11.43 ms (255 registers)
    if (e != cudaSuccess) return e;
  }
  count_launches(5);
  return cudaGetLastError();
}

}  // namespace bcast
