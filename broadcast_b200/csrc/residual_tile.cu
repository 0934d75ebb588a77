// Fused primal residual for sm_100a: one CTA per TI x TJ tile of cells.
//   stage 0  the tile of w plus its gh-wide halo is staged into shared memory (5 planes)
//   stage 1  primitives of every staged cell are computed once into shared memory
//   stage 2  5-point velocity gradients for the (TI+2) x (TJ+2) cells the sensor reads, including
//            the extrapolated first ghost layer at physical boundaries
//   stage 3  i-face fluxes -> shared exchange buffer -> partial balance in registers;
//            j-face fluxes -> same buffer            -> final balance, coalesced store of residu
// The face formulas are the templates of scheme.cuh evaluated on a shared-memory accessor, so this
// kernel computes exactly what the generic path (generic_impl.cuh) computes, without the global
// prims/gradient arrays (HBM traffic ~ the algorithmic 136 B per cell plus halo re-reads from L2).
// Reference: srcfv/rhs/flux_num_dnc5.F90:7-226.
#include "kernels.cuh"
#include <cstdlib>

namespace bcast {

namespace {

constexpr int H = 3;  // halo width = gh for order 5

// Thread tile TI x TJ: thread (tx,ty) owns cell (i0+tx, j0+ty), evaluates that cell's LEFT i-face and BOTTOM j-face
// (exactly two face fluxes per thread, no serial tail), and the CTA writes the (TI-1) x (TJ-1) cells whose four faces
// it holds.  Staged cells: i0-3 .. i0+TI+1, j0-3 .. j0+TJ+1.
template <int TI, int TJ>
struct Tile {
  static constexpr int OI = TI - 1, OJ = TJ - 1;  // output cells per CTA
  static constexpr int PI = TI + 2 * H - 1;
  static constexpr int PJ = TJ + 2 * H - 1;
  static constexpr int NC = PI * PJ;
  static constexpr int NARR = 14;  // w(5) u v wz T p mu h divu ducros
  static constexpr int XP = TI + 1;  // exchange pitch
  static constexpr int NX = 5 * TJ * XP;
  static constexpr size_t SMEM = (size_t)(NARR * NC + NX) * sizeof(double);
};
enum { A_W = 0, A_U = 5, A_V = 6, A_WZ = 7, A_T = 8, A_P = 9, A_MU = 10, A_H = 11, A_G = 12 };

template <int PI, int NC>
struct SmemAcc {
  using DT = Zero;
  const double* s;
  int k;  // shared index of the base cell
  const double *nx, *ny, *vol, *volf;
  long long c, n;
  int ldc, ldn;
  long long sc, sn;
  template <int OI, int OJ> __device__ __forceinline__ PVar ld(int arr) const { return PVar{s[arr * NC + k + OI + OJ * PI], {}}; }
  template <int OI, int OJ> __device__ __forceinline__ PVar W(int e) const { return ld<OI, OJ>(A_W + e); }
  template <int OI, int OJ> __device__ __forceinline__ PVar U() const { return ld<OI, OJ>(A_U); }
  template <int OI, int OJ> __device__ __forceinline__ PVar V() const { return ld<OI, OJ>(A_V); }
  template <int OI, int OJ> __device__ __forceinline__ PVar Wz() const { return ld<OI, OJ>(A_WZ); }
  template <int OI, int OJ> __device__ __forceinline__ PVar T() const { return ld<OI, OJ>(A_T); }
  template <int OI, int OJ> __device__ __forceinline__ PVar P() const { return ld<OI, OJ>(A_P); }
  template <int OI, int OJ> __device__ __forceinline__ PVar Mu() const { return ld<OI, OJ>(A_MU); }
  template <int OI, int OJ> __device__ __forceinline__ PVar H() const { return ld<OI, OJ>(A_H); }
  // per-cell sensor quantities precomputed in stage 2: dilatation and Ducros ratio
  template <int OI, int OJ> __device__ __forceinline__ auto SENS() const { return CellSens<Zero, Zero>{ld<OI, OJ>(A_G), ld<OI, OJ>(A_G + 1)}; }
  template <int OI, int OJ> __device__ __forceinline__ double NX(int kk) const { return __ldg(nx + kk * sn + n + OI + (long long)OJ * ldn); }
  template <int OI, int OJ> __device__ __forceinline__ double NY(int kk) const { return __ldg(ny + kk * sn + n + OI + (long long)OJ * ldn); }
  template <int OI, int OJ> __device__ __forceinline__ double VOL() const { return __ldg(vol + c + OI + (long long)OJ * ldc); }
  template <int OI, int OJ> __device__ __forceinline__ double VOLF(int kk) const { return __ldg(volf + kk * sc + c + OI + (long long)OJ * ldc); }
};

template <int TI, int TJ>
__global__ void __launch_bounds__(TI* TJ, 3)
    k_residual_tile(GridDesc g, SchemeConsts cst_, bool wall, const double* __restrict__ w, const double* __restrict__ nx,
                    const double* __restrict__ ny, const double* __restrict__ vol, const double* __restrict__ volf,
                    double* __restrict__ res) {
  using TL = Tile<TI, TJ>;
  constexpr int PI = TL::PI, PJ = TL::PJ, NC = TL::NC, NT = TI * TJ, XP = TL::XP;
  extern __shared__ double sm[];
  double* X = sm + TL::NARR * NC;
  const int tid = threadIdx.x;
  const int tx = tid % TI, ty = tid / TI;
  const int i0 = 1 + blockIdx.x * TL::OI, j0 = 1 + blockIdx.y * TL::OJ;
  const int im = g.im, jm = g.jm;
  const SchemeConsts c = cst_;
  using Acc = SmemAcc<PI, NC>;
  auto make_acc = [&](int a, int b) {  // shared-tile coordinates (a,b): cell (i0-H+a, j0-H+b)
    Acc A;
    A.s = sm;
    A.k = a + b * PI;
    A.nx = nx; A.ny = ny; A.vol = vol; A.volf = volf;
    A.c = g.cidx(i0 - H + a, j0 - H + b);
    A.n = g.nidx(i0 - H + a, j0 - H + b);
    A.ldc = g.ldc; A.ldn = g.ldn; A.sc = g.sc; A.sn = g.sn;
    return A;
  };

  // ---- stage 0 + 1: stage w, compute primitives ------------------------------------------------
  for (int idx = tid; idx < NC; idx += NT) {
    const int a = idx % PI, b = idx / PI;
    const int gi = i0 - H + a, gj = j0 - H + b;
    const bool inb = gi <= im + g.gh && gj <= jm + g.gh;
    PVar q[5];
    if (inb) {
      const long long k = g.cidx(gi, gj);
#pragma unroll
      for (int e = 0; e < 5; ++e) q[e].v = __ldg(w + e * g.sc + k);
    } else {
      q[0].v = 1.0; q[1].v = 0.0; q[2].v = 0.0; q[3].v = 0.0; q[4].v = 1.0;
    }
    const CellPrims<Zero> p = cell_prims(q, c);
#pragma unroll
    for (int e = 0; e < 5; ++e) sm[(A_W + e) * NC + idx] = q[e].v;
    sm[A_U * NC + idx] = p.u.v;
    sm[A_V * NC + idx] = p.v.v;
    sm[A_WZ * NC + idx] = p.w.v;
    sm[A_T * NC + idx] = p.t.v;
    sm[A_P * NC + idx] = p.p.v;
    sm[A_MU * NC + idx] = p.mu.v;
    sm[A_H * NC + idx] = p.h.v;
  }
  __syncthreads();

  // ---- stage 2: gradients on interior cells of [i0-1, i0+TI] x [j0-1, j0+TJ] ---------------------
  constexpr int GI = TI + 1, GJ = TJ + 1;
  for (int idx = tid; idx < GI * GJ; idx += NT) {
    const int a = idx % GI + (H - 1), b = idx / GI + (H - 1);
    const int ci = i0 - H + a, cj = j0 - H + b;
    if (ci >= g.glo() && ci <= g.ghi() && cj >= 1 && cj <= jm) {   // slab-internal edges: real gradients in the halo column
      const Acc A = make_acc(a, b);
      const auto r = cell_gradients<0, 0>(A);
      sm[(A_G + 0) * NC + A.k] = (r.u0 + r.v1).v;   // dilatation
      sm[(A_G + 1) * NC + A.k] = (r.v0 - r.u1).v;   // vorticity (replaced by the Ducros ratio below)
    }
  }
  __syncthreads();
  // first ghost layer by linear extrapolation (rhs/gradveloingh.F:1-19)
  for (int idx = tid; idx < GI * GJ; idx += NT) {
    const int a = idx % GI + (H - 1), b = idx / GI + (H - 1);
    const int ci = i0 - H + a, cj = j0 - H + b;
    const int k = a + b * PI;
    int d = 0;
    if (cj >= 1 && cj <= jm) {
      if (ci == 0 && !(g.edges & 1)) d = 1;
      else if (ci == im + 1 && !(g.edges & 2)) d = -1;
    } else if (ci >= 1 && ci <= im) {
      if (cj == 0) d = PI;
      else if (cj == jm + 1) d = -PI;
    }
    if (d != 0) {   // divu and vort are linear in the gradients: extrapolating them = extrapolating the gradients
#pragma unroll
      for (int q = 0; q < 2; ++q) sm[(A_G + q) * NC + k] = 2.0 * sm[(A_G + q) * NC + k + d] - sm[(A_G + q) * NC + k + 2 * d];
    }
  }
  __syncthreads();
  // Ducros ratio of every sensor cell (once per cell instead of once per face side)
  for (int idx = tid; idx < GI * GJ; idx += NT) {
    const int a = idx % GI + (H - 1), b = idx / GI + (H - 1);
    const int k = a + b * PI;
    const auto cs = sens_from_divu_vort(PVar{sm[A_G * NC + k], {}}, PVar{sm[(A_G + 1) * NC + k], {}});
    sm[(A_G + 1) * NC + k] = cs.ducros.v;
  }
  __syncthreads();

  double r[5] = {0.0, 0.0, 0.0, 0.0, 0.0};

  // ---- stage 3a: i-faces (rows ty < TJ-1) ----------------------------------------------------------
  {
    const int fi = i0 + tx, fj = j0 + ty;
    if (ty < TJ - 1 && fi <= im + 1 && fj <= jm) {
      const Acc A = make_acc(tx + H, ty + H);
      PVar hn[5];
      if (wall && fj <= 2)
        face_flux<0, true, FACE_MAIN>(A, c, hn);
      else
        face_flux<0, false, FACE_MAIN>(A, c, hn);
#pragma unroll
      for (int e = 0; e < 5; ++e) X[(e * TJ + ty) * XP + tx] = hn[e].v;
    }
  }
  __syncthreads();
  const bool mine = tx < TI - 1 && ty < TJ - 1 && i0 + tx <= im && j0 + ty <= jm;
  if (mine) {
#pragma unroll
    for (int e = 0; e < 5; ++e) r[e] = -(X[(e * TJ + ty) * XP + tx + 1] - X[(e * TJ + ty) * XP + tx]);
  }
  __syncthreads();

  // ---- stage 3b: j-faces (columns tx < TI-1) -------------------------------------------------------
  {
    const int fi = i0 + tx, fj = j0 + ty;
    if (tx < TI - 1 && fi <= im && fj <= jm + 1) {
      const Acc A = make_acc(tx + H, ty + H);
      PVar hn[5];
      if (wall && fj == 1)
        face_flux<1, true, FACE_WALL>(A, c, hn);
      else if (wall && fj == 2)
        face_flux<1, true, FACE_NEAR3>(A, c, hn);
      else if (wall && fj == 3)
        face_flux<1, false, FACE_NEAR5>(A, c, hn);
      else
        face_flux<1, false, FACE_MAIN>(A, c, hn);
#pragma unroll
      for (int e = 0; e < 5; ++e) X[(e * TJ + ty) * XP + tx] = hn[e].v;
    }
  }
  __syncthreads();
  if (mine) {
    const long long k = g.cidx(i0 + tx, j0 + ty);
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      const double v = r[e] - (X[(e * TJ + ty + 1) * XP + tx] - X[(e * TJ + ty) * XP + tx]);
      res[e * g.sc + k] = v;
    }
  }
}

}  // namespace

// which kernel RES_DEFAULT names for whole-block launches: decided by measurement (profiles/r2_b_summary.md)
constexpr bool BULK_IS_DEFAULT = true;   // C5: 2.25 ms against 2.39 ms for the LDG tile kernel (first bulk version, r2_11)

cudaError_t launch_residual_tiled(const GridDesc& g, const SchemeArgs& a, bool wall, double* res, const double* w, const double* nx,
                                  const double* ny, const double* vol, const double* volf, cudaStream_t st, int variant, int part) {
  if (g.im < 4 || g.jm < 6 || (part != 0 && (variant == RES_TILE_V1 || variant == RES_FAST_TMA || variant == RES_MARCH))) {
    if (part == 1) return cudaSuccess;   // kernels without a tile split do everything in the "ring" call
    if (g.im < 4 || g.jm < 6) return launch_residual_generic(g, a, wall, 0, res, w, nullptr, nx, ny, vol, volf, nullptr, st);
    part = 0;
  }
  // second-generation kernel (residual_fast.cu) unless the first one is asked for as a cross-check
  static const bool v1 = getenv("BROADCAST_B200_RESIDUAL_V1") != nullptr;
  static const bool tma = getenv("BROADCAST_B200_RESIDUAL_TMA") != nullptr;
  static const bool march_default = getenv("BROADCAST_B200_RESIDUAL_MARCH") != nullptr;
  if ((variant == RES_MARCH || (variant == RES_DEFAULT && march_default && !v1 && !tma)) && part == 0) {
    // the marching kernel; falls through to the tile kernel when TMA cannot describe the arrays
    const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
    bool done = false;
    cudaError_t e = launch_residual_march(g, c, ::sqrt(a.gam * a.rgaz), wall, res, w, nx, ny, vol, volf, st, &done);
    if (done || e != cudaSuccess) return e;
  }
  // bulk-staged tile kernel: the default for whole-block launches (BROADCAST_B200_RESIDUAL_LDG=1 keeps the LDG tile kernel)
  static const bool ldg_default = getenv("BROADCAST_B200_RESIDUAL_LDG") != nullptr;
  if (variant == RES_FAST_BULK || (variant == RES_DEFAULT && BULK_IS_DEFAULT && !ldg_default && !v1 && !tma && !march_default)) {
    // part 1 (inner tiles) runs on the bulk kernel too; part 2 (the ring) falls through to the LDG kernel's ring launch
    const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
    bool done = false;
    cudaError_t e = launch_residual_fast_bulk(g, c, ::sqrt(a.gam * a.rgaz), wall, res, w, nx, ny, vol, volf, st, &done, part);
    if (done || e != cudaSuccess) return e;
    // part 1 not launched by the bulk kernel (TMA cannot describe the arrays, or no inner tile): the LDG kernel takes it
  }
  if (variant != RES_TILE_V1 && !(variant == RES_DEFAULT && v1))
    return launch_residual_fast(g, a, wall, res, w, nx, ny, vol, volf, st, variant == RES_FAST_TMA || (variant == RES_DEFAULT && tma), part);
  constexpr int TI = 32, TJ = 8;
  using TL = Tile<TI, TJ>;
  const SchemeConsts c = make_consts(a.cp, a.cv, a.prandtl, a.gam, a.rgaz, a.cs, a.muref, a.tref, a.s_suth, a.k2, a.k4);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_residual_tile<TI, TJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL::SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((g.im + TL::OI - 1) / TL::OI, (g.jm + TL::OJ - 1) / TL::OJ);
  k_residual_tile<TI, TJ><<<grid, TI * TJ, TL::SMEM, st>>>(g, c, wall, w, nx, ny, vol, volf, res);
  return cudaGetLastError();
}

}  // namespace bcast
