// Generic (reference-shaped) CUDA kernels of the BROADCAST hot path for sm_100a:
//   primitives -> 5-point gradients (+ ghost-layer extrapolation) -> cell-centred flux balance,
// in passive and vector-tangent arithmetic, plus the boundary fills, the colouring seeds, the COO
// scatter and the residual norms.  The fast paths (residual_tile.cu, jacobian.cu) are validated
// against these kernels, and these against the CPU oracle.
#include <cstdio>
#include <map>
#include <mutex>
#include <vector>
#include "kernels.cuh"

namespace bcast {

// ---------------------------------------------------------------------------------------------
// scratch arena
// ---------------------------------------------------------------------------------------------
namespace {
struct Slot {
  double* p = nullptr;
  size_t n = 0;
};
std::mutex g_scratch_mu;
std::map<std::pair<int, int>, Slot> g_scratch;  // (device, slot)
}  // namespace

SlabInfo& current_slab() {
  static thread_local SlabInfo s{0, 0, 0};
  return s;
}

int& scratch_chain() {
  static thread_local int c = 0;
  return c;
}
RectFork& rect_fork() {
  static thread_local RectFork f[4];
  return f[scratch_chain() & 3];
}
bool RectFork::ready() {
  if (fork) return true;
  for (int k = 0; k < 4; ++k) {
    if (cudaStreamCreateWithFlags(&side[k], cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&join[k], cudaEventDisableTiming) != cudaSuccess) return false;
  }
  return cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess;
}

ColourRange& current_colours() {
  static thread_local ColourRange c{0, 0};
  return c;
}

double* scratch_doubles(int slot, size_t count) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  Slot& s = g_scratch[{dev, slot + 1000 * scratch_chain()}];
  if (s.n < count) {
    if (s.p) cudaFree(s.p);
    s.p = nullptr;
    s.n = 0;
    if (cudaMalloc(&s.p, count * sizeof(double)) != cudaSuccess) return nullptr;
    s.n = count;
  }
  return s.p;
}

void scratch_release_all() {
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  for (auto& kv : g_scratch)
    if (kv.second.p) cudaFree(kv.second.p);
  g_scratch.clear();
}

cudaError_t residual_generic_0(const GridDesc&, const SchemeArgs&, bool, double*, const double*, const double*, const double*,
                               const double*, const double*, const double*, const Rect*, cudaStream_t);
cudaError_t residual_generic_1(const GridDesc&, const SchemeArgs&, bool, double*, const double*, const double*, const double*,
                               const double*, const double*, const double*, const Rect*, cudaStream_t);
cudaError_t residual_generic_5(const GridDesc&, const SchemeArgs&, bool, double*, const double*, const double*, const double*,
                               const double*, const double*, const double*, const Rect*, cudaStream_t);

cudaError_t tangent_strips_5(const GridDesc&, const SchemeArgs&, bool, const RectList&, double*, const double*, const double*, const double*,
                             const double*, const double*, const double*, cudaStream_t);
cudaError_t launch_tangent_strips5(const GridDesc& g, const SchemeArgs& a, bool wall, const RectList& rows, double* out5, const double* w,
                                   const double* wd5, const double* nx, const double* ny, const double* vol, const double* volf,
                                   cudaStream_t st) {
  return tangent_strips_5(g, a, wall, rows, out5, w, wd5, nx, ny, vol, volf, st);
}

#define BCAST_DECL_GENERIC(NAME)                                                                                              \
  cudaError_t NAME(const GridDesc&, const SchemeArgs&, bool, double*, const double*, const double*, const double*, const double*, \
                   const double*, const double*, const Rect*, cudaStream_t);
BCAST_DECL_GENERIC(residual_generic_0_o3) BCAST_DECL_GENERIC(residual_generic_1_o3) BCAST_DECL_GENERIC(residual_generic_5_o3)
BCAST_DECL_GENERIC(residual_generic_0_o7) BCAST_DECL_GENERIC(residual_generic_1_o7) BCAST_DECL_GENERIC(residual_generic_5_o7)
BCAST_DECL_GENERIC(residual_generic_0_o9) BCAST_DECL_GENERIC(residual_generic_1_o9) BCAST_DECL_GENERIC(residual_generic_5_o9)
#undef BCAST_DECL_GENERIC

// The scheme order follows the ghost depth of the block, as in the reference's drivers (BROADCAST_npz.py:501-502: gh = (order + 1) / 2
// for the dnc family): gh = 2 / 3 / 4 / 5 = flux_num_dnc3 / 5 / 7 / 9.
cudaError_t launch_residual_generic(const GridDesc& g, const SchemeArgs& a, bool wall, int ndir, double* out, const double* w,
                                    const double* wd, const double* nx, const double* ny, const double* vol, const double* volf,
                                    const Rect* rect, cudaStream_t st) {
  if (ndir != 0 && ndir != 1 && ndir != 5) return cudaErrorInvalidValue;
  const double* wdn = ndir ? wd : nullptr;
#define BCAST_CALL(F0, F1, F5) \
  return ndir == 0 ? F0(g, a, wall, out, w, nullptr, nx, ny, vol, volf, rect, st) : ndir == 1 ? F1(g, a, wall, out, w, wdn, nx, ny, vol, volf, rect, st) : F5(g, a, wall, out, w, wdn, nx, ny, vol, volf, rect, st)
  switch (g.gh) {
    case 2: BCAST_CALL(residual_generic_0_o3, residual_generic_1_o3, residual_generic_5_o3);
    case 3: BCAST_CALL(residual_generic_0, residual_generic_1, residual_generic_5);
    case 4: BCAST_CALL(residual_generic_0_o7, residual_generic_1_o7, residual_generic_5_o7);
    case 5: BCAST_CALL(residual_generic_0_o9, residual_generic_1_o9, residual_generic_5_o9);
    default: return cudaErrorInvalidValue;
  }
#undef BCAST_CALL
}

// ---------------------------------------------------------------------------------------------
// boundary fills
// ---------------------------------------------------------------------------------------------
template <int N>
__global__ void k_bc_wall(StateRW<N> s, BcLine b, double gam) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < b.lmax) bc_wall_viscous_adia_line<N>(s, b, gam, l);
}
template <int N>
__global__ void k_bc_noref(StateRW<N> s, BcLine b, const double* wbd, int lm, const double* nx, const double* ny, double gam) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < b.lmax) bc_no_reflexion_line<N>(s, b, wbd, lm, nx, ny, gam, l);
}
template <int N>
__global__ void k_bc_inlet(StateRW<N> s, BcLine b, const double* field, int lm, const double* nx, const double* ny, double gam) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < b.lmax) bc_supandsubinlet_line<N>(s, b, field, lm, nx, ny, gam, l);
}
template <int N>
__global__ void k_bc_extrap(StateRW<N> s, BcLine b) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < b.lmax) bc_extrapolate_o2_line<N>(s, b, l);
}

template <int N>
__global__ void k_bc_general(StateRW<N> s, BcLine b, const double* field, int lm) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < b.lmax) bc_general_line<N>(s, b, field, lm, l);
}

template <int N>
__global__ void k_bc_wall_iso(StateRW<N> s, BcLine b, double twall, double gam, double rgaz) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < b.lmax) bc_wall_viscous_iso_line<N>(s, b, twall, gam, rgaz, l);
}
template <int N>
__global__ void k_bc_symmetry(StateRW<N> s, BcLine b, const double* nx, const double* ny, bool anti) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= b.lmax) return;
  if (anti)
    bc_symmetry_line<N, true>(s, b, nx, ny, l);
  else
    bc_symmetry_line<N, false>(s, b, nx, ny, l);
}
template <int N>
__global__ void k_bc_pressure(StateRW<N> s, BcLine b, double pext, bool noref, double gam, const double* nx, const double* ny) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < b.lmax) bc_pressure_line<N>(s, b, pext, noref, gam, nx, ny, l);
}

template <int N>
__global__ void k_bc_wall_profile(StateRW<N> s, BcLine b, const double* prof, const double* profd, double gam, double gamd, double rgaz,
                                  double rgazd, bool blow) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= b.lmax) return;
  if (blow)
    bc_wall_profile_line<N, true>(s, b, prof, profd, gam, gamd, rgaz, rgazd, l);
  else
    bc_wall_profile_line<N, false>(s, b, prof, profd, gam, gamd, rgaz, rgazd, l);
}

#define BC_DISPATCH(KERNEL, ...)                                                     \
  do {                                                                               \
    if (b.lmax <= 0) return cudaSuccess;                                             \
    const int nb = (b.lmax + 63) / 64;                                               \
    switch (ndir) {                                                                  \
      case 0: KERNEL<0><<<nb, 64, 0, st>>>(StateRW<0>{w, wd, g}, b, ##__VA_ARGS__); break; \
      case 1: KERNEL<1><<<nb, 64, 0, st>>>(StateRW<1>{w, wd, g}, b, ##__VA_ARGS__); break; \
      case 5: KERNEL<5><<<nb, 64, 0, st>>>(StateRW<5>{w, wd, g}, b, ##__VA_ARGS__); break; \
      default: return cudaErrorInvalidValue;                                         \
    }                                                                                \
    return cudaGetLastError();                                                       \
  } while (0)

cudaError_t launch_bc_wall(const GridDesc& g, const BcLine& b, double gam, int ndir, double* w, double* wd, cudaStream_t st) {
  BC_DISPATCH(k_bc_wall, gam);
}
cudaError_t launch_bc_noref(const GridDesc& g, const BcLine& b, double gam, int ndir, double* w, double* wd, const double* wbd, int lm,
                            const double* nx, const double* ny, cudaStream_t st) {
  BC_DISPATCH(k_bc_noref, wbd, lm, nx, ny, gam);
}
cudaError_t launch_bc_inlet(const GridDesc& g, const BcLine& b, double gam, int ndir, double* w, double* wd, const double* field, int lm,
                            const double* nx, const double* ny, cudaStream_t st) {
  BC_DISPATCH(k_bc_inlet, field, lm, nx, ny, gam);
}
cudaError_t launch_bc_wall_iso(const GridDesc& g, const BcLine& b, double twall, double gam, double rgaz, int ndir, double* w, double* wd,
                               cudaStream_t st) {
  BC_DISPATCH(k_bc_wall_iso, twall, gam, rgaz);
}
cudaError_t launch_bc_symmetry(const GridDesc& g, const BcLine& b, int ndir, double* w, double* wd, const double* nx, const double* ny,
                               bool anti, cudaStream_t st) {
  BC_DISPATCH(k_bc_symmetry, nx, ny, anti);
}
cudaError_t launch_bc_pressure(const GridDesc& g, const BcLine& b, double pext, bool noref, double gam, int ndir, double* w, double* wd,
                               const double* nx, const double* ny, cudaStream_t st) {
  BC_DISPATCH(k_bc_pressure, pext, noref, gam, nx, ny);
}
cudaError_t launch_bc_wall_profile(const GridDesc& g, const BcLine& b, bool blow, const double* prof, const double* profd, double gam,
                                   double gamd, double rgaz, double rgazd, int ndir, double* w, double* wd, cudaStream_t st) {
  BC_DISPATCH(k_bc_wall_profile, prof, profd, gam, gamd, rgaz, rgazd, blow);
}
cudaError_t launch_bc_extrap(const GridDesc& g, const BcLine& b, int ndir, double* w, double* wd, cudaStream_t st) {
  BC_DISPATCH(k_bc_extrap);
}
cudaError_t launch_bc_general(const GridDesc& g, const BcLine& b, int ndir, double* w, double* wd, const double* field, int lm,
                              cudaStream_t st) {
  BC_DISPATCH(k_bc_general, field, lm);
}

// ---------------------------------------------------------------------------------------------
// jn_match window copy (borders/jn_match.F90:24-64)
// ---------------------------------------------------------------------------------------------
__global__ void k_jn_match(double* wr, Window r, int prr_i, int prr_j, const double* wd, Window d, int pd_i, int pd_j, int ni, int nj,
                           int istep, int jstep, int em) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;  // receiver offset in first dimension
  const int b = blockIdx.y * blockDim.y + threadIdx.y;
  if (a >= ni || b >= nj) return;
  // donor loop index runs i1..i2 with step istep while the receiver offset counts up
  const int i = istep == 1 ? a + 1 : ni - a;
  const int j = jstep == 1 ? b + 1 : nj - b;
  const int ir = prr_i + a, jr = prr_j + b;
  const int id = pd_i + i - 1, jd = pd_j + j - 1;
  const long long kr = (long long)(ir - r.lo_i) + (long long)(jr - r.lo_j) * r.ld;
  const long long kd = (long long)(id - d.lo_i) + (long long)(jd - d.lo_j) * d.ld;
  for (int e = 0; e < em; ++e) wr[e * r.stride + kr] = wd[e * d.stride + kd];
}

cudaError_t launch_jn_match(double* wr, const Window& r, const int prr[4], const double* wd, const Window& d, const int prd[4],
                            const int tr[2], int em, cudaStream_t st) {
  // prr/prd = {imin, jmin, imax, jmax}  (Fortran p(1,1), p(1,2), p(2,1), p(2,2))
  const int istep = tr[0] >= 0 ? 1 : -1, jstep = tr[1] >= 0 ? 1 : -1;
  const int idir = tr[0] >= 0 ? tr[0] : -tr[0], jdir = tr[1] >= 0 ? tr[1] : -tr[1];
  auto lo = [&](int dir) { return dir == 1 ? prd[0] : prd[1]; };
  auto hi = [&](int dir) { return dir == 1 ? prd[2] : prd[3]; };
  const int ni = hi(idir) - lo(idir) + 1, nj = hi(jdir) - lo(jdir) + 1;
  if (ni <= 0 || nj <= 0) return cudaSuccess;
  dim3 blk(32, 4), grd((ni + 31) / 32, (nj + 3) / 4);
  k_jn_match<<<grd, blk, 0, st>>>(wr, r, prr[0], prr[1], wd, d, lo(idir), lo(jdir), ni, nj, istep, jstep, em);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// colouring seeds (misc/ComputeJacobian.f90:357-374, 1075-1092).  ndir == 5: direction n seeds
// variable n (vector mode), m ignored.
// ---------------------------------------------------------------------------------------------
__global__ void k_testvector(GridDesc g, double* __restrict__ wd, int ndir, int m, int l, int k, int is, int ie, int js, int je,
                             RectList wins /* storage index windows that are written */) {
  const Rect win = wins.r[blockIdx.z];
  const int ii = blockIdx.x * blockDim.x + threadIdx.x + win.i0;
  const int jj = blockIdx.y * blockDim.y + threadIdx.y + win.j0;
  if (ii > win.i1 || jj > win.j1) return;
  const int i = ii + 1 - g.gh + g.ioff, j = jj + 1 - g.gh + g.joff;   // global indices (i-slabs / strip windows: seeds also fall in the halo)
  const int s = 2 * g.gh + 1;
  const bool seed = i >= is + l + 1 && i <= ie && j >= js + k + 1 && j <= je && (i - (is + l + 1)) % s == 0 && (j - (js + k + 1)) % s == 0;
  const long long kk = ii + (long long)jj * g.ldc;
  if (ndir <= 1) {
#pragma unroll
    for (int e = 0; e < 5; ++e) wd[e * g.sc + kk] = (seed && e == m) ? 1.0 : 0.0;
  } else {
    for (int n = 0; n < ndir; ++n)
#pragma unroll
      for (int e = 0; e < 5; ++e) wd[(long long)(n * 5 + e) * g.sc + kk] = (seed && e == n) ? 1.0 : 0.0;
  }
}

cudaError_t launch_testvector(const GridDesc& g, double* wd, int ndir, int m, int l, int k, const int* zone, cudaStream_t st,
                              const RectList* rows) {
  // rows (optional): only the cells a tangent restricted to these rows can read are (re)written
  RectList win = one_rect(Rect{0, g.ni() - 1, 0, g.nj() - 1});
  if (rows) {
    win = *rows;
    for (int q = 0; q < rows->n; ++q)
      win.r[q] = Rect{max(0, rows->r[q].i0 - 5 + g.gh), min(g.ni() - 1, rows->r[q].i1 + 3 + g.gh), max(0, rows->r[q].j0 - 5 + g.gh),
                      min(g.nj() - 1, rows->r[q].j1 + 3 + g.gh)};
  }
  dim3 blk(32, 4);
  int is = 0, ie = g.img, js = 0, je = g.jmg;
  if (zone) {  // testvector_partial: i = istart+l+1 .. iend+1, j = jstart+k+1 .. jend+1
    is = zone[0];
    ie = zone[1] + 1;
    js = zone[2];
    je = zone[3] + 1;
  }
  for_each_rect(win, st, [&](const RectList& w1, int, cudaStream_t s1) {
    k_testvector<<<grid_of(w1, 32, 4), blk, 0, s1>>>(g, wd, ndir, m, l, k, is, ie, js, je, w1);
  });
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// COO scatter of one colour (misc/ComputeJacobian.f90).  Integer semantics
// are exact; the segment is the contiguous slot range of colour (m,l,k).
// ---------------------------------------------------------------------------------------------
__global__ void k_scatter(GridDesc g, int kind, double* __restrict__ jac, int* __restrict__ ia, int* __restrict__ ja,
                          const double* __restrict__ resd, int m, int l, int k, const double* __restrict__ coefdiag,
                          const double* __restrict__ vol) {
  // i-slabs: il = local column (addresses resd / coefdiag / the slot), i = global column (row, column numbering and the
  // nearest-seed rules of misc/ComputeJacobian.f90, which refer to the whole block of im = g.img columns)
  const int iml = g.im, im = g.img, jm = g.jm, gh = g.gh, s = 2 * gh + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= 5LL * iml * jm) return;
  const int il = (int)(t % iml) + 1;
  const int i = il + g.ioff;
  const int j = (int)((t / iml) % jm) + 1;
  const int e = (int)(t / ((long long)iml * jm)) + 1;
  const bool withjn = kind == SCATTER_JV_RELAXED_JN || kind == SCATTER_JV_JN;
  const int dummy_ia = 5 * im * jm - (withjn ? 2 : 1);
  int row = e - 1 + (j - 1) * 5 + (i - 1) * jm * 5;
  int col = 0;
  double val = 0.0;
  bool ok = true;
  int valj, vali = 0;
  if (kind == SCATTER_DZ) {
    if (k <= gh) {
      valj = (j <= k + 1 + gh) ? k : (j - (k + 1) + gh) / s * s + k;
    } else {
      valj = (j <= k - gh) ? jm + 1 : (j - (k + 1) + gh) / s * s + k;
    }
  } else {
    valj = (j <= k + 1 + gh) ? k : (j - gh - (k + 1) + 2 * gh) / s * s + k;
  }
  if (valj >= jm) ok = false;
  if (ok) {
    if (l <= gh) {
      vali = (i <= l + 1 + gh) ? l : (i - (l + 1) + gh) / s * s + l;
    } else {
      if (i <= l - gh)
        vali = withjn ? im - 1 - 2 * gh + l : im + 1;
      else
        vali = (i - (l + 1) + gh) / s * s + l;
    }
    if (vali >= im) {
      if (withjn && vali - im <= gh - 1)
        vali = l;
      else
        ok = false;
    }
  }
  if (ok) {
    col = m + valj * 5 + vali * jm * 5;
    const double r = resd[(long long)(e - 1) * g.sc + g.cidx(il, j)];
    if (kind == SCATTER_DZ) {
      val = r;
    } else {
      val = -r;
      if ((kind == SCATTER_JV_RELAXED || kind == SCATTER_JV_RELAXED_JN || kind == SCATTER_JV_RELAXED_DBYVOL) && row == col)
        val = -r + coefdiag[(il - 1) + (long long)(j - 1) * iml];
      if (kind == SCATTER_JV_DBYVOL || kind == SCATTER_JV_RELAXED_DBYVOL) val = val / vol[g.cidx(il, j)];
    }
  } else {
    row = dummy_ia;
    col = 0;
    val = 0.0;
  }
  jac[t] = val;
  ia[t] = row;
  ja[t] = col;
}

// computejacobianfromjv_relaxed_withjnandcheck (misc/ComputeJacobian.f90:1095-1204): two zones joined in i; zone n writes
// the slot range shifted by n * 25 s^2 im jm and only where the slot is still empty (|jac| < mini) -- the slot arrays are
// read-modify-write.  Integer rules restated statement by statement (dummy row 5 im jm - 2).
__global__ void k_scatter_check(GridDesc g, double* __restrict__ jac, int* __restrict__ ia, int* __restrict__ ja,
                                const double* __restrict__ resd, int m, int l, int k, const double* __restrict__ coefdiag, double mini,
                                int zone) {
  const int im = g.im, jm = g.jm, gh = g.gh, s = 2 * gh + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= 5LL * im * jm) return;
  const int i = (int)(t % im) + 1;
  const int j = (int)((t / im) % jm) + 1;
  const int e = (int)(t / ((long long)im * jm)) + 1;
  const int dummy = 5 * im * jm - 2;
  const int row = e - 1 + (j - 1) * 5 + (i - 1) * jm * 5;
  ia[t] = row;
  const int valj = (j <= k + 1 + gh) ? k : (j - gh - (k + 1) + 2 * gh) / s * s + k;
  if (valj >= jm) {
    ia[t] = dummy;
    ja[t] = 0;
    jac[t] = 0.0;
    return;
  }
  int vali;
  if (l <= gh)
    vali = (i <= l + 1 + gh) ? l : (i - (l + 1) + gh) / s * s + l;
  else
    vali = (i <= l - gh) ? l : (i - (l + 1) + gh) / s * s + l;
  bool clear_if_taken;   // what happens when the slot already holds an entry
  if (zone == 0) {
    if (vali >= im - 2 * gh) {
      vali = l;
      clear_if_taken = true;
    } else {
      clear_if_taken = false;
    }
  } else {
    if (vali <= 2 * gh) {
      vali = im - im % s + l;
      if (vali > im - 1) vali -= s;
      clear_if_taken = false;
    } else {
      clear_if_taken = true;
    }
  }
  if (::fabs(jac[t]) < mini) {
    const int col = m + valj * 5 + vali * jm * 5;
    const double r = resd[(long long)(e - 1) * g.sc + g.cidx(i, j)];
    ja[t] = col;
    jac[t] = (row == col) ? coefdiag[(i - 1) + (long long)(j - 1) * im] - r : -r;
  } else if (clear_if_taken) {
    ia[t] = dummy;
    ja[t] = 0;
    jac[t] = 0.0;
  }
}

cudaError_t launch_scatter_check(const GridDesc& g, double* seg_jac, int* seg_ia, int* seg_ja, const double* resd, int m, int l, int k,
                                 const double* coefdiag, double mini, int zone, cudaStream_t st) {
  const long long n = 5LL * g.im * g.jm;
  k_scatter_check<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, seg_jac, seg_ia, seg_ja, resd, m, l, k, coefdiag, mini, zone);
  return cudaGetLastError();
}

cudaError_t launch_scatter(const GridDesc& g, int kind, double* seg_jac, int* seg_ia, int* seg_ja, const double* resd, int m, int l,
                           int k, const double* coefdiag, const double* vol, cudaStream_t st) {
  const long long n = 5LL * g.im * g.jm;
  k_scatter<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, kind, seg_jac, seg_ia, seg_ja, resd, m, l, k, coefdiag, vol);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// norms: per-equation sum r^2 and sum r^10 over interior cells (norm.F90:34-77).  Reduction order:
// per-thread strided partial sums -> warp shuffle tree -> one atomicAdd per warp (order not fixed).
// ---------------------------------------------------------------------------------------------
__global__ void k_norms(GridDesc g, const double* __restrict__ res, double* __restrict__ out10) {
  const int e = blockIdx.y;
  const long long n = (long long)g.im * g.jm;
  double s2 = 0.0, s10 = 0.0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % g.im) + 1, j = (int)(t / g.im) + 1;
    const double r = res[e * g.sc + g.cidx(i, j)];
    const double r2 = r * r;
    s2 += r2;
    const double r4 = r2 * r2;
    s10 += r4 * r4 * r2;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
    s10 += __shfl_down_sync(0xffffffffu, s10, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out10 + e, s2);
    atomicAdd(out10 + 5 + e, s10);
  }
}

cudaError_t launch_norms(const GridDesc& g, const double* res, double* out10, cudaStream_t st) {
  cudaMemsetAsync(out10, 0, 10 * sizeof(double), st);
  k_norms<<<dim3(148, 5), 256, 0, st>>>(g, res, out10);
  return cudaGetLastError();
}

}  // namespace bcast
