// Device-resident colour loop: the reference algorithm (seed -> linearised boundary fills -> tangent
// -> scatter; BROADCAST_npz.py:1068-1127, misc/ComputeJacobian.f90) with the five variables of a
// colour (l,k) carried together as five tangent directions, so 49 passes instead of 245.
#include <cstdlib>
#include <cstring>
#include <list>
#include <string>
#include <vector>
#include "../../include/broadcast_b200.h"
#include "kernels.cuh"

namespace bcast {
void count_launches(int n);
cudaError_t launch_dz(const GridDesc& g, const SchemeArgs& a, int which, int ndir, double* out, const double* w, const double* wd,
                      const double* nx, const double* ny, const double* vol, const double* volf, const Rect* rect, cudaStream_t st);

static cudaError_t apply_bc_list(const GridDesc& g, double gam, int ndir, double* w, double* wd, const double* nx, const double* ny,
                                 const bc_desc_t* bcs, int nbcs, cudaStream_t st) {
  for (int b = 0; b < nbcs; ++b) {
    const bc_desc_t& d = bcs[b];
    cudaError_t e = cudaSuccess;
    if (d.kind == BC_KIND_JOIN) {
      Window win{g.ldc, g.sc, 1 - g.gh, 1 - g.gh};
      if (ndir > 0) e = launch_jn_match(wd, win, d.window, wd, win, d.prd, d.tr, 5 * ndir, st);
      if (e == cudaSuccess) e = launch_jn_match(w, win, d.window, w, win, d.prd, d.tr, 5, st);
      count_launches(ndir > 0 ? 2 : 1);
    } else {
      BcLine line;
      if (!decode_interface(d.loc, d.window, line)) return cudaErrorInvalidValue;
      switch (d.kind) {
        case BC_KIND_INLET: e = launch_bc_inlet(g, line, gam, ndir, w, wd, d.table, d.lm, nx, ny, st); break;
        case BC_KIND_NOREF: e = launch_bc_noref(g, line, gam, ndir, w, wd, d.table, d.lm, nx, ny, st); break;
        case BC_KIND_EXTRAP: e = launch_bc_extrap(g, line, ndir, w, wd, st); break;
        case BC_KIND_WALL: e = launch_bc_wall(g, line, gam, ndir, w, wd, st); break;
        case BC_KIND_WALL_ISO: e = launch_bc_wall_iso(g, line, d.param[0], gam, d.param[1], ndir, w, wd, st); break;
        case BC_KIND_SYMMETRY: e = launch_bc_symmetry(g, line, ndir, w, wd, nx, ny, false, st); break;
        case BC_KIND_ANTISYMMETRY: e = launch_bc_symmetry(g, line, ndir, w, wd, nx, ny, true, st); break;
        case BC_KIND_PRESSURE: e = launch_bc_pressure(g, line, d.param[0], d.param[1] != 0.0, gam, ndir, w, wd, nx, ny, st); break;
        case BC_KIND_WALL_BLOW_PROFILE:
        case BC_KIND_WALL_ISO_PROFILE:
          if (!d.table || d.lm < line.lmax) return cudaErrorInvalidValue;
          e = launch_bc_wall_profile(g, line, d.kind == BC_KIND_WALL_BLOW_PROFILE, d.table, nullptr, gam, 0.0, d.param[1], 0.0, ndir, w, wd, st);
          break;
        default: return cudaErrorInvalidValue;
      }
      count_launches(1);
    }
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// Linearised boundary fills a colour (l,k) can skip: the tangent a fill writes into the ghosts of a boundary line depends on wd
// of the first three interior cells of the SAME line only (bc.cuh: wall mirror / pressure extrapolation, o2 extrapolation,
// characteristic updates from the first interior cell and the ghosts already written; the isothermal wall, the (anti)symmetry mirror
// of layers 0 .. gh-1 and the pressure outlet read the same cells), and the seeds of a colour sit on the rows
// j = k+1 (mod 7), columns i = l+1 (mod 7): a side whose three first interior rows (columns) hold no seed row (column) of this
// colour receives zero tangents, which is what the seeding kernel has already written there.  Joins are always applied.
static int active_bcs(const GridDesc& g, const bc_desc_t* bcs, int nbcs, int l, int k, bc_desc_t* out) {
  const int s = 2 * g.gh + 1;
  auto near_lo = [&](int c) { return c <= g.gh - 1; };                                   // a seed line among lines 1 .. gh
  auto near_hi = [&](int c, int n) { return n >= c + 1 && (n - (c + 1)) % s <= g.gh - 1; };   // ... among n-gh+1 .. n
  int m = 0;
  for (int b = 0; b < nbcs; ++b) {
    const bc_desc_t& d = bcs[b];
    bool keep = true;
    if (d.kind != BC_KIND_JOIN) {
      const bool lo = d.loc[1] == 'l' || d.loc[1] == 'L';
      if (d.loc[0] == 'I' || d.loc[0] == 'i') keep = lo ? near_lo(l) : near_hi(l, g.img);
      else keep = lo ? near_lo(k) : near_hi(k, g.jmg);
    }
    if (keep) out[m++] = d;
  }
  return m;
}

// scatter of the five directions of colour (l,k): same integer rules as k_scatter, writing straight
// into the reference's slot order for m = 0..4
__global__ void k_scatter5(GridDesc g, int kind, double* __restrict__ jac, int* __restrict__ ia, int* __restrict__ ja,
                           const double* __restrict__ resd5, int l, int k, const double* __restrict__ coefdiag,
                           const double* __restrict__ vol, Rect rc, int compact) {
  // i-slabs / strip windows: il, jl = local indices (addresses), i, j = global indices (numbering, nearest-seed rules; see k_scatter)
  const int iml = g.im, jml = g.jm, im = g.img, jm = g.jmg, gh = g.gh, s = 2 * gh + 1;
  const int wi = rc.i1 - rc.i0 + 1, wj = rc.j1 - rc.j0 + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= 5LL * wi * wj) return;
  const int il = (int)(t % wi) + rc.i0;
  const int i = il + g.ioff;
  const int jl = (int)((t / wi) % wj) + rc.j0;
  const int j = jl + g.joff;
  const int e = (int)(t / ((long long)wi * wj)) + 1;
  const bool withjn = kind == SCATTER_JV_RELAXED_JN || kind == SCATTER_JV_JN;
  const int dummy_ia = 5 * im * jm - (withjn ? 2 : 1);
  const int row = e - 1 + (j - 1) * 5 + (i - 1) * jm * 5;
  bool ok = true;
  int valj, vali = 0;
  if (kind == SCATTER_DZ) {
    if (k <= gh) valj = (j <= k + 1 + gh) ? k : (j - (k + 1) + gh) / s * s + k;
    else valj = (j <= k - gh) ? jm + 1 : (j - (k + 1) + gh) / s * s + k;
  } else {
    valj = (j <= k + 1 + gh) ? k : (j - gh - (k + 1) + 2 * gh) / s * s + k;
  }
  if (valj >= jm) ok = false;
  if (ok) {
    if (l <= gh) {
      vali = (i <= l + 1 + gh) ? l : (i - (l + 1) + gh) / s * s + l;
    } else {
      if (i <= l - gh) vali = withjn ? im - 1 - 2 * gh + l : im + 1;
      else vali = (i - (l + 1) + gh) / s * s + l;
    }
    if (vali >= im) {
      if (withjn && vali - im <= gh - 1) vali = l;
      else ok = false;
    }
  }
  // reference slot order (ComputeJacobian.f90:524), over the whole grid or (compact) over the rectangle only
  const long long n = compact ? 5LL * wi * wj : 5LL * iml * jml;
  const long long cell = compact ? (long long)(il - rc.i0) + (long long)(jl - rc.j0) * wi + (long long)(e - 1) * wi * wj
                                 : (long long)(il - 1) + (long long)(jl - 1) * iml + (long long)(e - 1) * iml * jml;
  const long long kc = g.cidx(il, jl);
#pragma unroll
  for (int m = 0; m < 5; ++m) {
    const long long slot = cell + (long long)k * n + (long long)l * n * s + (long long)m * n * s * s;
    if (ok) {
      const int col = m + valj * 5 + vali * jm * 5;
      const double r = resd5[(long long)(m * 5 + (e - 1)) * g.sc + kc];
      double val;
      if (kind == SCATTER_DZ) {
        val = r;
      } else {
        val = -r;
        if ((kind == SCATTER_JV_RELAXED || kind == SCATTER_JV_RELAXED_JN || kind == SCATTER_JV_RELAXED_DBYVOL) && row == col)
          val = -r + coefdiag[(il - 1) + (long long)(jl - 1) * iml];
        if (kind == SCATTER_JV_DBYVOL || kind == SCATTER_JV_RELAXED_DBYVOL) val = val / vol[kc];
      }
      jac[slot] = val;
      ia[slot] = row;
      ja[slot] = col;
    } else {
      jac[slot] = 0.0;
      ia[slot] = dummy_ia;
      ja[slot] = 0;
    }
  }
}

}  // namespace bcast

using namespace bcast;

extern "C" int bcd_apply_bcs(double* w, const double* nx, const double* ny, double gam, int gh, int im, int jm, const bc_desc_t* bcs,
                             int nbcs, void* stream) {
  if (im < 1 || jm < 1 || gh < 2 || gh > 5) return BC_ERR_ARG;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  cudaError_t e = apply_bc_list(g, gam, 0, w, nullptr, nx, ny, bcs, nbcs, (cudaStream_t)stream);
  return e == cudaSuccess ? BC_OK : (int)e;
}

extern "C" int bcd_jacobian_coo(double* jac, int32_t* ia, int32_t* ja, double* w, const double* nx, const double* ny, const double* vol,
                                const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                                double muref, double tref, double s_suth, double k2, double k4, int im, int jm, int wall,
                                const bc_desc_t* bcs, int nbcs, int scatter_kind, const double* coefdiag, const int32_t* rect,
                                int compact, void* stream) {
  // gh = 2 / 4 / 5: the colour loop of the orders 3 / 7 / 9 ((2 gh + 1)^2 passes of five directions), tangent by the order-templated
  // pipeline of generic_impl.cuh; the fused tile tangent and the face skipping below are order 5
  if (im < 1 || jm < 1 || gh < 2 || gh > 5) return BC_ERR_ARG;
  if (scatter_kind < 0 || scatter_kind > 6) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const SchemeArgs a{cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4};
  const int s = 2 * gh + 1;
  double* wd5 = scratch_doubles(20, (size_t)g.sc * 25);
  double* resd5 = scratch_doubles(21, (size_t)g.sc * 25);
  if (!wd5 || !resd5) return BC_ERR_ALLOC;
  Rect rc{1, im, 1, jm};
  if (rect) rc = Rect{rect[0], rect[1], rect[2], rect[3]};
  if (rc.i1 < rc.i0 || rc.j1 < rc.j0) return BC_OK;
  const long long nt = 5LL * (rc.i1 - rc.i0 + 1) * (rc.j1 - rc.j0 + 1);
  // the seeds may be restricted to the window the rows of `rect` can read, unless a join copies tangents from the far
  // side of the block (periodic cut): then every seed matters
  bool has_join = false;
  for (int b = 0; b < nbcs; ++b) has_join = has_join || bcs[b].kind == BC_KIND_JOIN;
  const RectList seed_list = one_rect(rc);
  const RectList* seed_rows = (rect && !has_join) ? &seed_list : nullptr;
  // every ghost of w holds the value of this list before the passes start (they refresh only the sides their seeds reach)
  {
    cudaError_t e0 = apply_bc_list(g, gam, 0, w, nullptr, nx, ny, bcs, nbcs, st);
    if (e0 != cudaSuccess) return (int)e0;
    count_launches(nbcs);
  }
  for (int l = 0; l < s; ++l)
    for (int k = 0; k < s; ++k) {
      if (!current_colours().has(l * s + k)) continue;   // colour sharding: this rank's passes only
      cudaError_t e = launch_testvector(g, wd5, 5, 0, l, k, nullptr, st, seed_rows);
      if (e != cudaSuccess) return (int)e;
      static const bool skip_bcs = getenv("BROADCAST_B200_NO_BC_SKIP") == nullptr;
      bc_desc_t act[16];
      const int nact = (skip_bcs && nbcs <= 16) ? active_bcs(g, bcs, nbcs, l, k, act) : -1;
      e = nact >= 0 ? apply_bc_list(g, gam, 5, w, wd5, nx, ny, act, nact, st) : apply_bc_list(g, gam, 5, w, wd5, nx, ny, bcs, nbcs, st);
      if (e != cudaSuccess) return (int)e;
      // tangent of the rows of rc, five directions: one face per thread, every face once, faces without a tangent input skipped
      // (k_strip_faces5 over bands of three rows); BROADCAST_B200_COO_GENERIC=1: the cell-centred k_balance<5>
      static const bool coo_generic = getenv("BROADCAST_B200_COO_GENERIC") != nullptr;
      if (coo_generic || gh != 3) e = launch_residual_generic(g, a, wall != 0, 5, resd5, w, wd5, nx, ny, vol, volf, &rc, st);
      else e = launch_tangent_strips5(g, a, wall != 0, one_rect(rc), resd5, w, wd5, nx, ny, vol, volf, st);
      if (e != cudaSuccess) return (int)e;
      k_scatter5<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(g, scatter_kind, jac, ia, ja, resd5, l, k, coefdiag, vol, rc, (rect && compact) ? 1 : 0);
      e = cudaGetLastError();
      if (e != cudaSuccess) return (int)e;
      count_launches(7);
    }
  return BC_OK;
}

// Colour loop of the spanwise operators (BROADCAST_npz.py:1231-1246): seeds -> linearised boundary fills ->
// coeffs_5p_dz / coeffs_5p_dz2 -> computejacobianfromdz, five variables per pass; both COO triplets at once.
extern "C" int bcd_dz_coo(double* jac1, int32_t* ia1, int32_t* ja1, double* jac2, int32_t* ia2, int32_t* ja2, double* w, const double* nx,
                          const double* ny, const double* vol, const double* volf, int gh, double cp, double cv, double prandtl, double gam,
                          double rgaz, double cs, double muref, double tref, double s_suth, int im, int jm, const bc_desc_t* bcs, int nbcs,
                          void* stream) {
  if (im < 1 || jm < 1 || gh != 3) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const SchemeArgs a{cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, 0.0, 0.0};
  const int s = 2 * gh + 1;
  double* wd5 = scratch_doubles(20, (size_t)g.sc * 25);
  double* out5 = scratch_doubles(21, (size_t)g.sc * 25);
  if (!wd5 || !out5) return BC_ERR_ALLOC;
  const Rect rc{1, im, 1, jm};
  const long long nt = 5LL * im * jm;
  // every ghost of w holds the value of this list before the passes start (they refresh only the sides their seeds reach)
  {
    cudaError_t e0 = apply_bc_list(g, gam, 0, w, nullptr, nx, ny, bcs, nbcs, st);
    if (e0 != cudaSuccess) return (int)e0;
    count_launches(nbcs);
  }
  for (int l = 0; l < s; ++l)
    for (int k = 0; k < s; ++k) {
      if (!current_colours().has(l * s + k)) continue;   // colour sharding: this rank's passes only
      cudaError_t e = launch_testvector(g, wd5, 5, 0, l, k, nullptr, st);
      if (e != cudaSuccess) return (int)e;
      static const bool skip_bcs = getenv("BROADCAST_B200_NO_BC_SKIP") == nullptr;
      bc_desc_t act[16];
      const int nact = (skip_bcs && nbcs <= 16) ? active_bcs(g, bcs, nbcs, l, k, act) : -1;
      e = nact >= 0 ? apply_bc_list(g, gam, 5, w, wd5, nx, ny, act, nact, st) : apply_bc_list(g, gam, 5, w, wd5, nx, ny, bcs, nbcs, st);
      if (e != cudaSuccess) return (int)e;
      for (int which = 1; which <= 2; ++which) {
        double* jac = which == 1 ? jac1 : jac2;
        int32_t* ia = which == 1 ? ia1 : ia2;
        int32_t* ja = which == 1 ? ja1 : ja2;
        if (!jac) continue;
        e = launch_dz(g, a, which, 5, out5, w, wd5, nx, ny, vol, volf, &rc, st);
        if (e != cudaSuccess) return (int)e;
        k_scatter5<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(g, SCATTER_DZ, jac, ia, ja, out5, l, k, nullptr, vol, rc, 0);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        count_launches(1);
      }
      count_launches(1);
    }
  return BC_OK;
}

// Colour loop of the sensitivity driver (BROADCAST_npz_sens.py:1741-1800) on the device: derivative of the Dz / Dz2 operator rows applied
// to a mode (real part wmoder, imaginary part wmodei or null) with respect to the base flow, column by column.  Per colour (l,k): the
// five seeds at once, the linearised boundary fills of the list, then for every seed m the fused pass k_dz_tangent (both operators,
// wd0 = seed m, wd = mode) and the scatter rule computejacobianfromdz (misc/ComputeJacobian.f90:708-779).  ia / ja are shared by the
// real and imaginary value lists, as in the driver (IAdz, JAdz).  Any of the four value lists may be null.
namespace bcast {
cudaError_t launch_dz_tangent_raw(const GridDesc& g, double cp, double cv, double prandtl, double gam, double cs, double muref, double tref,
                                  double s_suth, double* out1, double* out2, const double* w, const double* wd0, const double* wd,
                                  const double* nx, const double* ny, const double* vol, const Rect& rc, cudaStream_t st);
}

extern "C" int bcd_dz_tangent_coo(double* jac1r, double* jac1i, int32_t* ia1, int32_t* ja1, double* jac2r, double* jac2i, int32_t* ia2,
                                  int32_t* ja2, double* w, const double* wmoder, const double* wmodei, const double* nx, const double* ny,
                                  const double* vol, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                                  double muref, double tref, double s_suth, int im, int jm, const bc_desc_t* bcs, int nbcs, void* stream) {
  (void)rgaz;
  if (im < 1 || jm < 1 || gh != 3 || !wmoder) return BC_ERR_ARG;
  if (((jac1r || jac1i) && (!ia1 || !ja1)) || ((jac2r || jac2i) && (!ia2 || !ja2))) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const int s = 2 * gh + 1;
  double* wd5 = scratch_doubles(20, (size_t)g.sc * 25);
  double* out5a = scratch_doubles(21, (size_t)g.sc * 25);
  double* out5b = scratch_doubles(22, (size_t)g.sc * 25);
  if (!wd5 || !out5a || !out5b) return BC_ERR_ALLOC;
  const Rect rc{1, im, 1, jm};
  const long long nt = 5LL * im * jm;
  const bool want1 = jac1r || jac1i, want2 = jac2r || jac2i;
  // every ghost of w holds the value of this list before the passes start (they refresh only the sides their seeds reach)
  {
    cudaError_t e0 = apply_bc_list(g, gam, 0, w, nullptr, nx, ny, bcs, nbcs, st);
    if (e0 != cudaSuccess) return (int)e0;
    count_launches(nbcs);
  }
  for (int l = 0; l < s; ++l)
    for (int k = 0; k < s; ++k) {
      if (!current_colours().has(l * s + k)) continue;
      cudaError_t e = launch_testvector(g, wd5, 5, 0, l, k, nullptr, st);
      if (e != cudaSuccess) return (int)e;
      bc_desc_t act[16];
      const int nact = nbcs <= 16 ? active_bcs(g, bcs, nbcs, l, k, act) : -1;
      e = nact >= 0 ? apply_bc_list(g, gam, 5, w, wd5, nx, ny, act, nact, st) : apply_bc_list(g, gam, 5, w, wd5, nx, ny, bcs, nbcs, st);
      if (e != cudaSuccess) return (int)e;
      count_launches(1);
      for (int part = 0; part < 2; ++part) {
        const double* mode = part == 0 ? wmoder : wmodei;
        double* j1 = part == 0 ? jac1r : jac1i;
        double* j2 = part == 0 ? jac2r : jac2i;
        if (!mode || (!j1 && !j2)) continue;
        for (int m = 0; m < 5; ++m) {
          e = launch_dz_tangent_raw(g, cp, cv, prandtl, gam, cs, muref, tref, s_suth, j1 ? out5a + (size_t)m * 5 * g.sc : nullptr,
                                    j2 ? out5b + (size_t)m * 5 * g.sc : nullptr, w, wd5 + (size_t)m * 5 * g.sc, mode, nx, ny, vol, rc, st);
          if (e != cudaSuccess) return (int)e;
        }
        if (j1) k_scatter5<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(g, SCATTER_DZ, j1, ia1, ja1, out5a, l, k, nullptr, vol, rc, 0);
        if (j2) k_scatter5<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(g, SCATTER_DZ, j2, ia2, ja2, out5b, l, k, nullptr, vol, rc, 0);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        count_launches((j1 ? 1 : 0) + (j2 ? 1 : 0));
      }
    }
  (void)want1; (void)want2;
  return BC_OK;
}

constexpr int MAXCHAIN = 16;

// The boundary strips of the Jacobian (rows within gh of a physical side) by the reference colour loop, all strips in
// the SAME 49 passes: seeds (windows of the strips only) -> linearised boundary fills -> tangent of the strip rows with
// one thread per (cell, face) -> scatter into one compact COO triple per strip (slot order of
// misc/ComputeJacobian.f90:524 over the strip's own index space, ia/ja global).
static int strips_impl(int nrect, const int32_t* rects /* [nrect][4] i0,i1,j0,j1 */, double* const* jac, int32_t* const* ia,
                       int32_t* const* ja, double* w, const double* nx, const double* ny, const double* vol, const double* volf,
                       int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref,
                       double tref, double s_suth, double k2, double k4, int im, int jm, int wall, const bc_desc_t* bcs,
                       int nbcs, int scatter_kind, const double* coefdiag, void* stream, int kcap, int chain0 = 0) {
  // chain0: first scratch chain of this call (the sub-block loops of the four strips run CONCURRENTLY on streams of their own:
  // each needs its own set of chain buffers)
  if (im < 1 || jm < 1 || gh != 3 || nrect < 0 || nrect > 4) return BC_ERR_ARG;
  if (scatter_kind < 0 || scatter_kind > 6) return BC_ERR_ARG;
  if (nrect == 0) return BC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const SchemeArgs a{cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4};
  const int s = 2 * gh + 1;
  // colour chains: up to four colours of the loop run at a time inside the graph (each chain has its own wd / tangent / primitive
  // scratch and side streams): on the reference's own grid sizes a pass is a chain of ~10 dependent small kernels, latency bound
  static const int chains_env = getenv("BROADCAST_B200_STRIP_CHAINS") ? atoi(getenv("BROADCAST_B200_STRIP_CHAINS")) : 0;
  static const bool no_graph = getenv("BROADCAST_B200_NO_GRAPH") != nullptr;
  // each chain holds 116 full-grid planes of scratch (0.93 KB per cell): four chains up to 4.5 M cells, two up to 9 M, one beyond
  const long long ncells_ = (long long)im * jm;
  int K = chains_env > 0 ? chains_env : (ncells_ <= 4500000LL ? 4 : (ncells_ <= 9000000LL ? 2 : 1));
  if (kcap > 4 && ncells_ <= 400000LL) K = kcap;   // a strip sub-block: tiny kernels, every chain costs ~100 MB of scratch
  if (K > MAXCHAIN) K = MAXCHAIN;
  if (no_graph) K = 1;
  double* wd5c[MAXCHAIN] = {};
  double* resd5c[MAXCHAIN] = {};
  double* sc_[MAXCHAIN][4];
  for (int c = 0; c < K; ++c) {   // every scratch slot the loop touches must exist before capture (allocation is not capturable)
    scratch_chain() = chain0 + c;
    wd5c[c] = scratch_doubles(20, (size_t)g.sc * 25);
    resd5c[c] = scratch_doubles(21, (size_t)g.sc * 25);
    sc_[c][0] = scratch_doubles(0, (size_t)g.sc * NPRIM);
    sc_[c][1] = scratch_doubles(1, (size_t)g.sc * NGRAD);
    sc_[c][2] = scratch_doubles(2, (size_t)g.sc * NPRIM * 5);
    sc_[c][3] = scratch_doubles(3, (size_t)g.sc * NGRAD * 5);
    scratch_chain() = 0;
    if (!wd5c[c] || !resd5c[c] || !sc_[c][0] || !sc_[c][1] || !sc_[c][2] || !sc_[c][3]) return BC_ERR_ALLOC;
  }
  RectList rows;
  rows.n = nrect;
  for (int q = 0; q < 4; ++q) rows.r[q] = q < nrect ? Rect{rects[4 * q], rects[4 * q + 1], rects[4 * q + 2], rects[4 * q + 3]} : Rect{1, 0, 1, 0};
  for (int q = 0; q < nrect; ++q)
    if (rows.r[q].i1 < rows.r[q].i0 || rows.r[q].j1 < rows.r[q].j0) return BC_ERR_ARG;
  bool has_join = false;
  for (int b = 0; b < nbcs; ++b) has_join = has_join || bcs[b].kind == BC_KIND_JOIN;
  int nlaunch = 0;
  static const bool skip_bcs = getenv("BROADCAST_B200_NO_BC_SKIP") == nullptr;
  // one pass = one colour (l,k) on stream s_, with the buffers of chain c
  auto pass = [&](int l, int k, int c, cudaStream_t s_) -> cudaError_t {
    double* wd5 = wd5c[c];
    double* resd5 = resd5c[c];
    scratch_chain() = chain0 + c;
    cudaError_t e = launch_testvector(g, wd5, 5, 0, l, k, nullptr, s_, has_join ? nullptr : &rows);
    if (e == cudaSuccess) {
      bc_desc_t act[16];
      const int nact = (skip_bcs && nbcs <= 16) ? active_bcs(g, bcs, nbcs, l, k, act) : -1;
      e = nact >= 0 ? apply_bc_list(g, gam, 5, w, wd5, nx, ny, act, nact, s_) : apply_bc_list(g, gam, 5, w, wd5, nx, ny, bcs, nbcs, s_);
    }
    if (e == cudaSuccess) e = launch_tangent_strips5(g, a, wall != 0, rows, resd5, w, wd5, nx, ny, vol, volf, s_);
    if (e == cudaSuccess) {
      for_each_rect(rows, s_, [&](const RectList& r1, int q, cudaStream_t s1) {
        const Rect rc = r1.r[0];
        const long long nt = 5LL * (rc.i1 - rc.i0 + 1) * (rc.j1 - rc.j0 + 1);
        k_scatter5<<<(unsigned)((nt + 255) / 256), 256, 0, s1>>>(g, scatter_kind, jac[q], ia[q], ja[q], resd5, l, k, coefdiag, vol, rc, 1);
      });
      e = cudaGetLastError();
    }
    scratch_chain() = 0;
    nlaunch += 6 + nrect;
    return e;
  };
  // the 49 passes on one stream (K = 1) or dealt round-robin to K chain streams forked from s_ and joined at the end
  auto run = [&](cudaStream_t s_, cudaStream_t* cs, cudaEvent_t fork, cudaEvent_t* join) -> cudaError_t {
    // The passes below apply the combined primal + tangent fills only on the sides their seeds can reach (active_bcs), and with
    // K > 1 several colours write w's ghosts while other chains read them: both are correct only if every ghost of w already
    // holds the value of the SAME list (the reference loop refreshes all of them in every colour, BROADCAST_npz.py:1079-1082).
    // Establish that here, once, before the fork -- the caller does not have to.
    {
      cudaError_t e0 = apply_bc_list(g, gam, 0, w, nullptr, nx, ny, bcs, nbcs, s_);
      if (e0 != cudaSuccess) return e0;
      nlaunch += nbcs;
    }
    if (K > 1) {
      cudaEventRecord(fork, s_);
      for (int c = 0; c < K; ++c) cudaStreamWaitEvent(cs[c], fork, 0);
    }
    int n = 0;
    cudaError_t err = cudaSuccess;
    for (int l = 0; l < s && err == cudaSuccess; ++l)
      for (int k = 0; k < s && err == cudaSuccess; ++k) {
        if (!current_colours().has(l * s + k)) continue;   // colour sharding: this rank's passes only
        const int c = K > 1 ? n % K : 0;
        err = pass(l, k, c, K > 1 ? cs[c] : s_);
        ++n;
      }
    if (K > 1)
      for (int c = 0; c < K; ++c) {
        cudaEventRecord(join[c], cs[c]);
        cudaStreamWaitEvent(s_, join[c], 0);
      }
    return err;
  };
  // CUDA graph of the whole loop, cached on everything a launch depends on (pointers, sizes, scalars, slab and colour context,
  // scratch arena addresses).  BROADCAST_B200_NO_GRAPH=1 launches kernel by kernel.
  if (no_graph) {
    cudaError_t e = run(st, nullptr, nullptr, nullptr);
    count_launches(nlaunch);
    return e == cudaSuccess ? BC_OK : (int)e;
  }
  static thread_local cudaStream_t chain_st[MAXCHAIN] = {};
  static thread_local cudaEvent_t chain_join[MAXCHAIN] = {};
  static thread_local cudaEvent_t chain_fork = nullptr;
  if (K > 1 && !chain_fork) {
    for (int c = 0; c < MAXCHAIN; ++c) {
      if (cudaStreamCreateWithFlags(&chain_st[c], cudaStreamNonBlocking) != cudaSuccess) return BC_ERR_ALLOC;
      if (cudaEventCreateWithFlags(&chain_join[c], cudaEventDisableTiming) != cudaSuccess) return BC_ERR_ALLOC;
    }
    if (cudaEventCreateWithFlags(&chain_fork, cudaEventDisableTiming) != cudaSuccess) return BC_ERR_ALLOC;
  }
  std::string key;
  auto put = [&](const void* p, size_t n) { key.append(reinterpret_cast<const char*>(p), n); };
  int dev = 0;
  cudaGetDevice(&dev);
  // (field by field: raw struct bytes would drag uninitialised padding into the key and miss the cache at random)
  const long long gk[] = {g.im, g.jm, g.gh, g.ldc, g.ldn, g.sc, g.sn, g.ioff, g.img, g.edges, g.joff, g.jmg};
  put(&dev, sizeof dev); put(gk, sizeof gk); put(&a, sizeof a); put(&wall, sizeof wall);
  {
    const WallIso wi = current_wall_iso();   // baked into the captured kernel arguments (SchemeConsts)
    put(&wi.on, sizeof wi.on); put(&wi.twall, sizeof wi.twall);
  }
  for (int q = 0; q < 4; ++q) {
    const int rk[] = {rows.n, rows.r[q].i0, rows.r[q].i1, rows.r[q].j0, rows.r[q].j1};
    put(rk, sizeof rk);
  }
  put(&scatter_kind, sizeof scatter_kind); put(&nbcs, sizeof nbcs); put(&gam, sizeof gam);
  for (int b = 0; b < nbcs; ++b) {
    const bc_desc_t& d = bcs[b];
    put(&d.kind, sizeof d.kind); put(d.loc, 3); put(d.window, sizeof d.window); put(d.prd, sizeof d.prd); put(d.tr, sizeof d.tr);
    put(&d.lm, sizeof d.lm); put(&d.table, sizeof d.table); put(d.param, sizeof d.param);   // param: baked into the captured arguments
  }
  for (int q = 0; q < nrect; ++q) { put(&jac[q], sizeof(void*)); put(&ia[q], sizeof(void*)); put(&ja[q], sizeof(void*)); }
  const void* ptrs[] = {w, nx, ny, vol, volf, coefdiag};
  put(ptrs, sizeof ptrs);
  put(&K, sizeof K);
  for (int c = 0; c < K; ++c) {
    const void* pc[] = {wd5c[c], resd5c[c], sc_[c][0], sc_[c][1], sc_[c][2], sc_[c][3]};
    put(pc, sizeof pc);
  }
  const int crk[] = {current_colours().c0, current_colours().c1};
  put(crk, sizeof crk);
  struct Entry { std::string key; cudaGraphExec_t exec; int nlaunch; };
  static thread_local std::list<Entry> cache;
  // graphs cannot be captured on the legacy default stream: use a blocking side stream, which the legacy stream orders with
  static thread_local cudaStream_t side = nullptr;
  cudaStream_t gs = st;
  if (st == nullptr || st == cudaStreamLegacy) {
    if (!side && cudaStreamCreate(&side) != cudaSuccess) return BC_ERR_ALLOC;
    gs = side;
  }
  for (auto it = cache.begin(); it != cache.end(); ++it)
    if (it->key == key) {
      cudaError_t e = cudaGraphLaunch(it->exec, gs);
      count_launches(it->nlaunch);
      cache.splice(cache.begin(), cache, it);
      return e == cudaSuccess ? BC_OK : (int)e;
    }
  static const bool no_fork = getenv("BROADCAST_B200_NO_GRAPH_FORK") != nullptr;
  for (int c = 0; c < K; ++c) {   // side streams and events exist before the capture starts
    scratch_chain() = chain0 + c;
    rect_fork().ready();
  }
  scratch_chain() = 0;
  cudaError_t e = cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) return (int)e;
  for (int c = 0; c < K; ++c) {   // per-rectangle launches become parallel branches of the graph
    scratch_chain() = chain0 + c;
    rect_fork().on = !no_fork;
  }
  scratch_chain() = 0;
  const cudaError_t er = run(gs, chain_st, chain_fork, chain_join);
  for (int c = 0; c < K; ++c) {
    scratch_chain() = chain0 + c;
    rect_fork().on = false;
  }
  scratch_chain() = 0;
  cudaGraph_t graph = nullptr;
  e = cudaStreamEndCapture(gs, &graph);
  if (er != cudaSuccess || e != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    return (int)(er != cudaSuccess ? er : e);
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return (int)e;
  cache.push_front(Entry{key, exec, nlaunch});
  if (cache.size() > 16) {
    cudaGraphExecDestroy(cache.back().exec);
    cache.pop_back();
  }
  e = cudaGraphLaunch(exec, gs);
  count_launches(nlaunch);
  return e == cudaSuccess ? BC_OK : (int)e;
}

// ---- boundary strips on SUB-BLOCKS --------------------------------------------------------------------------------------------
// The colour loop above works on arrays indexed over the whole block: every colour chain needs 116 full-grid planes of scratch, so
// C5 on one GPU could afford one chain and its 49 passes ran back to back -- ~370 us each, every kernel a single latency-bound wave
// over 0.4 % of the cells (18 ms of the 64 ms assembly).  A strip of gh rows only reads cells within gh of itself:
// each strip is therefore cut out of the block with a margin of gh + 2 rows (columns) into a compact SUB-BLOCK of its own -- state,
// metrics and coefdiag copied by k_copy_window, the boundary list clipped to the window and re-expressed in its local indices,
// global numbering through GridDesc::ioff / joff -- and the colour loop runs on that small grid with up to 16 chains (a chain costs
// ~100 MB there).  The cut sides need no special treatment: the sub-block's ghost layers on a cut hold the parent's real cells, and
// whatever the kernels do differently next to a cut (extrapolated instead of computed sensor gradients in the first layer) is
// farther than the stencil from the strip rows.  Same COO slots and values as the full-grid loop.
namespace {

__global__ void k_copy_window(double* __restrict__ dst, int dld, long long dps, const double* __restrict__ src, int sld, long long sps,
                              int x0, int y0, int w, int h, int planes) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)w * h * planes) return;
  const int x = (int)(t % w), y = (int)((t / w) % h), p = (int)(t / ((long long)w * h));
  dst[p * dps + x + (long long)y * dld] = src[p * sps + (x0 + x) + (long long)(y0 + y) * sld];
}
cudaError_t copy_window(double* dst, int dld, long long dps, const double* src, int sld, long long sps, int x0, int y0, int w, int h,
                        int planes, cudaStream_t st) {
  const long long n = (long long)w * h * planes;
  k_copy_window<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, dld, dps, src, sld, sps, x0, y0, w, h, planes);
  return cudaGetLastError();
}

struct SubBlock {
  int ia, ib, ja, jb;   // parent-local cell range of the sub-block's interior
  int im, jm;           // its size
  bool ok;
};

// clip the boundary list of the parent to the sub-block [ia, ib] x [ja, jb] of a parent of im x jm cells and express the windows in
// its local indices; false = this list cannot be windowed (the caller falls back to the full-grid loop)
bool clip_bcs(const bc_desc_t* bcs, int nbcs, const SubBlock& sb, int pim, int pjm, int gh, bc_desc_t* out, int* nout) {
  int m = 0;
  for (int b = 0; b < nbcs; ++b) {
    bc_desc_t d = bcs[b];
    if (d.kind == BC_KIND_JOIN) return false;
    BcLine line;
    if (!decode_interface(d.loc, d.window, line)) return false;
    const bool iside = line.kdir == 1, high = line.high != 0;
    // the side must belong to the sub-block
    if (iside && (high ? sb.ib != pim : sb.ia != 1)) continue;
    if (!iside && (high ? sb.jb != pjm : sb.ja != 1)) continue;
    int imin = d.window[0], jmin = d.window[1], imax = d.window[2], jmax = d.window[3];
    int shift = 0;   // entries of the table the clipped line starts behind the original one
    if (iside) {   // line along j
      const int lo = sb.ja == 1 ? jmin : (jmin > sb.ja ? jmin : sb.ja);
      const int hi = sb.jb == pjm ? jmax : (jmax < sb.jb ? jmax : sb.jb);
      if (lo > hi) continue;
      shift = lo - jmin;
      // the inlet reads its table both by position on the line and by absolute row (bc.cuh): the two agree only if the shift
      // equals the row offset of the sub-block
      if (d.kind == BC_KIND_INLET && shift != sb.ja - 1) return false;
      jmin = lo; jmax = hi;
    } else {
      const int lo = sb.ia == 1 ? imin : (imin > sb.ia ? imin : sb.ia);
      const int hi = sb.ib == pim ? imax : (imax < sb.ib ? imax : sb.ib);
      if (lo > hi) continue;
      shift = lo - imin;
      if (d.kind == BC_KIND_INLET && shift != sb.ia - 1) return false;
      imin = lo; imax = hi;
    }
    d.window[0] = imin - (sb.ia - 1); d.window[1] = jmin - (sb.ja - 1); d.window[2] = imax - (sb.ia - 1); d.window[3] = jmax - (sb.ja - 1);
    if (d.table) d.table = d.table + shift;   // leading dimension lm of the table is unchanged
    if (m >= 16) return false;
    out[m++] = d;
  }
  (void)gh;
  *nout = m;
  return true;
}

}  // namespace

extern "C" int bcd_jacobian_strips(int nrect, const int32_t* rects /* [nrect][4] i0,i1,j0,j1 */, double* const* jac, int32_t* const* ia,
                                   int32_t* const* ja, double* w, const double* nx, const double* ny, const double* vol, const double* volf,
                                   int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref,
                                   double tref, double s_suth, double k2, double k4, int im, int jm, int wall, const bc_desc_t* bcs,
                                   int nbcs, int scatter_kind, const double* coefdiag, void* stream) {
  if (im < 1 || jm < 1 || gh != 3 || nrect < 0 || nrect > 4) return BC_ERR_ARG;
  const bool full_env = getenv("BROADCAST_B200_STRIPS_FULL") != nullptr;   // (read per call: the tests switch it)
  static const bool no_graph = getenv("BROADCAST_B200_NO_GRAPH") != nullptr;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  constexpr int M = 5;   // margin of a sub-block beyond its strip: gh for the stencil + 2 (see above)
  bool windowed = !full_env && !no_graph && nrect > 0 && g.joff == 0 && g.jmg == jm && (long long)im * jm >= 4096 && im > 3 * gh + M &&
                  jm > 3 * gh + M;
  SubBlock sb[4];
  bc_desc_t cb[4][16];
  int ncb[4] = {0, 0, 0, 0};
  for (int q = 0; q < nrect && windowed; ++q) {
    const int i0 = rects[4 * q], i1 = rects[4 * q + 1], j0 = rects[4 * q + 2], j1 = rects[4 * q + 3];
    if (i1 < i0 || j1 < j0) { windowed = false; break; }
    SubBlock& s_ = sb[q];
    if (i1 - i0 >= j1 - j0) {   // row-shaped strip: all columns, rows with the margin
      s_.ia = 1; s_.ib = im;
      s_.ja = j0 - M < 1 ? 1 : j0 - M; s_.jb = j1 + M > jm ? jm : j1 + M;
    } else {
      s_.ja = 1; s_.jb = jm;
      s_.ia = i0 - M < 1 ? 1 : i0 - M; s_.ib = i1 + M > im ? im : i1 + M;
    }
    s_.im = s_.ib - s_.ia + 1; s_.jm = s_.jb - s_.ja + 1;
    if (s_.im < 4 || s_.jm < 6) { windowed = false; break; }
    if (!clip_bcs(bcs, nbcs, s_, im, jm, gh, cb[q], &ncb[q])) windowed = false;
  }
  if (!windowed)
    return strips_impl(nrect, rects, jac, ia, ja, w, nx, ny, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4, im, jm,
                       wall, bcs, nbcs, scatter_kind, coefdiag, stream, 4);
  cudaStream_t st = (cudaStream_t)stream;
  // every ghost of the parent's w holds the value of its list before the windows are cut
  {
    cudaError_t e0 = apply_bc_list(g, gam, 0, w, nullptr, nx, ny, bcs, nbcs, st);
    if (e0 != cudaSuccess) return (int)e0;
    count_launches(nbcs);
  }
  // the four sub-block loops run side by side: the two thin column strips are latency bound (49 passes of small launches on 6 k
  // cells) and hide behind the two long row strips, which are throughput bound
  static thread_local cudaStream_t wstream[4] = {};
  static thread_local cudaEvent_t wjoin[4] = {};
  static thread_local cudaEvent_t wfork = nullptr;
  static const bool serial_env = getenv("BROADCAST_B200_STRIPS_SERIAL") != nullptr;
  const bool concurrent = !serial_env && nrect > 1;   // (events order the window streams with the legacy default stream as well)
  if (concurrent && !wfork) {
    for (int q = 0; q < 4; ++q) {
      if (cudaStreamCreateWithFlags(&wstream[q], cudaStreamNonBlocking) != cudaSuccess) return BC_ERR_ALLOC;
      if (cudaEventCreateWithFlags(&wjoin[q], cudaEventDisableTiming) != cudaSuccess) return BC_ERR_ALLOC;
    }
    if (cudaEventCreateWithFlags(&wfork, cudaEventDisableTiming) != cudaSuccess) return BC_ERR_ALLOC;
  }
  if (concurrent) {
    cudaEventRecord(wfork, st);
    for (int q = 0; q < nrect; ++q) cudaStreamWaitEvent(wstream[q], wfork, 0);
  }
  cudaStream_t st_main = st;
  const SlabInfo saved = current_slab();
  int rc = BC_OK;
  for (int q = 0; q < nrect && rc == BC_OK; ++q) {
    const SubBlock& s_ = sb[q];
    if (concurrent) st = wstream[q];
    const GridDesc gs = make_grid(s_.im, s_.jm, gh);
    // compact copies: slots 60 + 8 q .. of the scratch arena (chain 0), stable across calls so that the cached graphs replay
    scratch_chain() = 0;
    double* sw = scratch_doubles(60 + 8 * q, (size_t)gs.sc * 5);
    double* snx = scratch_doubles(61 + 8 * q, (size_t)gs.sn * 2);
    double* sny = scratch_doubles(62 + 8 * q, (size_t)gs.sn * 2);
    double* svol = scratch_doubles(63 + 8 * q, (size_t)gs.sc);
    double* svolf = scratch_doubles(64 + 8 * q, (size_t)gs.sc * 2);
    double* scd = coefdiag ? scratch_doubles(65 + 8 * q, (size_t)s_.im * s_.jm) : nullptr;
    if (!sw || !snx || !sny || !svol || !svolf || (coefdiag && !scd)) { rc = BC_ERR_ALLOC; break; }
    const int x0 = s_.ia - 1, y0 = s_.ja - 1;   // storage offset of the window in the parent's padded arrays
    cudaError_t e = copy_window(sw, gs.ldc, gs.sc, w, g.ldc, g.sc, x0, y0, gs.ni(), gs.nj(), 5, st);
    if (e == cudaSuccess) e = copy_window(snx, gs.ldn, gs.sn, nx, g.ldn, g.sn, x0, y0, gs.ni() + 1, gs.nj() + 1, 2, st);
    if (e == cudaSuccess) e = copy_window(sny, gs.ldn, gs.sn, ny, g.ldn, g.sn, x0, y0, gs.ni() + 1, gs.nj() + 1, 2, st);
    if (e == cudaSuccess) e = copy_window(svol, gs.ldc, gs.sc, vol, g.ldc, g.sc, x0, y0, gs.ni(), gs.nj(), 1, st);
    if (e == cudaSuccess) e = copy_window(svolf, gs.ldc, gs.sc, volf, g.ldc, g.sc, x0, y0, gs.ni(), gs.nj(), 2, st);
    if (e == cudaSuccess && coefdiag) e = copy_window(scd, s_.im, 0, coefdiag, im, 0, s_.ia - 1, s_.ja - 1, s_.im, s_.jm, 1, st);
    if (e != cudaSuccess) { rc = (int)e; break; }
    count_launches(coefdiag ? 6 : 5);
    // the sub-block in the numbering of the whole block: a cut in i is a slab-internal edge (computed gradients in the halo
    // column, which holds the parent's cells); rows are numbered through joff / jmg
    SlabInfo si;
    si.ioff = g.ioff + s_.ia - 1;
    si.img = g.img;
    si.edges = (s_.ia > 1 ? 1 : (g.edges & 1)) | (s_.ib < im ? 2 : (g.edges & 2));
    si.joff = s_.ja - 1;
    si.jmg = jm;
    current_slab() = si;
    const int32_t lr[4] = {rects[4 * q] - (s_.ia - 1), rects[4 * q + 1] - (s_.ia - 1), rects[4 * q + 2] - (s_.ja - 1), rects[4 * q + 3] - (s_.ja - 1)};
    double* jq[1] = {jac[q]};
    int32_t* iq[1] = {ia[q]};
    int32_t* kq[1] = {ja[q]};
    rc = strips_impl(1, lr, jq, iq, kq, sw, snx, sny, svol, svolf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4, s_.im, s_.jm,
                     (wall && s_.ja == 1) ? 1 : 0, cb[q], ncb[q], scatter_kind, scd, (void*)st, MAXCHAIN, concurrent ? MAXCHAIN * q : 0);
    current_slab() = saved;
  }
  current_slab() = saved;
  if (concurrent)
    for (int q = 0; q < nrect; ++q) {   // join every window's stream, also after an error (no work may outlive the call's stream order)
      cudaEventRecord(wjoin[q], wstream[q]);
      cudaStreamWaitEvent(st_main, wjoin[q], 0);
    }
  return rc;
}
