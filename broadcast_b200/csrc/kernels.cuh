// Internal C++ interface between the translation units of libbroadcast_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include "bc.cuh"
#include "grid.cuh"
#include "scheme.cuh"

namespace bcast {

// grow-only per-device scratch arena (prims / gradients / tangents of the generic path)
double* scratch_doubles(int slot, size_t count);
void scratch_release_all();

// slab descriptor of the calling thread (bcd_slab_begin / bcd_slab_end); make_grid_ctx applies it
struct SlabInfo {
  int ioff, img, edges;
  int joff = 0, jmg = 0;   // j window (strip sub-blocks); jmg = 0: none
};
SlabInfo& current_slab();
inline GridDesc make_grid_ctx(int im, int jm, int gh) {
  GridDesc g = make_grid(im, jm, gh);
  const SlabInfo& s = current_slab();
  if (s.img > 0) {
    g.ioff = s.ioff;
    g.img = s.img;
    g.edges = s.edges;
  }
  if (s.jmg > 0) {
    g.joff = s.joff;
    g.jmg = s.jmg;
  }
  return g;
}

// colour range of the calling thread (bcd_colour_range): the colour loops visit only the (l,k) passes with
// c0 <= l * (2gh+1) + k < c1 -- colour sharding over GPUs for grids too small for i-slabs
struct ColourRange {
  int c0, c1;
  bool has(int c) const { return c1 <= c0 || (c >= c0 && c < c1); }   // empty range = all colours
};
ColourRange& current_colours();

struct SchemeArgs {
  double cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4;
};

// rectangle of interior cells (Fortran indices, inclusive) a kernel is restricted to
struct Rect {
  int i0, i1, j0, j1;
};
// up to four rectangles handled by one launch (blockIdx.z selects the rectangle)
struct RectList {
  int n;
  Rect r[4];
};
inline RectList one_rect(const Rect& r) {
  RectList l;
  l.n = 1;
  l.r[0] = r;
  l.r[1] = l.r[2] = l.r[3] = Rect{1, 0, 1, 0};
  return l;
}
inline dim3 grid_of(const RectList& l, int bx, int by) {   // grid covering the largest rectangle, one z-slice per rectangle
  int wi = 1, wj = 1;
  for (int k = 0; k < l.n; ++k) {
    wi = l.r[k].i1 - l.r[k].i0 + 1 > wi ? l.r[k].i1 - l.r[k].i0 + 1 : wi;
    wj = l.r[k].j1 - l.r[k].j0 + 1 > wj ? l.r[k].j1 - l.r[k].j0 + 1 : wj;
  }
  return dim3((wi + bx - 1) / bx, (wj + by - 1) / by, l.n);
}
// Rectangles of very different shapes (the boundary strips: im x gh and gh x jm) are launched one by one instead:
// a common grid would be (im/bx) x (jm/by) blocks, almost all of them empty.  The launches of one call are independent of
// each other; when the calling thread has switched its fork context on (bcd_jacobian_strips while it captures its CUDA graph)
// each rectangle goes to its own side stream between a fork and a join event, so that the graph holds parallel branches
// instead of a chain (the strip passes are latency bound: ~8 us per dependent kernel on the reference's own grids).
struct RectFork {
  bool on = false;
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[4] = {nullptr, nullptr, nullptr, nullptr};
  bool ready();   // creates streams and events on first use
};
RectFork& rect_fork();   // the fork context of the calling thread's current chain (scratch_chain())
// Colour chains: the strip colour loop runs up to four colours at a time inside its CUDA graph, each chain with its own scratch
// buffers and side streams.  scratch_doubles(slot, n) resolves to slot + 1000 * scratch_chain().
int& scratch_chain();
template <class F>
inline void for_each_rect(const RectList& l, cudaStream_t st, F&& launch) {   // launch(one-rect list, index, stream)
  RectFork& rf_ = rect_fork();
  if (!rf_.on || l.n < 2 || !rf_.ready()) {
    for (int k = 0; k < l.n; ++k) launch(one_rect(l.r[k]), k, st);
    return;
  }
  cudaEventRecord(rf_.fork, st);
  for (int k = 0; k < l.n; ++k) {
    cudaStreamWaitEvent(rf_.side[k], rf_.fork, 0);
    launch(one_rect(l.r[k]), k, rf_.side[k]);
    cudaEventRecord(rf_.join[k], rf_.side[k]);
    cudaStreamWaitEvent(st, rf_.join[k], 0);
  }
}

// Generic ("reference-shaped") residual and tangent: prims -> gradients (+ ghost-layer extrapolation)
// -> cell-centred balance of four face fluxes.  ndir == 0: residual into `out` (5 planes, interior
// cells of `rect` only).  ndir in {1,5}: tangent(s) into `out` ([ndir][5] planes).
cudaError_t launch_residual_generic(const GridDesc& g, const SchemeArgs& a, bool wall, int ndir, double* out, const double* w,
                                    const double* wd, const double* nx, const double* ny, const double* vol, const double* volf,
                                    const Rect* rect, cudaStream_t st);

// boundary fills, ndir == 0 primal, else tangent (w and wd ghosts are both written)
cudaError_t launch_bc_wall(const GridDesc& g, const BcLine& b, double gam, int ndir, double* w, double* wd, cudaStream_t st);
cudaError_t launch_bc_noref(const GridDesc& g, const BcLine& b, double gam, int ndir, double* w, double* wd, const double* wbd, int lm,
                            const double* nx, const double* ny, cudaStream_t st);
cudaError_t launch_bc_inlet(const GridDesc& g, const BcLine& b, double gam, int ndir, double* w, double* wd, const double* field, int lm,
                            const double* nx, const double* ny, cudaStream_t st);
cudaError_t launch_bc_extrap(const GridDesc& g, const BcLine& b, int ndir, double* w, double* wd, cudaStream_t st);
cudaError_t launch_bc_general(const GridDesc& g, const BcLine& b, int ndir, double* w, double* wd, const double* field, int lm,
                              cudaStream_t st);
cudaError_t launch_bc_wall_iso(const GridDesc& g, const BcLine& b, double twall, double gam, double rgaz, int ndir, double* w, double* wd,
                               cudaStream_t st);
cudaError_t launch_bc_wall_profile(const GridDesc& g, const BcLine& b, bool blow, const double* prof, const double* profd, double gam,
                                   double gamd, double rgaz, double rgazd, int ndir, double* w, double* wd, cudaStream_t st);
cudaError_t launch_bc_symmetry(const GridDesc& g, const BcLine& b, int ndir, double* w, double* wd, const double* nx, const double* ny,
                               bool anti, cudaStream_t st);
cudaError_t launch_bc_pressure(const GridDesc& g, const BcLine& b, double pext, bool noref, double gam, int ndir, double* w, double* wd,
                               const double* nx, const double* ny, cudaStream_t st);

// rectangular window copy (jn_match): arrays described by (ld, plane stride, origin offsets)
struct Window {
  int ld;            // leading dimension
  long long stride;  // plane stride
  int lo_i, lo_j;    // Fortran lower bounds of the two dimensions
};
cudaError_t launch_jn_match(double* wr, const Window& r, const int prr[4], const double* wd, const Window& d, const int prd[4],
                            const int tr[2], int em, cudaStream_t st);

// colouring seeds and COO scatter (misc/ComputeJacobian.f90)
cudaError_t launch_testvector(const GridDesc& g, double* wd, int ndir, int m, int l, int k, const int* zone /*null or istart,iend,jstart,jend*/,
                              cudaStream_t st, const RectList* rows = nullptr);
// tangent (5 directions) of the rows of up to four rectangles in one pass; one thread per (cell, face)
cudaError_t launch_tangent_strips5(const GridDesc& g, const SchemeArgs& a, bool wall, const RectList& rows, double* out5, const double* w,
                                   const double* wd5, const double* nx, const double* ny, const double* vol, const double* volf,
                                   cudaStream_t st);
enum ScatterKind {
  SCATTER_JV = 0,            // computejacobianfromjv            :292-355
  SCATTER_JV_RELAXED = 1,    // computejacobianfromjv_relaxed    :503-570
  SCATTER_DZ = 2,            // computejacobianfromdz            :708-779
  SCATTER_JV_RELAXED_JN = 3, // computejacobianfromjv_relaxed_withjn :847-926
  SCATTER_JV_JN = 4,         // computejacobianfromjv_withjn     :929-999
  SCATTER_JV_DBYVOL = 5,     // computejacobianfromjv_dbyvol     :643-706
  SCATTER_JV_RELAXED_DBYVOL = 6,
};
// writes the contiguous slot range [base, base + 5*im*jm) of colour (m,l,k) into seg_* (length 5*im*jm)
cudaError_t launch_scatter(const GridDesc& g, int kind, double* seg_jac, int* seg_ia, int* seg_ja, const double* resd, int m, int l,
                           int k, const double* coefdiag /*im x jm or null*/, const double* vol /*cell layout or null*/, cudaStream_t st);

// two-zone variant computejacobianfromjv_relaxed_withjnandcheck (:1095-1204): read-modify-write of the slot segment of zone `zone`
cudaError_t launch_scatter_check(const GridDesc& g, double* seg_jac, int* seg_ia, int* seg_ja, const double* resd, int m, int l, int k,
                                 const double* coefdiag, double mini, int zone, cudaStream_t st);

// norms (srcfv/norm.F90)
cudaError_t launch_norms(const GridDesc& g, const double* res, double* out10 /*device: sum r^2 [5], sum r^10 [5]*/, cudaStream_t st);

// fused, shared-memory tiled primal residual (residual_tile.cu)
// variant: RES_DEFAULT (= k_residual_fast), RES_FAST_TMA (persistent CTAs + TMA staging of w), RES_TILE_V1 (first generation)
//          RES_DEFAULT = k_residual_fast (32 x 9 tile kernel); RES_MARCH = k_residual_march (j-marching persistent kernel, rings fed by
//          TMA / bulk copies, residual_march.cu; falls back to the tile kernel when TMA cannot describe the arrays).  Measured at C5
//          (profiles/r2_a_summary.md): tile 2.385 ms, march 2.478 ms -- the faster one is the default, BROADCAST_B200_RESIDUAL_MARCH=1
//          swaps them.  RES_FAST_TILE always names the tile kernel.
//          RES_FAST_BULK = k_residual_fast_bulk (residual_bulk.cu): the tile kernel with w, vol, volf and the node rows of nx / ny delivered
//          by TMA / bulk copies on one mbarrier, metrics read from shared memory; bit-identical to the tile kernel.
enum ResidualVariant { RES_DEFAULT = 0, RES_GENERIC = 1, RES_FAST_TMA = 2, RES_TILE_V1 = 3, RES_FAST_TILE = 4, RES_MARCH = 5, RES_FAST_BULK = 6 };
cudaError_t launch_residual_tiled(const GridDesc& g, const SchemeArgs& a, bool wall, double* res, const double* w, const double* nx,
                                  const double* ny, const double* vol, const double* volf, cudaStream_t st, int variant = RES_DEFAULT,
                                  int part = 0);

// second-generation fused residual (residual_fast.cu): re-associated face formulas, shared normal-direction interpolations
cudaError_t launch_residual_fast(const GridDesc& g, const SchemeArgs& a, bool wall, double* res, const double* w, const double* nx,
                                 const double* ny, const double* vol, const double* volf, cudaStream_t st, bool tma, int part = 0);

// bulk-staged tile kernel (residual_bulk.cu); *done = false -> caller falls back to the LDG tile kernel
cudaError_t launch_residual_fast_bulk(const GridDesc& g, const SchemeConsts& c, double sqgr, bool wall, double* res, const double* w,
                                      const double* nx, const double* ny, const double* vol, const double* volf, cudaStream_t st, bool* done,
                                      int part = 0);

// third-generation fused residual (residual_march.cu): persistent j-marching CTAs; *done = false -> caller falls back
cudaError_t launch_residual_march(const GridDesc& g, const SchemeConsts& c, double sqgr, bool wall, double* res, const double* w,
                                  const double* nx, const double* ny, const double* vol, const double* volf, cudaStream_t st, bool* done);

}  // namespace bcast
