// Block-Jacobian assembly of the regular interior rows into the fixed 29-slot pattern
// (values only, structure of arrays: V[slot][e][m][cell]).
#include "../../include/broadcast_b200.h"
#include "jac_blocks.cuh"

namespace bcast {
void count_launches(int n);
cudaError_t prepare_prims_grads(const GridDesc& g, const SchemeArgs& a, const double* w, const double* nx, const double* ny,
                                const double* vol, const double* volf, FieldPtrs& f, cudaStream_t st);
cudaError_t launch_jacobian_faces(const GridDesc& g, const SchemeArgs& a, const double* w, const double* nx, const double* ny,
                                  const double* vol, const double* volf, const Rect& rc, double* values, const double* coefdiag,
                                  cudaStream_t st, int* counts = nullptr, double thresh = 0.0);
#define DECL(G) jac_block_fn jac_block_launcher_g##G(int, int);
DECL(0) DECL(1) DECL(2) DECL(3) DECL(4) DECL(5) DECL(6) DECL(7)
#undef DECL
jac_block_fn jac_block_launcher(int di, int dj) {
  jac_block_fn f = nullptr;
#define TRY(G) if (!f) f = jac_block_launcher_g##G(di, dj);
  TRY(0) TRY(1) TRY(2) TRY(3) TRY(4) TRY(5) TRY(6) TRY(7)
#undef TRY
  return f;
}
static const int kOffsets[JAC_NSLOT][2] = {
#define X(a, b) {a, b},
    BCAST_JAC_OFFSETS(X)
#undef X
};
}  // namespace bcast

using namespace bcast;

extern "C" int bcd_jacobian_slots(int32_t* offsets /* [29][2] */) {
  for (int s = 0; s < JAC_NSLOT; ++s) {
    offsets[2 * s] = kOffsets[s][0];
    offsets[2 * s + 1] = kOffsets[s][1];
  }
  return JAC_NSLOT;
}

extern "C" int bcd_jacobian_interior(double* values, const double* w, const double* nx, const double* ny, const double* vol,
                                     const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                                     double muref, double tref, double s_suth, double k2, double k4, int im, int jm,
                                     const double* coefdiag, const int32_t* rect, void* stream) {
  if (im < 1 || jm < 1 || gh != 3) return BC_ERR_ARG;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const SchemeArgs a{cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4};
  // regular rows: gh away from the physical sides; a slab-internal edge has real data in its halo, rows go up to it
  const int ilo = (g.edges & 1) ? 1 : gh + 1, ihi = (g.edges & 2) ? im : im - gh;
  Rect rc{ilo, ihi, gh + 1, jm - gh};
  if (rect) {
    rc = Rect{rect[0], rect[1], rect[2], rect[3]};
    if (rc.i0 < ilo || rc.i1 > ihi || rc.j0 < gh + 1 || rc.j1 > jm - gh) return BC_ERR_ARG;
  }
  if (rc.i1 < rc.i0 || rc.j1 < rc.j0) return BC_OK;
  cudaError_t e = launch_jacobian_faces(g, a, w, nx, ny, vol, volf, rc, values, coefdiag, (cudaStream_t)stream);
  return e == cudaSuccess ? BC_OK : (int)e;
}

// bcd_jacobian_interior + the per-row counts of the entries the CSR conversion keeps (|v| > thresh): `counts` (5 im jm + 1 ints, rows
// in the reference's numbering) is zeroed, then the rows of the regular region receive their counts.  bcd_hybrid_csr_indptr called
// with values = null and counted = 1 adds the strip rows and scans, without reading the block values again.
extern "C" int bcd_jacobian_interior_counted(double* values, int32_t* counts, double thresh, const double* w, const double* nx,
                                             const double* ny, const double* vol, const double* volf, int gh, double cp, double cv,
                                             double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth,
                                             double k2, double k4, int im, int jm, const double* coefdiag, const int32_t* rect,
                                             void* stream) {
  if (im < 1 || jm < 1 || gh != 3 || !counts) return BC_ERR_ARG;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const SchemeArgs a{cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4};
  const int ilo = (g.edges & 1) ? 1 : gh + 1, ihi = (g.edges & 2) ? im : im - gh;
  Rect rc{ilo, ihi, gh + 1, jm - gh};
  if (rect) {
    rc = Rect{rect[0], rect[1], rect[2], rect[3]};
    if (rc.i0 < ilo || rc.i1 > ihi || rc.j0 < gh + 1 || rc.j1 > jm - gh) return BC_ERR_ARG;
  }
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * (5LL * im * jm + 1), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  if (rc.i1 < rc.i0 || rc.j1 < rc.j0) return BC_OK;
  e = launch_jacobian_faces(g, a, w, nx, ny, vol, volf, rc, values, coefdiag, (cudaStream_t)stream, counts, thresh);
  return e == cudaSuccess ? BC_OK : (int)e;
}

extern "C" int bcd_jacobian_interior_ad(double* values, const double* w, const double* nx, const double* ny, const double* vol,
                                     const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                                     double muref, double tref, double s_suth, double k2, double k4, int im, int jm,
                                     const double* coefdiag, const int32_t* rect, void* stream) {
  if (im < 1 || jm < 1 || gh != 3) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const GridDesc g = make_grid_ctx(im, jm, gh);
  const SchemeArgs a{cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4};
  const SchemeConsts c = make_consts(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4);
  // regular rows: gh away from the physical sides; a slab-internal edge has real data in its halo, rows go up to it
  const int ilo = (g.edges & 1) ? 1 : gh + 1, ihi = (g.edges & 2) ? im : im - gh;
  Rect rc{ilo, ihi, gh + 1, jm - gh};
  if (rect) {
    rc = Rect{rect[0], rect[1], rect[2], rect[3]};
    if (rc.i0 < ilo || rc.i1 > ihi || rc.j0 < gh + 1 || rc.j1 > jm - gh) return BC_ERR_ARG;
  }
  if (rc.i1 < rc.i0 || rc.j1 < rc.j0) return BC_OK;
  FieldPtrs f;
  cudaError_t e = prepare_prims_grads(g, a, w, nx, ny, vol, volf, f, st);
  if (e != cudaSuccess) return (int)e;
  const long long slot_stride = 25LL * im * jm;
  for (int s = 0; s < JAC_NSLOT; ++s) {
    jac_block_fn fn = jac_block_launcher(kOffsets[s][0], kOffsets[s][1]);
    if (!fn) return BC_ERR_UNSUPPORTED;
    e = fn(g, c, f, rc, values + s * slot_stride, coefdiag, st);
    if (e != cudaSuccess) return (int)e;
  }
  count_launches(3 + JAC_NSLOT);
  return BC_OK;
}
