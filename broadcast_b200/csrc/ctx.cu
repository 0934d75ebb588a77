// Resident-mode context of the C ABI (bcast_ctx_create / destroy, ..._upload_state, ..._residual, ..._jacobian_csr,
// ..._download): one structured block kept on the device between calls, for C / Fortran (ISO_C_BINDING) hosts that have no torch
// to own device memory.  Pure orchestration: every number is produced by the bcd_* entry points (the same ones the Python
// resident layer, broadcast_b200/resident.py, drives), in the order of the reference drivers:
//   residual step   BROADCAST_npz.py:1018-1035  (boundary fills, flux_num_dnc5_2d, compute_norml2inf)
//   Jacobian        BROADCAST_npz.py:1068-1127, 129-135, 1206-1209, misc/PETSc_func.py:71-95
//                   (colour loop -> remove_zero_jac -> csr_matrix -> division by the cell volume) = hybrid assembly + csr.cu
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/broadcast_b200.h"

struct bcast_ctx {
  int im, jm, gh, wall;
  double phys[11];   // cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4
  long long sc, sn;  // plane sizes of cell / node arrays
  double *w = nullptr, *res = nullptr, *nx = nullptr, *ny = nullptr, *vol = nullptr, *volf = nullptr, *out10 = nullptr;
  bool have_geom = false, have_state = false, have_res = false;
  std::vector<bc_desc_t> bcs;     // tables = device copies owned by the context
  std::vector<double*> tables;
  bool has_join = false;
  // Jacobian
  double* blocks = nullptr;       // 29 x 25 x im x jm
  double* coefdiag = nullptr;     // im x jm
  int nstrip = 0;
  int32_t srect[16];
  double* sjac[4] = {nullptr, nullptr, nullptr, nullptr};
  int32_t *sia[4] = {nullptr, nullptr, nullptr, nullptr}, *sja[4] = {nullptr, nullptr, nullptr, nullptr};
  long long slen[4] = {0, 0, 0, 0};
  long long* indptr = nullptr;
  int32_t* counts = nullptr;
  long long* bsum = nullptr;
  int32_t* indices = nullptr;
  double* data = nullptr;
  long long nnz = -1, nnz_cap = 0;
  cudaStream_t st = nullptr;
};

namespace {
template <class T>
int dalloc(T** p, size_t n) {
  if (*p) return BC_OK;
  return cudaMalloc((void**)p, n * sizeof(T)) == cudaSuccess ? BC_OK : BC_ERR_ALLOC;
}
#define CTX_CK(call)                              \
  do {                                            \
    cudaError_t e__ = (call);                     \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)
#define CTX_RC(call)          \
  do {                        \
    int rc__ = (call);        \
    if (rc__) return rc__;    \
  } while (0)
void free_tables(bcast_ctx* c) {
  for (double* t : c->tables) cudaFree(t);
  c->tables.clear();
  c->bcs.clear();
  c->has_join = false;
}
}  // namespace

extern "C" int bcast_ctx_create(bcast_ctx_t** out, int im, int jm, int gh, double cp, double cv, double prandtl, double gam, double rgaz,
                                double cs, double muref, double tref, double s_suth, double k2, double k4, int wall) {
  if (!out) return BC_ERR_ARG;
  *out = nullptr;
  if (bc_device_count() <= 0) return BC_ERR_NODEV;
  if (im < 1 || jm < 1) return BC_ERR_ARG;
  // gh = 2 / 3 / 4 / 5: flux_num_dnc3 / 5 / 7 / 9 (state, fills, residual, norms); the block-Jacobian -> CSR entry is order 5
  if (gh < 2 || gh > 5) return BC_ERR_UNSUPPORTED;
  bcast_ctx* c = new bcast_ctx();
  c->im = im; c->jm = jm; c->gh = gh; c->wall = wall ? 1 : 0;
  const double p[11] = {cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4};
  std::memcpy(c->phys, p, sizeof p);
  c->sc = (long long)(im + 2 * gh) * (jm + 2 * gh);
  c->sn = (long long)(im + 2 * gh + 1) * (jm + 2 * gh + 1);
  int rc = dalloc(&c->w, c->sc * 5);
  if (!rc) rc = dalloc(&c->res, c->sc * 5);
  if (!rc) rc = dalloc(&c->nx, c->sn * 2);
  if (!rc) rc = dalloc(&c->ny, c->sn * 2);
  if (!rc) rc = dalloc(&c->vol, c->sc);
  if (!rc) rc = dalloc(&c->volf, c->sc * 2);
  if (!rc) rc = dalloc(&c->out10, 16);
  if (!rc && cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) rc = BC_ERR_ALLOC;
  if (!rc && cudaMemsetAsync(c->res, 0, sizeof(double) * c->sc * 5, c->st) != cudaSuccess) rc = BC_ERR_ALLOC;
  if (rc) {
    bcast_ctx_destroy(c);
    return rc;
  }
  *out = c;
  return BC_OK;
}

extern "C" int bcast_ctx_destroy(bcast_ctx_t* c) {
  if (!c) return BC_OK;
  if (c->st) cudaStreamSynchronize(c->st);
  free_tables(c);
  for (double* p : {c->w, c->res, c->nx, c->ny, c->vol, c->volf, c->out10, c->blocks, c->coefdiag, c->data}) cudaFree(p);
  for (int k = 0; k < 4; ++k) {
    cudaFree(c->sjac[k]);
    cudaFree(c->sia[k]);
    cudaFree(c->sja[k]);
  }
  cudaFree(c->indptr);
  cudaFree(c->counts);
  cudaFree(c->bsum);
  cudaFree(c->indices);
  if (c->st) cudaStreamDestroy(c->st);
  delete c;
  return BC_OK;
}

// nx, ny, vol, volf: HOST arrays as f_geom.computegeom_2d leaves them (BROADCAST_npz.py:702)
extern "C" int bcast_ctx_set_geometry(bcast_ctx_t* c, const double* nx, const double* ny, const double* vol, const double* volf) {
  if (!c || !nx || !ny || !vol || !volf) return BC_ERR_ARG;
  CTX_CK(cudaMemcpyAsync(c->nx, nx, sizeof(double) * c->sn * 2, cudaMemcpyHostToDevice, c->st));
  CTX_CK(cudaMemcpyAsync(c->ny, ny, sizeof(double) * c->sn * 2, cudaMemcpyHostToDevice, c->st));
  CTX_CK(cudaMemcpyAsync(c->vol, vol, sizeof(double) * c->sc, cudaMemcpyHostToDevice, c->st));
  CTX_CK(cudaMemcpyAsync(c->volf, volf, sizeof(double) * c->sc * 2, cudaMemcpyHostToDevice, c->st));
  CTX_CK(cudaStreamSynchronize(c->st));
  c->have_geom = true;
  return BC_OK;
}

// the driver's ordered boundary list; here the tables are HOST Fortran arrays (field(lm,gh,5), wbd(lm,5)): copied to the device
extern "C" int bcast_ctx_set_bcs(bcast_ctx_t* c, const bc_desc_t* bcs, int nbcs) {
  if (!c || nbcs < 0 || (nbcs && !bcs)) return BC_ERR_ARG;
  // the new list is built aside and swapped in only when every entry was accepted: an error leaves the installed list untouched
  std::vector<bc_desc_t> list;
  std::vector<double*> tabs;
  bool join = false;
  auto fail = [&](int code) {
    for (double* t : tabs) cudaFree(t);
    return code;
  };
  for (int k = 0; k < nbcs; ++k) {
    bc_desc_t d = bcs[k];
    size_t n = 0;
    if (d.kind == BC_KIND_INLET) n = (size_t)d.lm * c->gh * 5;
    else if (d.kind == BC_KIND_NOREF) n = (size_t)d.lm * 5;
    else if (d.kind == BC_KIND_WALL_BLOW_PROFILE || d.kind == BC_KIND_WALL_ISO_PROFILE) n = (size_t)d.lm;
    else if (d.kind == BC_KIND_JOIN) join = true;
    else if (d.kind != BC_KIND_EXTRAP && d.kind != BC_KIND_WALL && (d.kind < BC_KIND_WALL_ISO || d.kind > BC_KIND_WALL_ISO_PROFILE)) return fail(BC_ERR_ARG);
    if (n) {
      if (!d.table || d.lm < 1) return fail(BC_ERR_ARG);
      double* t = nullptr;
      if (cudaMalloc((void**)&t, n * sizeof(double)) != cudaSuccess) return fail(BC_ERR_ALLOC);
      tabs.push_back(t);
      if (cudaMemcpy(t, d.table, n * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return fail(BC_ERR_ALLOC);
      d.table = t;
    } else {
      d.table = nullptr;
    }
    list.push_back(d);
  }
  free_tables(c);
  c->bcs.swap(list);
  c->tables.swap(tabs);
  c->has_join = join;
  return BC_OK;
}

extern "C" int bcast_ctx_upload_state(bcast_ctx_t* c, const double* w) {
  if (!c || !w) return BC_ERR_ARG;
  CTX_CK(cudaMemcpyAsync(c->w, w, sizeof(double) * c->sc * 5, cudaMemcpyHostToDevice, c->st));
  CTX_CK(cudaStreamSynchronize(c->st));
  c->have_state = true;
  c->have_res = false;
  c->nnz = -1;
  return BC_OK;
}

extern "C" int bcast_ctx_download_state(bcast_ctx_t* c, double* w) {
  if (!c || !w || !c->have_state) return BC_ERR_ARG;
  CTX_CK(cudaMemcpyAsync(w, c->w, sizeof(double) * c->sc * 5, cudaMemcpyDeviceToHost, c->st));
  CTX_CK(cudaStreamSynchronize(c->st));
  return BC_OK;
}

extern "C" int bcast_ctx_apply_bcs(bcast_ctx_t* c) {
  if (!c || !c->have_state || !c->have_geom) return BC_ERR_ARG;
  if (c->bcs.empty()) return BC_OK;
  return bcd_apply_bcs(c->w, c->nx, c->ny, c->phys[3], c->gh, c->im, c->jm, c->bcs.data(), (int)c->bcs.size(), c->st);
}

// boundary fills + residual of the resident state (asynchronous; the download / norm calls synchronise)
extern "C" int bcast_ctx_residual(bcast_ctx_t* c) {
  if (!c) return BC_ERR_ARG;
  CTX_RC(bcast_ctx_apply_bcs(c));
  const double* p = c->phys;
  CTX_RC(bcd_residual(c->res, c->w, c->nx, c->ny, c->vol, c->volf, c->gh, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10],
                      c->im, c->jm, c->wall, 0, c->st));
  c->have_res = true;
  return BC_OK;
}

extern "C" int bcast_ctx_download_residual(bcast_ctx_t* c, double* res) {
  if (!c || !res || !c->have_res) return BC_ERR_ARG;
  CTX_CK(cudaMemcpyAsync(res, c->res, sizeof(double) * c->sc * 5, cudaMemcpyDeviceToHost, c->st));
  CTX_CK(cudaStreamSynchronize(c->st));
  return BC_OK;
}

// f_norm.compute_norml2inf of the resident residual (srcfv/norm.F90:34-77): norm = sqrt(sum r^2), ninf = (sum r^10)^0.1
extern "C" int bcast_ctx_norms(bcast_ctx_t* c, double* norm5, double* ninf5) {
  if (!c || !norm5 || !ninf5 || !c->have_res) return BC_ERR_ARG;
  CTX_RC(bcd_norm_sums(c->out10, c->res, c->im, c->jm, c->gh, c->st));
  double h[10];
  CTX_CK(cudaMemcpyAsync(h, c->out10, sizeof h, cudaMemcpyDeviceToHost, c->st));
  CTX_CK(cudaStreamSynchronize(c->st));
  for (int e = 0; e < 5; ++e) {
    norm5[e] = std::sqrt(h[e]);
    ninf5[e] = std::pow(h[5 + e], 0.1);
  }
  return BC_OK;
}

// Jacobian of the resident state as the CSR row block the reference's PETSc path consumes.  coefdiag: HOST (im, jm) Fortran array
// of the relaxation term (BROADCAST_npz.py:1127) or NULL (= 0); divide_by_vol != 0 applies BROADCAST_npz.py:1206-1209;
// thresh = remove_zero_jac's 2e-16; scatter_kind: 1 (computejacobianfromjv_relaxed), 3 (..._relaxed_withjn) or -1 = 3 iff the
// boundary list holds a join (the cylinder driver).  The boundary fills are applied to the state first, as the drivers do.
extern "C" int bcast_ctx_jacobian_csr(bcast_ctx_t* c, const double* coefdiag, int divide_by_vol, double thresh, int scatter_kind,
                                      long long* nnz) {
  if (!c || !nnz || !c->have_state || !c->have_geom) return BC_ERR_ARG;
  const int im = c->im, jm = c->jm, gh = c->gh;
  if (gh != 3) return BC_ERR_UNSUPPORTED;   // (the other orders assemble through bcd_jacobian_coo on device pointers)
  if (im < 2 * gh || jm < 2 * gh) return BC_ERR_UNSUPPORTED;
  if (scatter_kind < 0) scatter_kind = c->has_join ? 3 : 1;
  if (scatter_kind != 1 && scatter_kind != 3) return BC_ERR_ARG;
  const double* p = c->phys;
  const long long ncell = (long long)im * jm, n = 5 * ncell;
  const bool fresh_blocks = !c->blocks;
  CTX_RC(dalloc(&c->blocks, (size_t)29 * 25 * ncell));
  CTX_RC(dalloc(&c->coefdiag, (size_t)ncell));
  if (fresh_blocks) CTX_CK(cudaMemsetAsync(c->blocks, 0, sizeof(double) * 29 * 25 * ncell, c->st));
  if (coefdiag)
    CTX_CK(cudaMemcpyAsync(c->coefdiag, coefdiag, sizeof(double) * ncell, cudaMemcpyHostToDevice, c->st));
  else
    CTX_CK(cudaMemsetAsync(c->coefdiag, 0, sizeof(double) * ncell, c->st));
  CTX_RC(bcast_ctx_apply_bcs(c));
  // regular rows: face-linearisation block kernels, no colouring
  CTX_RC(bcd_jacobian_interior(c->blocks, c->w, c->nx, c->ny, c->vol, c->volf, gh, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9],
                               p[10], im, jm, c->coefdiag, nullptr, c->st));
  // irregular rows: the reference colour loop restricted to the four boundary strips
  if (!c->nstrip) {
    const int32_t r[4][4] = {{1, im, 1, gh}, {1, im, jm - gh + 1, jm}, {1, gh, gh + 1, jm - gh}, {im - gh + 1, im, gh + 1, jm - gh}};
    for (int k = 0; k < 4; ++k) {
      if (r[k][1] < r[k][0] || r[k][3] < r[k][2]) continue;
      const int q = c->nstrip++;
      std::memcpy(c->srect + 4 * q, r[k], sizeof r[k]);
      const int s = 2 * gh + 1;
      c->slen[q] = 25LL * s * s * (r[k][1] - r[k][0] + 1) * (r[k][3] - r[k][2] + 1);
      CTX_RC(dalloc(&c->sjac[q], (size_t)c->slen[q]));
      CTX_RC(dalloc(&c->sia[q], (size_t)c->slen[q]));
      CTX_RC(dalloc(&c->sja[q], (size_t)c->slen[q]));
      CTX_CK(cudaMemsetAsync(c->sjac[q], 0, sizeof(double) * c->slen[q], c->st));
      CTX_CK(cudaMemsetAsync(c->sia[q], 0, sizeof(int32_t) * c->slen[q], c->st));
      CTX_CK(cudaMemsetAsync(c->sja[q], 0, sizeof(int32_t) * c->slen[q], c->st));
    }
  }
  CTX_RC(bcd_jacobian_strips(c->nstrip, c->srect, c->sjac, c->sia, c->sja, c->w, c->nx, c->ny, c->vol, c->volf, gh, p[0], p[1], p[2], p[3],
                             p[4], p[5], p[6], p[7], p[8], p[9], p[10], im, jm, c->wall, c->bcs.data(), (int)c->bcs.size(), scatter_kind,
                             c->coefdiag, c->st));
  // zero filter + CSR (+ division by the volume)
  CTX_RC(dalloc(&c->indptr, (size_t)n + 1));
  CTX_RC(dalloc(&c->counts, (size_t)n + 1));
  CTX_RC(dalloc(&c->bsum, (size_t)n / 2048 + 2));
  const int32_t region[4] = {gh + 1, im - gh, gh + 1, jm - gh};
  CTX_RC(bcd_hybrid_csr_indptr(c->indptr, c->counts, c->bsum, c->blocks, region, c->nstrip, c->sjac, c->sia, c->slen, thresh, gh, im, jm,
                               c->st));
  long long total = 0;
  CTX_CK(cudaMemcpyAsync(&total, c->indptr + n, sizeof total, cudaMemcpyDeviceToHost, c->st));
  CTX_CK(cudaStreamSynchronize(c->st));
  if (total > c->nnz_cap) {
    cudaFree(c->indices);
    cudaFree(c->data);
    c->indices = nullptr;
    c->data = nullptr;
    c->nnz_cap = 0;
    CTX_RC(dalloc(&c->indices, (size_t)total));
    CTX_RC(dalloc(&c->data, (size_t)total));
    c->nnz_cap = total;
  }
  CTX_RC(bcd_hybrid_csr_fill(c->indices, c->data, c->counts, c->indptr, c->blocks, region, c->nstrip, c->srect, c->sjac, c->sia, c->sja,
                             c->slen, thresh, divide_by_vol ? c->vol : nullptr, gh, im, jm, c->st));
  CTX_CK(cudaStreamSynchronize(c->st));
  c->nnz = total;
  *nnz = total;
  return BC_OK;
}

// indptr: 5 im jm + 1 int64; indices / data: nnz entries (rows e + 5 (j-1) + 5 jm (i-1), columns ascending in a row)
extern "C" int bcast_ctx_download_csr(bcast_ctx_t* c, long long* indptr, int32_t* indices, double* data) {
  if (!c || !indptr || !indices || !data || c->nnz < 0) return BC_ERR_ARG;
  const long long n = 5LL * c->im * c->jm;
  CTX_CK(cudaMemcpyAsync(indptr, c->indptr, sizeof(long long) * (n + 1), cudaMemcpyDeviceToHost, c->st));
  CTX_CK(cudaMemcpyAsync(indices, c->indices, sizeof(int32_t) * c->nnz, cudaMemcpyDeviceToHost, c->st));
  CTX_CK(cudaMemcpyAsync(data, c->data, sizeof(double) * c->nnz, cudaMemcpyDeviceToHost, c->st));
  CTX_CK(cudaStreamSynchronize(c->st));
  return BC_OK;
}
