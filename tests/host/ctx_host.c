/* A C host of the resident C-ABI context (include/broadcast_b200.h, bcast_ctx_*): no Python, no torch -- what a Fortran / C driver
 * of the reference would link.  TEST PROGRAM: reads one block (sizes, physics, geometry, boundary list, state) from a flat binary
 * file written by tests/test_ctx_host.py, runs boundary fills + residual + norms + Jacobian -> CSR on the GPU and writes the results
 * to a second file; the test compares them with the Python paths.  Compiled as C (gcc) to prove the header is plain C. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/broadcast_b200.h"

#define RD(ptr, n) do { if (fread((ptr), sizeof(*(ptr)), (size_t)(n), f) != (size_t)(n)) { fprintf(stderr, "short read\n"); return 2; } } while (0)
#define CK(call) do { int rc_ = (call); if (rc_) { fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, bc_last_error()); return 3; } } while (0)

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int32_t hd[5];           /* im, jm, gh, wall, nbcs */
  double phys[11];
  RD(hd, 5);
  RD(phys, 11);
  const int im = hd[0], jm = hd[1], gh = hd[2], nbcs = hd[4];
  const size_t sc = (size_t)(im + 2 * gh) * (jm + 2 * gh), sn = (size_t)(im + 2 * gh + 1) * (jm + 2 * gh + 1);
  double *nx = malloc(sn * 2 * 8), *ny = malloc(sn * 2 * 8), *vol = malloc(sc * 8), *volf = malloc(sc * 2 * 8), *w = malloc(sc * 5 * 8);
  double *res = malloc(sc * 5 * 8), *coef = malloc((size_t)im * jm * 8);
  RD(nx, sn * 2); RD(ny, sn * 2); RD(vol, sc); RD(volf, sc * 2); RD(w, sc * 5); RD(coef, (size_t)im * jm);
  bc_desc_t* bcs = calloc((size_t)(nbcs ? nbcs : 1), sizeof(bc_desc_t));
  for (int k = 0; k < nbcs; ++k) {
    int32_t rec[16];       /* kind, loc[4 chars as ints], window[4], prd[4], tr[2], lm */
    RD(rec, 16);
    bcs[k].kind = rec[0];
    for (int q = 0; q < 4; ++q) bcs[k].loc[q] = (char)rec[1 + q];
    memcpy(bcs[k].window, rec + 5, 16);
    memcpy(bcs[k].prd, rec + 9, 16);
    memcpy(bcs[k].tr, rec + 13, 8);
    bcs[k].lm = rec[15];
    RD(bcs[k].param, 2);
    size_t n = bcs[k].kind == BC_KIND_INLET ? (size_t)rec[15] * gh * 5 : bcs[k].kind == BC_KIND_NOREF ? (size_t)rec[15] * 5
               : bcs[k].kind >= BC_KIND_WALL_BLOW_PROFILE ? (size_t)rec[15] : 0;
    if (n) {
      double* t = malloc(n * 8);
      RD(t, n);
      bcs[k].table = t;
    }
  }
  fclose(f);

  bcast_ctx_t* ctx = NULL;
  CK(bcast_ctx_create(&ctx, im, jm, gh, phys[0], phys[1], phys[2], phys[3], phys[4], phys[5], phys[6], phys[7], phys[8], phys[9], phys[10],
                      hd[3]));
  CK(bcast_ctx_set_geometry(ctx, nx, ny, vol, volf));
  CK(bcast_ctx_set_bcs(ctx, bcs, nbcs));
  CK(bcast_ctx_upload_state(ctx, w));
  CK(bcast_ctx_residual(ctx));
  CK(bcast_ctx_download_residual(ctx, res));
  double n2[5], ninf[5];
  CK(bcast_ctx_norms(ctx, n2, ninf));
  long long nnz = 0;
  CK(bcast_ctx_jacobian_csr(ctx, coef, 1, 2e-16, -1, &nnz));
  const long long nrow = 5LL * im * jm;
  long long* indptr = malloc((size_t)(nrow + 1) * 8);
  int32_t* indices = malloc((size_t)nnz * 4);
  double* data = malloc((size_t)nnz * 8);
  CK(bcast_ctx_download_csr(ctx, indptr, indices, data));
  CK(bcast_ctx_destroy(ctx));

  FILE* o = fopen(argv[2], "wb");
  if (!o) return 1;
  fwrite(&nnz, 8, 1, o);
  fwrite(n2, 8, 5, o);
  fwrite(ninf, 8, 5, o);
  fwrite(res, 8, sc * 5, o);
  fwrite(indptr, 8, (size_t)(nrow + 1), o);
  fwrite(indices, 4, (size_t)nnz, o);
  fwrite(data, 8, (size_t)nnz, o);
  fclose(o);
  printf("ctx_host: %d x %d cells, nnz = %lld, launches = %lld\n", im, jm, nnz, bc_launch_count());
  return 0;
}
