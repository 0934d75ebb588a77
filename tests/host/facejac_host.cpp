// Host build of the product's block-Jacobian math (broadcast_b200/csrc/facejac.cuh is host+device code):
// TEST INFRASTRUCTURE so that the semi-analytic face linearisation can be checked against the oracle on a
// machine without a GPU.  Not part of the product; the product runs the same templates in CUDA kernels.
#include <vector>
#include "../../broadcast_b200/csrc/facejac.cuh"

using namespace bcast;

static const JacTab kTab = fj::make_jac_tab();

extern "C" int fj_host_blocks_impl(int table_driven, double* values, const double* w, const double* nx, const double* ny, const double* vol,
                              const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                              double muref, double tref, double s_suth, double k2, double k4, int im, int jm);
extern "C" int fj_host_blocks(double* values /* [29][25][im*jm] */, const double* w, const double* nx, const double* ny, const double* vol,
                              const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                              double muref, double tref, double s_suth, double k2, double k4, int im, int jm) {
  return fj_host_blocks_impl(0, values, w, nx, ny, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4, im, jm);
}
extern "C" int fj_host_blocks_table(double* values, const double* w, const double* nx, const double* ny, const double* vol,
                              const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                              double muref, double tref, double s_suth, double k2, double k4, int im, int jm) {
  return fj_host_blocks_impl(1, values, w, nx, ny, vol, volf, gh, cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4, im, jm);
}
extern "C" int fj_host_blocks_impl(int table_driven, double* values, const double* w, const double* nx, const double* ny, const double* vol,
                              const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                              double muref, double tref, double s_suth, double k2, double k4, int im, int jm) {
  const GridDesc g = make_grid(im, jm, gh);
  const SchemeConsts c = make_consts(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4);
  std::vector<double> prim((size_t)NPRIM * g.sc, 0.0), grad((size_t)NGRAD * g.sc, 0.0), pkg((size_t)2 * FPK_N * g.sc, 0.0);
  for (int jj = 0; jj < g.nj(); ++jj)
    for (int ii = 0; ii < g.ni(); ++ii) {
      const long long k = ii + (long long)jj * g.ldc;
      PVar q[5];
      for (int e = 0; e < 5; ++e) q[e].v = w[e * g.sc + k];
      const CellPrims<Zero> p = cell_prims(q, c);
      const double out[NPRIM] = {p.u.v, p.v.v, p.w.v, p.t.v, p.p.v, p.mu.v, p.h.v};
      for (int s = 0; s < NPRIM; ++s) prim[s * g.sc + k] = out[s];
    }
  FieldPtrs f{w, prim.data(), grad.data(), nx, ny, vol, volf, nullptr, nullptr, nullptr};
  for (int j = 1; j <= jm; ++j)
    for (int i = 1; i <= im; ++i) {
      GlobalAcc<0> a(f, g, i, j);
      const auto r = cell_gradients<0, 0>(a);
      const long long k = g.cidx(i, j);
      grad[0 * g.sc + k] = r.u0.v;
      grad[1 * g.sc + k] = r.u1.v;
      grad[2 * g.sc + k] = r.v0.v;
      grad[3 * g.sc + k] = r.v1.v;
    }
  const int i0 = gh + 1, i1 = im - gh, j0 = gh + 1, j1 = jm - gh;
  for (int j = j0; j <= j1 + 1; ++j)
    for (int i = i0; i <= i1 + 1; ++i) {
      GlobalAcc<0> a(f, g, i, j);
      const long long k = g.cidx(i, j);
      face_package<0>(a, c, [&](int fld, double v) { pkg[(size_t)(0 * FPK_N + fld) * g.sc + k] = v; });
      face_package<1>(a, c, [&](int fld, double v) { pkg[(size_t)(1 * FPK_N + fld) * g.sc + k] = v; });
    }
  const long long ncell = (long long)im * jm;
  auto ctx = [&](int dir, int i, int j) {
    FaceCtx x;
    x.pk = pkg.data() + (size_t)dir * FPK_N * g.sc + g.cidx(i, j);
    x.stride = g.sc;
    x.vs = x.pk + (size_t)FPK_VS * g.sc;
    x.vstride = (int)g.sc;
    return x;
  };
  for (int j = j0; j <= j1; ++j)
    for (int i = i0; i <= i1; ++i) {
      const FaceCtx fi0 = ctx(0, i, j), fi1 = ctx(0, i + 1, j), fj0 = ctx(1, i, j), fj1 = ctx(1, i, j + 1);
      const long long cell = (long long)(i - 1) + (long long)(j - 1) * im;
      if (table_driven) {
        for (int s = 0; s < JAC_NSLOT; ++s) {
          double wc[5], B[25];
          for (int e = 0; e < 5; ++e) wc[e] = w[e * g.sc + g.cidx(i + kTab.di[s], j + kTab.dj[s])];
          block_of_rt(kTab, s, fi0, fi1, fj0, fj1, wc, c, B);
          for (int q = 0; q < 25; ++q) values[((long long)s * 25 + q) * ncell + cell] = B[q];
        }
        continue;
      }
      int slot = 0;
#define X(DI, DJ)                                                                                  \
  {                                                                                                \
    double wc[5], B[25];                                                                           \
    for (int e = 0; e < 5; ++e) wc[e] = w[e * g.sc + g.cidx(i + (DI), j + (DJ))];                  \
    block_of<DI, DJ>(fi0, fi1, fj0, fj1, wc, c, B);                                                \
    for (int q = 0; q < 25; ++q) values[((long long)slot * 25 + q) * ncell + cell] = B[q];         \
    ++slot;                                                                                        \
  }
      BCAST_JAC_OFFSETS(X)
#undef X
    }
  return 0;
}
