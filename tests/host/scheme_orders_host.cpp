// Host build of the scheme-family templates (broadcast_b200/csrc/scheme.cuh is host+device code): the reference-shaped pipeline
// primitives -> cell gradients -> ghost gradients -> four faces per cell -> balance, exactly what k_prims / k_grads / k_grad_ghost /
// k_balance of generic_impl.cuh do on the device, for the orders 3 / 5 / 7 / 9 in passive and one-direction tangent arithmetic.
// TEST INFRASTRUCTURE (lets the order tables be checked against the oracle on a machine without a GPU); not part of the product.
#include <vector>
#include "../../broadcast_b200/csrc/scheme.cuh"

using namespace bcast;

template <int N, int ORD>
static int run(double* out, const double* w, const double* wd, const double* nx, const double* ny, const double* vol, const double* volf,
               const GridDesc& g, const SchemeConsts& c, bool wall) {
  using DT = TanOf<N>;
  std::vector<double> prim((size_t)NPRIM * g.sc, 0.0), grad((size_t)NGRAD * g.sc, 0.0);
  std::vector<double> primd(N ? (size_t)NPRIM * g.sc : 1, 0.0), gradd(N ? (size_t)NGRAD * g.sc : 1, 0.0);
  for (int jj = 0; jj < g.nj(); ++jj)
    for (int ii = 0; ii < g.ni(); ++ii) {
      const long long k = ii + (long long)jj * g.ldc;
      Var<DT> q[5];
      for (int e = 0; e < 5; ++e) {
        q[e].v = w[e * g.sc + k];
        if constexpr (N > 0) q[e].d.d[0] = wd[e * g.sc + k];
      }
      const CellPrims<DT> p = cell_prims(q, c);
      const Var<DT> o[NPRIM] = {p.u, p.v, p.w, p.t, p.p, p.mu, p.h};
      for (int s = 0; s < NPRIM; ++s) {
        prim[s * g.sc + k] = o[s].v;
        if constexpr (N > 0) primd[s * g.sc + k] = o[s].d.d[0];
      }
    }
  FieldPtrs f{w, prim.data(), grad.data(), nx, ny, vol, volf, wd, primd.data(), gradd.data()};
  for (int j = 1; j <= g.jm; ++j)
    for (int i = 1; i <= g.im; ++i) {
      GlobalAcc<N> a(f, g, i, j);
      const auto r = cell_gradients<0, 0, ORD>(a);
      const long long k = g.cidx(i, j);
      const decltype(r.u0) o[NGRAD] = {r.u0, r.u1, r.v0, r.v1};
      for (int s = 0; s < NGRAD; ++s) {
        grad[s * g.sc + k] = o[s].v;
        if constexpr (N > 0) gradd[s * g.sc + k] = o[s].d.d[0];
      }
    }
  // rhs/gradveloingh.F: first ghost layer by linear extrapolation
  auto ghost = [&](std::vector<double>& a) {
    for (int s = 0; s < NGRAD; ++s) {
      double* p = a.data() + (size_t)s * g.sc;
      for (int i = 1; i <= g.im; ++i) {
        p[g.cidx(i, 0)] = 2.0 * p[g.cidx(i, 1)] - p[g.cidx(i, 2)];
        p[g.cidx(i, g.jm + 1)] = 2.0 * p[g.cidx(i, g.jm)] - p[g.cidx(i, g.jm - 1)];
      }
      for (int j = 1; j <= g.jm; ++j) {
        p[g.cidx(0, j)] = 2.0 * p[g.cidx(1, j)] - p[g.cidx(2, j)];
        p[g.cidx(g.im + 1, j)] = 2.0 * p[g.cidx(g.im, j)] - p[g.cidx(g.im - 1, j)];
      }
    }
  };
  ghost(grad);
  if (N) ghost(gradd);
  for (int j = 1; j <= g.jm; ++j)
    for (int i = 1; i <= g.im; ++i) {
      Var<DT> a0[5], a1[5], b0[5], b1[5];
      face_by_row<0, ORD>(GlobalAcc<N>(f, g, i, j), c, wall, j, a0);
      face_by_row<0, ORD>(GlobalAcc<N>(f, g, i + 1, j), c, wall, j, a1);
      face_by_row<1, ORD>(GlobalAcc<N>(f, g, i, j), c, wall, j, b0);
      face_by_row<1, ORD>(GlobalAcc<N>(f, g, i, j + 1), c, wall, j + 1, b1);
      const long long k = g.cidx(i, j);
      for (int e = 0; e < 5; ++e) {
        const Var<DT> r = -(a1[e] - a0[e]) - (b1[e] - b0[e]);
        if constexpr (N == 0) out[e * g.sc + k] = r.v;
        else out[e * g.sc + k] = r.d.d[0];
      }
    }
  return 0;
}

// out: residual (wd null) or tangent of the residual along wd, interior cells of a (im + 2 gh) x (jm + 2 gh) x 5 array
extern "C" int so_host_residual(int order, double* out, const double* w, const double* wd, const double* nx, const double* ny,
                                const double* vol, const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz,
                                double cs, double muref, double tref, double s_suth, double k2, double k4, int im, int jm, int wall) {
  const GridDesc g = make_grid(im, jm, gh);
  const SchemeConsts c = make_consts(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4);
#define RUN(O) return wd ? run<1, O>(out, w, wd, nx, ny, vol, volf, g, c, wall != 0) : run<0, O>(out, w, nullptr, nx, ny, vol, volf, g, c, wall != 0)
  switch (order) {
    case 3: RUN(3);
    case 5: RUN(5);
    case 7: RUN(7);
    case 9: RUN(9);
    default: return 1;
  }
#undef RUN
}
