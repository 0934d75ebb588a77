// Direct block-Jacobian of the regular interior rows (no colouring): for a fixed column offset
// (DI,DJ), every thread owns one row cell (i,j) and evaluates the tangent of its four face fluxes
// with respect to the five conservative variables of cell (i+DI, j+DJ) ONLY.  All other cells are
// passive at compile time (dual.cuh), so each instantiation contains just the terms that depend on
// that cell: one pass over the stencil emits a whole 5x5 block per row cell, 29 instantiations cover
// the structural stencil of the order-5 scheme (5x5 box + (+-3,0), (0,+-3)).
//
// The result equals what the reference's colour loop attributes to that (row cell, column cell) pair
// for rows whose stencil touches neither ghost cells nor the wall rows (4 <= i <= im-3,
// 4 <= j <= jm-3): jac = -(d residu / d w_col) [+ coefdiag on the diagonal]
// (misc/ComputeJacobian.f90:558-563, srcfv/tangent/flux_num_dnc5_d.f90).
#pragma once
#include "kernels.cuh"
#include "facejac.cuh"

namespace bcast {

constexpr bool iface_dep(int pi, int pj) { return (pj == 0 && pi >= -3 && pi <= 2) || (pi >= -2 && pi <= 1 && pj >= -2 && pj <= 2); }
constexpr bool jface_dep(int pi, int pj) { return iface_dep(pj, pi); }

// accessor with exactly one active cell at compile-time offset (PI,PJ) from the face cell
template <int PI, int PJ>
struct PrunedAcc {
  using DT = Tan<5>;
  GlobalAcc<0> g;
  const Var<Tan<5>>* q;          // seeded conservative state of the active cell
  const CellPrims<Tan<5>>* pp;   // its primitives

  __device__ __forceinline__ PrunedAcc(const FieldPtrs& f, const GridDesc& gd, int i, int j, const Var<Tan<5>>* q_,
                                       const CellPrims<Tan<5>>* pp_)
      : g(f, gd, i, j), q(q_), pp(pp_) {}

#define BC_PRUNED(NAME, FIELD)                                            \
  template <int OI, int OJ> __device__ __forceinline__ auto NAME() const { \
    if constexpr (OI == PI && OJ == PJ)                                   \
      return pp->FIELD;                                                   \
    else                                                                  \
      return g.template NAME<OI, OJ>();                                   \
  }
  BC_PRUNED(U, u)
  BC_PRUNED(V, v)
  BC_PRUNED(Wz, w)
  BC_PRUNED(T, t)
  BC_PRUNED(P, p)
  BC_PRUNED(Mu, mu)
  BC_PRUNED(H, h)
#undef BC_PRUNED
  template <int OI, int OJ> __device__ __forceinline__ auto W(int e) const {
    if constexpr (OI == PI && OJ == PJ)
      return q[e];
    else
      return g.template W<OI, OJ>(e);
  }
  template <int OI, int OJ> __device__ __forceinline__ auto GR() const {
    constexpr int a = PI - OI, b = PJ - OJ;
    constexpr bool dep = (b == 0 && (a == 1 || a == -1 || a == 2 || a == -2)) || (a == 0 && (b == 1 || b == -1 || b == 2 || b == -2));
    if constexpr (dep)
      return cell_gradients<OI, OJ>(*this);
    else
      return g.template GR<OI, OJ>();
  }
  template <int OI, int OJ> __device__ __forceinline__ auto SENS() const { return sens_from_grad(GR<OI, OJ>()); }
  template <int OI, int OJ> __device__ __forceinline__ double NX(int k) const { return g.template NX<OI, OJ>(k); }
  template <int OI, int OJ> __device__ __forceinline__ double NY(int k) const { return g.template NY<OI, OJ>(k); }
  template <int OI, int OJ> __device__ __forceinline__ double VOL() const { return g.template VOL<OI, OJ>(); }
  template <int OI, int OJ> __device__ __forceinline__ double VOLF(int k) const { return g.template VOLF<OI, OJ>(k); }
};

// values layout: V[(e*5 + m) * ncell + cell], cell = (i-1) + (j-1)*im  (one slot = 25 planes of im*jm)
template <int DI, int DJ>
__global__ void __launch_bounds__(128) k_jac_block(GridDesc g, SchemeConsts c, FieldPtrs f, Rect rc, double* __restrict__ V,
                                                   const double* __restrict__ coefdiag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + rc.i0;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + rc.j0;
  if (i > rc.i1 || j > rc.j1) return;
  // active cell (i+DI, j+DJ): identity seeds
  Var<Tan<5>> q[5];
  {
    const long long k = g.cidx(i + DI, j + DJ);
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      q[e].v = __ldg(f.w + e * g.sc + k);
#pragma unroll
      for (int m = 0; m < 5; ++m) q[e].d.d[m] = (e == m) ? 1.0 : 0.0;
    }
  }
  const CellPrims<Tan<5>> pp = cell_prims(q, c);
  // residud = -(hn_i(i+1) - hn_i(i)) - (hn_j(j+1) - hn_j(j))   (balance.F);  jac = -residud (+ coefdiag)
  double t[25];
  {
    Var<Tan<5>> a0[5], a1[5];
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      a0[e] = promote<Tan<5>>(cst(0.0));
      a1[e] = a0[e];
    }
    if constexpr (iface_dep(DI, DJ)) {
      PrunedAcc<DI, DJ> A(f, g, i, j, q, &pp);
      face_flux<0, false, FACE_MAIN>(A, c, a0);
    }
    if constexpr (iface_dep(DI - 1, DJ)) {
      PrunedAcc<DI - 1, DJ> A(f, g, i + 1, j, q, &pp);
      face_flux<0, false, FACE_MAIN>(A, c, a1);
    }
#pragma unroll
    for (int e = 0; e < 5; ++e)
#pragma unroll
      for (int m = 0; m < 5; ++m) t[e * 5 + m] = -(a1[e].d.d[m] - a0[e].d.d[m]);
  }
  {
    Var<Tan<5>> b0[5], b1[5];
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      b0[e] = promote<Tan<5>>(cst(0.0));
      b1[e] = b0[e];
    }
    if constexpr (jface_dep(DI, DJ)) {
      PrunedAcc<DI, DJ> A(f, g, i, j, q, &pp);
      face_flux<1, false, FACE_MAIN>(A, c, b0);
    }
    if constexpr (jface_dep(DI, DJ - 1)) {
      PrunedAcc<DI, DJ - 1> A(f, g, i, j + 1, q, &pp);
      face_flux<1, false, FACE_MAIN>(A, c, b1);
    }
#pragma unroll
    for (int e = 0; e < 5; ++e)
#pragma unroll
      for (int m = 0; m < 5; ++m) t[e * 5 + m] = t[e * 5 + m] - (b1[e].d.d[m] - b0[e].d.d[m]);
  }
  const long long ncell = (long long)g.im * g.jm;
  const long long cell = (long long)(i - 1) + (long long)(j - 1) * g.im;
  double cd = 0.0;
  if (DI == 0 && DJ == 0 && coefdiag) cd = coefdiag[cell];
#pragma unroll
  for (int e = 0; e < 5; ++e) {
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      const double rd = t[e * 5 + m];
      double v = -rd;
      if (DI == 0 && DJ == 0 && e == m && coefdiag) v = cd - rd;
      V[(long long)(e * 5 + m) * ncell + cell] = v;
    }
  }
}

template <int DI, int DJ>
cudaError_t launch_jac_block(const GridDesc& g, const SchemeConsts& c, const FieldPtrs& f, const Rect& rc, double* V, const double* coefdiag,
                             cudaStream_t st) {
  dim3 blk(32, 4), grd((rc.i1 - rc.i0 + 32) / 32, (rc.j1 - rc.j0 + 4) / 4);
  k_jac_block<DI, DJ><<<grd, blk, 0, st>>>(g, c, f, rc, V, coefdiag);
  return cudaGetLastError();
}

// the 29 structural offsets: BCAST_JAC_OFFSETS / JAC_NSLOT in facejac.cuh

typedef cudaError_t (*jac_block_fn)(const GridDesc&, const SchemeConsts&, const FieldPtrs&, const Rect&, double*, const double*, cudaStream_t);
jac_block_fn jac_block_launcher(int di, int dj);

}  // namespace bcast
