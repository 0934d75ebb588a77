// Forward-mode AD scalar types for the BROADCAST hot path (device + host).
//
// Var<Zero>    : passive value (plain double arithmetic, no tangent is ever computed)
// Var<Tan<N>>  : value + N tangent directions (vector forward mode)
//
// Mixed expressions Var<Zero> (op) Var<Tan<N>> resolve at COMPILE time, so a kernel in which only
// one stencil cell carries a tangent pays only for the terms that depend on that cell (used by the
// direct block-Jacobian kernels, jacobian.cu).
//
// Non-smooth intrinsics follow the conventions of the reference's Tapenade 3.16 tangent code so that
// branch choices are identical (reference srcfv/tangent/flux_num_dnc5_d.f90):
//   abs  : x >= 0 ? xd : -xd                                  (:945-951)
//   max  : max(a,b) takes b iff a < b                         (:1054-1074, :1117-1123)
//   sqrt : tangent forced to 0 where the argument is 0        (:923-927)
//   tanh : (1 - tanh^2) xd                                    (:1025)
//   pow  : x**y, real y: 0 if x <= 0 and (y == 0 or y not integer) else y x^(y-1) xd
//                                                             (srcfv/tangent/bc_wall_viscous_d.f90:131-146)
//   sign : piecewise constant, never differentiated           (srcfv/tangent/bc_no_reflexion_d.f90)
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define BC_HD __host__ __device__ __forceinline__
#else
#define BC_HD inline
#endif

namespace bcast {

struct Zero {};
template <int N>
struct Tan {
  double d[N];
};

template <class D>
struct Var {
  double v;
  D d;
};
using PVar = Var<Zero>;
template <int N>
using DVar = Var<Tan<N>>;

// ---- tangent-part algebra ------------------------------------------------------------------
BC_HD Zero t_add(Zero, Zero) { return {}; }
template <int N> BC_HD Tan<N> t_add(Tan<N> a, Zero) { return a; }
template <int N> BC_HD Tan<N> t_add(Zero, Tan<N> b) { return b; }
template <int N> BC_HD Tan<N> t_add(Tan<N> a, Tan<N> b) {
  Tan<N> r;
#pragma unroll
  for (int n = 0; n < N; ++n) r.d[n] = a.d[n] + b.d[n];
  return r;
}
BC_HD Zero t_sub(Zero, Zero) { return {}; }
template <int N> BC_HD Tan<N> t_sub(Tan<N> a, Zero) { return a; }
template <int N> BC_HD Tan<N> t_sub(Zero, Tan<N> b) {
  Tan<N> r;
#pragma unroll
  for (int n = 0; n < N; ++n) r.d[n] = -b.d[n];
  return r;
}
template <int N> BC_HD Tan<N> t_sub(Tan<N> a, Tan<N> b) {
  Tan<N> r;
#pragma unroll
  for (int n = 0; n < N; ++n) r.d[n] = a.d[n] - b.d[n];
  return r;
}
BC_HD Zero t_scale(double, Zero) { return {}; }
template <int N> BC_HD Tan<N> t_scale(double s, Tan<N> a) {
  Tan<N> r;
#pragma unroll
  for (int n = 0; n < N; ++n) r.d[n] = s * a.d[n];
  return r;
}
BC_HD Zero t_neg(Zero) { return {}; }
template <int N> BC_HD Tan<N> t_neg(Tan<N> a) {
  Tan<N> r;
#pragma unroll
  for (int n = 0; n < N; ++n) r.d[n] = -a.d[n];
  return r;
}
// s1*a + s2*b
BC_HD Zero t_lin2(double, Zero, double, Zero) { return {}; }
template <int N> BC_HD Tan<N> t_lin2(double s1, Tan<N> a, double, Zero) { return t_scale(s1, a); }
template <int N> BC_HD Tan<N> t_lin2(double, Zero, double s2, Tan<N> b) { return t_scale(s2, b); }
template <int N> BC_HD Tan<N> t_lin2(double s1, Tan<N> a, double s2, Tan<N> b) {
  Tan<N> r;
#pragma unroll
  for (int n = 0; n < N; ++n) r.d[n] = s1 * a.d[n] + s2 * b.d[n];
  return r;
}
// runtime selection between two tangent parts (result type = join)
BC_HD Zero t_sel(bool, Zero, Zero) { return {}; }
template <int N> BC_HD Tan<N> t_sel(bool c, Tan<N> a, Zero) {
  Tan<N> r;
#pragma unroll
  for (int n = 0; n < N; ++n) r.d[n] = c ? a.d[n] : 0.0;
  return r;
}
template <int N> BC_HD Tan<N> t_sel(bool c, Zero, Tan<N> b) {
  Tan<N> r;
#pragma unroll
  for (int n = 0; n < N; ++n) r.d[n] = c ? 0.0 : b.d[n];
  return r;
}
template <int N> BC_HD Tan<N> t_sel(bool c, Tan<N> a, Tan<N> b) {
  Tan<N> r;
#pragma unroll
  for (int n = 0; n < N; ++n) r.d[n] = c ? a.d[n] : b.d[n];
  return r;
}
template <class D> struct TanTraits;
template <> struct TanTraits<Zero> {
  static constexpr int n = 0;
  BC_HD static Zero zero() { return {}; }
};
template <int N> struct TanTraits<Tan<N>> {
  static constexpr int n = N;
  BC_HD static Tan<N> zero() {
    Tan<N> r;
#pragma unroll
    for (int k = 0; k < N; ++k) r.d[k] = 0.0;
    return r;
  }
};

// promotion of a result to a wider tangent type
template <class D> BC_HD D t_promote(D a, D*) { return a; }
template <int N> BC_HD Tan<N> t_promote(Zero, Tan<N>*) { return TanTraits<Tan<N>>::zero(); }
template <class DT, class D> BC_HD Var<DT> promote(Var<D> a) {
  return Var<DT>{a.v, t_promote(a.d, (DT*)nullptr)};
}

BC_HD PVar cst(double v) { return PVar{v, {}}; }

// ---- arithmetic ----------------------------------------------------------------------------
template <class A, class B> BC_HD auto operator+(Var<A> a, Var<B> b) { return Var<decltype(t_add(a.d, b.d))>{a.v + b.v, t_add(a.d, b.d)}; }
template <class A, class B> BC_HD auto operator-(Var<A> a, Var<B> b) { return Var<decltype(t_sub(a.d, b.d))>{a.v - b.v, t_sub(a.d, b.d)}; }
template <class A, class B> BC_HD auto operator*(Var<A> a, Var<B> b) {
  return Var<decltype(t_lin2(b.v, a.d, a.v, b.d))>{a.v * b.v, t_lin2(b.v, a.d, a.v, b.d)};
}
template <class A, class B> BC_HD auto operator/(Var<A> a, Var<B> b) {
  const double q = a.v / b.v;
  // (ad - q*bd)/b   (Tapenade quotient form)
  auto num = t_sub(a.d, t_scale(q, b.d));
  return Var<decltype(num)>{q, t_scale(1.0 / b.v, num)};
}
template <class A> BC_HD Var<A> operator+(Var<A> a, double b) { return Var<A>{a.v + b, a.d}; }
template <class A> BC_HD Var<A> operator+(double b, Var<A> a) { return Var<A>{b + a.v, a.d}; }
template <class A> BC_HD Var<A> operator-(Var<A> a, double b) { return Var<A>{a.v - b, a.d}; }
template <class A> BC_HD Var<A> operator-(double b, Var<A> a) { return Var<A>{b - a.v, t_neg(a.d)}; }
template <class A> BC_HD Var<A> operator*(Var<A> a, double b) { return Var<A>{a.v * b, t_scale(b, a.d)}; }
template <class A> BC_HD Var<A> operator*(double b, Var<A> a) { return Var<A>{b * a.v, t_scale(b, a.d)}; }
template <class A> BC_HD Var<A> operator/(Var<A> a, double b) { return Var<A>{a.v / b, t_scale(1.0 / b, a.d)}; }
template <class A> BC_HD Var<A> operator/(double b, Var<A> a) {
  const double q = b / a.v;
  return Var<A>{q, t_scale(-q / a.v, a.d)};
}
template <class A> BC_HD Var<A> operator-(Var<A> a) { return Var<A>{-a.v, t_neg(a.d)}; }

// ---- intrinsics ----------------------------------------------------------------------------
template <class A> BC_HD Var<A> sqrt(Var<A> a) {
  const double s = ::sqrt(a.v);
  return Var<A>{s, t_scale(a.v == 0.0 ? 0.0 : 1.0 / (2.0 * s), a.d)};
}
template <class A> BC_HD Var<A> fabs(Var<A> a) { return Var<A>{::fabs(a.v), t_scale(a.v >= 0.0 ? 1.0 : -1.0, a.d)}; }
template <class A> BC_HD Var<A> tanh(Var<A> a) {
  const double t = ::tanh(a.v);
  return Var<A>{t, t_scale(1.0 - t * t, a.d)};
}
template <class A, class B> BC_HD auto fmax(Var<A> a, Var<B> b) {
  const bool takeb = a.v < b.v;
  return Var<decltype(t_sel(takeb, b.d, a.d))>{takeb ? b.v : a.v, t_sel(takeb, b.d, a.d)};
}
template <class A> BC_HD Var<A> fmax(double a, Var<A> b) {  // max(ZERO, x)
  const bool takeb = a < b.v;
  return Var<A>{takeb ? b.v : a, t_scale(takeb ? 1.0 : 0.0, b.d)};
}
template <class A> BC_HD Var<A> pow(Var<A> x, double y) {
  const double p = ::pow(x.v, y);
  double fac;
  if (x.v <= 0.0 && (y == 0.0 || y != (double)(int)y))
    fac = 0.0;
  else
    fac = y * ::pow(x.v, y - 1.0);
  return Var<A>{p, t_scale(fac, x.d)};
}
// x**y with BOTH active (Tapenade, e.g. tangent/bc_wall_blow_profile_d.f90:139-147): the exponent contributes x^y log(x) yd for x > 0 only
template <class A> BC_HD Var<A> pow(Var<A> x, Var<A> y) {
  const double p = ::pow(x.v, y.v);
  double fac;
  if (x.v <= 0.0 && (y.v == 0.0 || y.v != (double)(int)y.v))
    fac = 0.0;
  else
    fac = y.v * ::pow(x.v, y.v - 1.0);
  Var<A> r{p, t_scale(fac, x.d)};
  if (x.v > 0.0) r.d = t_add(r.d, t_scale(p * ::log(x.v), y.d));
  return r;
}
// Fortran SIGN(a,b) with passive result
BC_HD double fsign(double a, double b) { return ::copysign(::fabs(a), b); }

}  // namespace bcast
