// Device-side description of one structured block and the accessors the scheme templates use.
//
// Layout in HBM (structure of arrays, i fastest -- the reference's Fortran order):
//   cell arrays   (im+2gh)   x (jm+2gh)   [x planes]   leading dimension ldc, plane stride sc
//   node arrays   (im+2gh+1) x (jm+2gh+1) [x planes]   leading dimension ldn, plane stride sn
// Fortran index (i,j) (lower bound 1-gh) <-> linear (i-1+gh) + (j-1+gh)*ld.
#pragma once
#include "dual.cuh"
#include <type_traits>

namespace bcast {

// read-only load (LDG.NC on the device, plain load on the host build used by the CPU-side tests)
BC_HD double BC_LDG(const double* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

struct GridDesc {
  int im, jm, gh;
  int ldc, ldn;
  long long sc, sn;
  // i-slab of a larger block: global index of local cell i is i + ioff, the block has img columns;
  // edges bit 0 / 1: the Ilo / Ihi side is a slab-internal edge whose gh halo columns hold real neighbour data
  // (gradients are then computed there instead of extrapolated).  Whole block: ioff = 0, img = im, edges = 0.
  int ioff, img, edges;
  // window of a larger block in j as well (the boundary-strip sub-blocks of the Jacobian assembly, jacobian.cu): global row of
  // local row j is j + joff, the block has jmg rows.  Only the colouring seeds and the row / column numbers of the COO scatter use
  // global indices; a j-cut has no special treatment in the kernels (the sub-blocks keep a margin wider than the stencil).
  int joff, jmg;
  BC_HD int glo() const { return (edges & 1) ? 0 : 1; }        // first / last column with a computed gradient
  BC_HD int ghi() const { return (edges & 2) ? im + 1 : im; }
  BC_HD long long cidx(int i, int j) const { return (long long)(i - 1 + gh) + (long long)(j - 1 + gh) * ldc; }
  BC_HD long long nidx(int i, int j) const { return (long long)(i - 1 + gh) + (long long)(j - 1 + gh) * ldn; }
  BC_HD int ni() const { return im + 2 * gh; }
  BC_HD int nj() const { return jm + 2 * gh; }
};

inline GridDesc make_grid(int im, int jm, int gh) {
  GridDesc g;
  g.im = im;
  g.jm = jm;
  g.gh = gh;
  g.ldc = im + 2 * gh;
  g.ldn = im + 2 * gh + 1;
  g.sc = (long long)g.ldc * (jm + 2 * gh);
  g.sn = (long long)g.ldn * (jm + 2 * gh + 1);
  g.ioff = 0;
  g.img = im;
  g.edges = 0;
  g.joff = 0;
  g.jmg = jm;
  return g;
}

// number of primitive planes: u, v, w, T, p, mu, htot
constexpr int NPRIM = 7;
enum { PR_U = 0, PR_V = 1, PR_W = 2, PR_T = 3, PR_P = 4, PR_MU = 5, PR_H = 6 };
// gradient planes: gradu(.,1), gradu(.,2), gradv(.,1), gradv(.,2)
constexpr int NGRAD = 4;

struct FieldPtrs {
  const double* w;      // 5 planes
  const double* prim;   // NPRIM planes
  const double* grad;   // NGRAD planes
  const double* nx;     // 2 planes (node layout)
  const double* ny;     // 2 planes
  const double* vol;    // 1 plane
  const double* volf;   // 2 planes (cell layout)
  // tangents: direction-major [n][plane]
  const double* wd;
  const double* primd;
  const double* gradd;
};

template <int N>
using TanOf = std::conditional_t<N == 0, Zero, Tan<(N == 0 ? 1 : N)>>;

// Accessor over global arrays; every cell carries the same tangent type (none if N == 0).
template <int N>
struct GlobalAcc {
  using DT = TanOf<N>;
  FieldPtrs f;
  long long c, n;  // linear indices of the base cell in cell / node layout
  int ldc, ldn;
  long long sc, sn;

  BC_HD GlobalAcc(const FieldPtrs& f_, const GridDesc& g, int i, int j)
      : f(f_), c(g.cidx(i, j)), n(g.nidx(i, j)), ldc(g.ldc), ldn(g.ldn), sc(g.sc), sn(g.sn) {}

  BC_HD Var<DT> load(const double* v, const double* d, int plane, int nplanes, long long k) const {
    Var<DT> r;
    r.v = BC_LDG(v + plane * sc + k);
    if constexpr (N > 0) {
#pragma unroll
      for (int q = 0; q < N; ++q) r.d.d[q] = BC_LDG(d + (long long)(q * nplanes + plane) * sc + k);
    }
    return r;
  }
  template <int OI, int OJ> BC_HD long long ck() const { return c + OI + (long long)OJ * ldc; }
  template <int OI, int OJ> BC_HD long long nk() const { return n + OI + (long long)OJ * ldn; }

  template <int OI, int OJ> BC_HD Var<DT> W(int e) const { return load(f.w, f.wd, e, 5, ck<OI, OJ>()); }
  template <int OI, int OJ> BC_HD Var<DT> U() const { return load(f.prim, f.primd, PR_U, NPRIM, ck<OI, OJ>()); }
  template <int OI, int OJ> BC_HD Var<DT> V() const { return load(f.prim, f.primd, PR_V, NPRIM, ck<OI, OJ>()); }
  template <int OI, int OJ> BC_HD Var<DT> Wz() const { return load(f.prim, f.primd, PR_W, NPRIM, ck<OI, OJ>()); }
  template <int OI, int OJ> BC_HD Var<DT> T() const { return load(f.prim, f.primd, PR_T, NPRIM, ck<OI, OJ>()); }
  template <int OI, int OJ> BC_HD Var<DT> P() const { return load(f.prim, f.primd, PR_P, NPRIM, ck<OI, OJ>()); }
  template <int OI, int OJ> BC_HD Var<DT> Mu() const { return load(f.prim, f.primd, PR_MU, NPRIM, ck<OI, OJ>()); }
  template <int OI, int OJ> BC_HD Var<DT> H() const { return load(f.prim, f.primd, PR_H, NPRIM, ck<OI, OJ>()); }
  template <int OI, int OJ> BC_HD Var<DT> GU(int cc) const { return load(f.grad, f.gradd, cc, NGRAD, ck<OI, OJ>()); }
  template <int OI, int OJ> BC_HD Var<DT> GV(int cc) const { return load(f.grad, f.gradd, 2 + cc, NGRAD, ck<OI, OJ>()); }
  template <int OI, int OJ> BC_HD auto GR() const {
    struct R { Var<DT> u0, u1, v0, v1; };
    return R{GU<OI, OJ>(0), GU<OI, OJ>(1), GV<OI, OJ>(0), GV<OI, OJ>(1)};
  }
  template <int OI, int OJ> BC_HD auto SENS() const;   // defined in scheme.cuh (needs sens_from_grad)
  template <int OI, int OJ> BC_HD double NX(int k) const { return BC_LDG(f.nx + k * sn + nk<OI, OJ>()); }
  template <int OI, int OJ> BC_HD double NY(int k) const { return BC_LDG(f.ny + k * sn + nk<OI, OJ>()); }
  template <int OI, int OJ> BC_HD double VOL() const { return BC_LDG(f.vol + ck<OI, OJ>()); }
  template <int OI, int OJ> BC_HD double VOLF(int k) const { return BC_LDG(f.volf + k * sc + ck<OI, OJ>()); }
};

}  // namespace bcast
