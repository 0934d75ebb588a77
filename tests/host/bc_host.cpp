// Host build of the product's boundary-fill line functions (broadcast_b200/csrc/bc.cuh), passive and tangent arithmetic.  TEST
// INFRASTRUCTURE: the header is compiled unchanged with the CUDA qualifiers defined away, the one-thread-per-line-cell kernels are
// replaced by a loop over the line, and the result is checked against the reference routines run on oracle/_ref
// (tests/test_bc_host_cpu.py).  On the GPU the same functions run in k_bc_* (csrc/kernels.cu).
#define __device__
#define __forceinline__ inline
#include <cstdint>
#include "../../broadcast_b200/csrc/bc.cuh"

using namespace bcast;

// which: 0 adiabatic wall, 1 non-reflecting (tab = wbd(lm,5)), 2 inlet (tab = field(lm,gh,5)), 3 extrapolation, 4 isothermal wall
// (p0 = twall, p1 = rgaz), 5 symmetry, 6 antisymmetry, 7 pressure outlet (p0 = pext, p1 = noref), 8 blowing-profile wall (tab = velprof,
// tabd = velprofd, p0 = gamd), 9 temperature-profile wall (tab, tabd, p0 = gamd, p1 = rgaz, p2 = rgazd)
template <int N>
static int run(int which, double* w, double* wd, const BcLine& b, const GridDesc& g, double gam, double p0, double p1, double p2,
               const double* nx, const double* ny, const double* tab, const double* tabd, int lm) {
  const StateRW<N> s{w, wd, g};
  for (int l = 0; l < b.lmax; ++l) {
    switch (which) {
      case 0: bc_wall_viscous_adia_line<N>(s, b, gam, l); break;
      case 1: bc_no_reflexion_line<N>(s, b, tab, lm, nx, ny, gam, l); break;
      case 2: bc_supandsubinlet_line<N>(s, b, tab, lm, nx, ny, gam, l); break;
      case 3: bc_extrapolate_o2_line<N>(s, b, l); break;
      case 4: bc_wall_viscous_iso_line<N>(s, b, p0, gam, p1, l); break;
      case 5: bc_symmetry_line<N, false>(s, b, nx, ny, l); break;
      case 6: bc_symmetry_line<N, true>(s, b, nx, ny, l); break;
      case 7: bc_pressure_line<N>(s, b, p0, p1 != 0.0, gam, nx, ny, l); break;
      case 8: bc_wall_profile_line<N, true>(s, b, tab, tabd, gam, p0, 1.0, 0.0, l); break;
      case 9: bc_wall_profile_line<N, false>(s, b, tab, tabd, gam, p0, p1, p2, l); break;
      default: return 2;
    }
  }
  return 0;
}

extern "C" int bc_host_fill(int which, double* w, double* wd /* null = primal */, const char* loc, const int32_t* interf, double gam, double p0,
                            double p1, double p2, const double* nx, const double* ny, const double* tab, const double* tabd, int lm, int gh,
                            int im, int jm) {
  const GridDesc g = make_grid(im, jm, gh);
  BcLine b;
  int it[4] = {interf[0], interf[1], interf[2], interf[3]};
  if (!decode_interface(loc, it, b)) return 1;
  return wd ? run<1>(which, w, wd, b, g, gam, p0, p1, p2, nx, ny, tab, tabd, lm) : run<0>(which, w, nullptr, b, g, gam, p0, p1, p2, nx, ny, tab, tabd, lm);
}
