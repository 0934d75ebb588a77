// Host build of the product's fused tangent-of-Dz tile algorithm (broadcast_b200/csrc/dz_tangent.cuh).  TEST INFRASTRUCTURE: the CTA
// is emulated phase by phase (A: own cells, B: cross halo, C: d/dz rows) and checked against the reference's Tapenade code
// (srcfv/tangentdz/coeffs_5p_dz_d.f90, coeffs_5p_dz2_d.f90) run on oracle/_ref.  On the GPU the same phase functions run in k_dz_tangent.
#include <cmath>
#include <vector>
#include "../../broadcast_b200/csrc/dz_tangent.cuh"

using namespace bcast;

extern "C" int dzt_host(double* out1, double* out2, const double* w, const double* wd0, const double* wd, const double* nx, const double* ny,
                        const double* vol, int gh, double cp, double cv, double prandtl, double gam, double cs, double muref, double tref,
                        double s_suth, int im, int jm) {
  if (gh != 3) return 1;
  dzt::Tile t;
  t.g = make_grid(im, jm, gh);
  t.c = dzt::make_dz_consts(cp, cv, prandtl, gam, cs, muref, tref, s_suth);
  t.w = w; t.wa = wd; t.wb = wd0; t.nx = nx; t.ny = ny; t.vol = vol; t.out1 = out1; t.out2 = out2;
  std::vector<double> sm(dzt::NSM);
  std::vector<dzt::Carry> cells(dzt::NT);
  t.sm = sm.data();
  t.i1 = im; t.j1 = jm;
  for (int j0 = 1; j0 <= jm; j0 += dzt::TJ)
    for (int i0 = 1; i0 <= im; i0 += dzt::TI) {
      t.i0 = i0; t.j0 = j0;
      for (auto& x : sm) x = std::nan("");   // a read of an unstaged cell must show
      for (int tid = 0; tid < dzt::NT; ++tid) {   // the kernel's order: both loads, halo cell, own cell
        long long ka, kb;
        const int sa = dzt::own_cell(t, tid, &ka), sb = dzt::halo_cell(t, tid, &kb);
        const dzt::Raw ra = dzt::load_raw(t, sa >= 0, ka), rb = dzt::load_raw(t, sb >= 0, kb);
        dzt::phase_b(t, sb, rb);
        cells[tid] = dzt::carry_of(dzt::phase_a(t, tid, ra, ka));
      }
      for (int tid = 0; tid < dzt::NT; ++tid) dzt::phase_c(t, tid, cells[tid], dzt::load_metrics(t, tid));
    }
  return 0;
}
