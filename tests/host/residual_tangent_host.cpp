// Host build of the product's fused TANGENT tile algorithm: residual_fast.cuh compiled in dual-number arithmetic (BCAST_RF_DUAL,
// 32 x 3 tiles = the strips of the Jacobian assembly).  TEST INFRASTRUCTURE: the CTA is emulated phase by phase and checked
// against the reference's Tapenade tangent run on oracle/_ref.  The product runs the same phase functions in k_tangent_tile.
#define BCAST_RF_DUAL 1
#define BCAST_RF_OJ 3
#define BCAST_RF_NS rfd
#include <vector>
#include "../../broadcast_b200/csrc/residual_fast.cuh"

using namespace bcast;

extern "C" int rfd_host_tangent(double* resd, const double* w, const double* wd, const double* nx, const double* ny, const double* vol,
                                const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                                double muref, double tref, double s_suth, double k2, double k4, int im, int jm, int wall, int ri0, int ri1,
                                int rj0, int rj1) {
  if (gh != rfd::H) return 1;
  const GridDesc g = make_grid(im, jm, gh);
  const SchemeConsts sc = make_consts(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4);
  std::vector<rfd::real> sm(rfd::NSM);
  std::vector<rfd::real> r((size_t)rfd::NT * 5);
  rfd::TileCtx t(g, sc);
  t.wsm = sm.data();
  t.sm = sm.data() + rfd::WBUF;
  t.sqgr = std::sqrt(gam * rgaz);
  t.wall = wall != 0;
  t.w = w; t.wd = wd; t.nx = nx; t.ny = ny; t.vol = vol; t.volf = volf; t.res = resd;
  t.i1 = ri1; t.j1 = rj1;
  std::vector<unsigned char> flags(rfd::NC);
  t.flags = flags.data();   // face skipping on, as in the kernel
  for (int j0 = rj0; j0 <= rj1; j0 += rfd::OJ)
    for (int i0 = ri0; i0 <= ri1; i0 += rfd::OI) {
      t.i0 = i0;
      t.j0 = j0;
      for (auto& x : sm) x = rfd::Dual{std::nan(""), std::nan("")};
      for (int tid = 0; tid < rfd::NT; ++tid) rfd::phase0<false>(t, tid);
      for (int tid = 0; tid < rfd::NT; ++tid) rfd::phase1(t, tid, rfd::prefetch_sensor(t, tid, 0), rfd::prefetch_sensor(t, tid, 1));
      if (t.has_ghost_sensor())
        for (int tid = 0; tid < rfd::NT; ++tid) rfd::phase1b(t, tid);
      for (int tid = 0; tid < rfd::NT; ++tid) rfd::phase2(t, tid, rfd::prefetch_iface(t, tid));
      for (int tid = 0; tid < rfd::NT; ++tid) {
        rfd::real (&rr)[5] = *reinterpret_cast<rfd::real(*)[5]>(&r[(size_t)tid * 5]);
        rfd::balance_i(t, tid, rr);
        rfd::phase_rj(t, tid);
      }
      for (int tid = 0; tid < rfd::NT; ++tid) rfd::phase3(t, tid, rfd::prefetch_jface(t, tid));
      for (int tid = 0; tid < rfd::NT; ++tid) {
        const rfd::real (&rr)[5] = *reinterpret_cast<rfd::real(*)[5]>(&r[(size_t)tid * 5]);
        rfd::balance_j_store(t, tid, rr);
      }
    }
  return 0;
}
