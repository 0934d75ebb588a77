// Host build of the product's fused residual tile algorithm (broadcast_b200/csrc/residual_fast.cuh is host+device code):
// TEST INFRASTRUCTURE so that the tile indexing and the re-associated face formulas can be checked against the oracle
// on a machine without a GPU.  The CTA is emulated phase by phase (a loop over the 288 thread ids per phase, barriers
// = loop boundaries).  Not part of the product; the product runs the same phase functions inside k_residual_fast.
#include <cstdint>
#include <vector>
#include "../../broadcast_b200/csrc/residual_fast.cuh"

using namespace bcast;

// isothermal-wall scheme variant (flux_num_dnc5_iso.F90): wall context of the calling thread for the next rf_host_residual calls
extern "C" void rf_host_wall_iso(int on, double twall) { current_wall_iso() = WallIso{on, twall}; }

extern "C" int rf_host_residual(double* res, const double* w, const double* nx, const double* ny, const double* vol, const double* volf,
                                int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref,
                                double tref, double s_suth, double k2, double k4, int im, int jm, int wall, int ioff, int img, int edges, int staged) {
  if (gh != rf::H) return 1;
  GridDesc g = make_grid(im, jm, gh);
  if (img > 0) {
    g.ioff = ioff;
    g.img = img;
    g.edges = edges;
  }
  std::vector<double> sm(rf::NSM_BULK);
  std::vector<double> r((size_t)rf::NT * 5);
  const SchemeConsts sc = make_consts(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4);
  rf::TileCtx t(g, sc);
  t.wsm = sm.data();
  t.sm = sm.data() + rf::WBUF;
  t.sqgr = std::sqrt(gam * rgaz);
  t.wall = wall != 0;
  t.w = w; t.nx = nx; t.ny = ny; t.vol = vol; t.volf = volf; t.res = res;
  for (int by = 0; by < (jm + rf::OJ - 1) / rf::OJ; ++by)
    for (int bx = 0; bx < (im + rf::OI - 1) / rf::OI; ++bx) {
      t.i0 = 1 + bx * rf::OI;
      t.j0 = 1 + by * rf::OJ;
      std::fill(sm.begin(), sm.end(), std::nan(""));   // reading an unwritten shared entry must show up
      if (staged == 2) {   // k_residual_fast_bulk: w box, vol / volf boxes and the node rows by the op list of rf::bulk_op
        t.met = sm.data() + rf::O_MET;
        t.volbox = sm.data() + rf::O_VOLBOX;
        t.nsh = rf::node_shift_mask(g, t.i0, t.j0);
        double* met = sm.data();
        for (int op = 0; op < rf::NBULK; ++op) {
          const rf::BulkOp o = rf::bulk_op(g, t.i0, t.j0, op);
          if (o.kind < 0 || (o.x & 1)) return 8;   // TMA: the first byte of a box must be 16-byte aligned (B200 faults otherwise)
          if (o.dst % 16) return 9;                // ... and its shared-memory destination 128-byte aligned
          if (o.kind == 0) {
            for (int e = 0; e < 5; ++e)
              for (int b = 0; b < rf::PJ; ++b)
                for (int a = 0; a < rf::PI; ++a) {
                  const int si = o.x + a, sj = o.y + b;
                  t.wsm[e * rf::NC + a + b * rf::PI] = (si < g.ni() && sj < g.nj()) ? w[e * g.sc + si + (long long)sj * g.ldc] : 0.0;
                }
          } else if (o.kind == 1) {
            for (int b = 0; b < rf::MV_H; ++b)
              for (int a = 0; a < rf::MV_W; ++a) {
                const int si = o.x + a, sj = o.y + b;
                met[o.dst + a + b * rf::MV_W] = (si < g.ni() && sj < g.nj()) ? vol[si + (long long)sj * g.ldc] : 0.0;
              }
          } else if (o.kind == 2) {
            for (int e = 0; e < 2; ++e)
              for (int b = 0; b < rf::MF_H; ++b)
                for (int a = 0; a < rf::MF_W; ++a) {
                  const int si = o.x + a, sj = o.y + b;
                  met[o.dst + (e * rf::MF_H + b) * rf::MF_W + a] = (si < g.ni() && sj < g.nj()) ? volf[e * g.sc + si + (long long)sj * g.ldc] : 0.0;
                }
          } else {   // node box: the array seen as (nj + 1) rows of 2 ldn elements, zero fill outside
            const double* src = o.kind == 3 ? nx : ny;
            const long long rowlen = 2LL * g.ldn, nrows = g.nj() + 1;
            for (int b = 0; b < rf::MN_HALF; ++b)
              for (int a = 0; a < rf::MN_SLOT; ++a) {
                const long long xx = o.x + a, yy = o.y + b;
                met[o.dst + b * rf::MN_SLOT + a] = (xx < rowlen && yy < nrows) ? src[yy * rowlen + xx] : 0.0;
              }
          }
        }
        for (int tid = 0; tid < rf::NT; ++tid) rf::phase0<true>(t, tid);
        for (int tid = 0; tid < rf::NT; ++tid) rf::phase1(t, tid, rf::sensor_geom_sm(t, tid, 0), rf::sensor_geom_sm(t, tid, 1));
        if (t.has_ghost_sensor())
          for (int tid = 0; tid < rf::NT; ++tid) rf::phase1b(t, tid);
        for (int tid = 0; tid < rf::NT; ++tid) rf::phase2(t, tid, rf::geom_iface_sm(t, tid));
        for (int tid = 0; tid < rf::NT; ++tid) {
          double (&rr)[5] = *reinterpret_cast<double(*)[5]>(&r[(size_t)tid * 5]);
          rf::balance_i(t, tid, rr);
          rf::phase_rj(t, tid);
        }
        for (int tid = 0; tid < rf::NT; ++tid) rf::phase3(t, tid, rf::geom_jface_sm(t, tid));
        for (int tid = 0; tid < rf::NT; ++tid) {
          const double (&rr)[5] = *reinterpret_cast<double(*)[5]>(&r[(size_t)tid * 5]);
          rf::balance_j_store(t, tid, rr);
        }
        continue;
      }
      if (staged) {   // what the TMA load of k_residual_fast_tma delivers: the (PI, PJ, 5) box of w, zero fill outside the array
        for (int e = 0; e < 5; ++e)
          for (int b = 0; b < rf::PJ; ++b)
            for (int a = 0; a < rf::PI; ++a) {
              const int si = t.i0 - 1 + a, sj = t.j0 - 1 + b;   // storage coordinates of cell (i0-3+a, j0-3+b)
              t.wsm[e * rf::NC + a + b * rf::PI] = (si < g.ni() && sj < g.nj()) ? w[e * g.sc + si + (long long)sj * g.ldc] : 0.0;
            }
        for (int tid = 0; tid < rf::NT; ++tid) rf::phase0<true>(t, tid);
      } else {
        for (int tid = 0; tid < rf::NT; ++tid) rf::phase0<false>(t, tid);
      }
      for (int tid = 0; tid < rf::NT; ++tid) rf::phase1(t, tid, rf::prefetch_sensor(t, tid, 0), rf::prefetch_sensor(t, tid, 1));
      if (t.has_ghost_sensor())
        for (int tid = 0; tid < rf::NT; ++tid) rf::phase1b(t, tid);
      for (int tid = 0; tid < rf::NT; ++tid) rf::phase2(t, tid, rf::prefetch_iface(t, tid));
      for (int tid = 0; tid < rf::NT; ++tid) {
        double (&rr)[5] = *reinterpret_cast<double(*)[5]>(&r[(size_t)tid * 5]);
        rf::balance_i(t, tid, rr);
        rf::phase_rj(t, tid);
      }
      for (int tid = 0; tid < rf::NT; ++tid) rf::phase3(t, tid, rf::prefetch_jface(t, tid));
      for (int tid = 0; tid < rf::NT; ++tid) {
        const double (&rr)[5] = *reinterpret_cast<double(*)[5]>(&r[(size_t)tid * 5]);
        rf::balance_j_store(t, tid, rr);
      }
    }
  return 0;
}
