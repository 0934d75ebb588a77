// Host build of the product's j-marching residual algorithm (broadcast_b200/csrc/residual_march.cuh is host+device code):
// TEST INFRASTRUCTURE so that the ring indexing, the step schedule and the face formulas can be checked against the oracle on a
// machine without a GPU.  The CTA is emulated phase by phase (a loop over the 288 thread ids per phase, barriers = loop
// boundaries); what TMA / LDGSTS deliver is copied row by row at the point where the kernel issues it, and the shared arrays start
// as NaN so that reading an entry nobody wrote shows up.  Not part of the product; the product runs the same phase functions
// inside k_residual_march (residual_march.cu).
#include <algorithm>
#include <cmath>
#include <vector>
#include "../../broadcast_b200/csrc/residual_march.cuh"

using namespace bcast;

// what one generation of asynchronous copies delivers: the operations of rm::copy_op, executed the way the copy engine does
// (tensor boxes with zero fill outside the array; node rows as plain runs of `bytes` starting at the aligned source element)
static void issue_rows_host(const rm::MCtx& t, int cq0, int cn, int mq0, int mn) {
  const GridDesc& g = t.g;
  for (int op = 0; op < rm::copy_count(cn, mn); ++op) {
    const rm::CopyOp o = rm::copy_op(t, op, cq0, cn, mq0, mn);
    if (o.kind < 0) continue;
    double* dst = t.sm + o.dst;
    if (o.kind == 3) {
      for (int k = 0; k < o.bytes / 8; ++k) dst[k] = o.src[k];
      continue;
    }
    const int planes = o.kind == 0 ? 5 : (o.kind == 1 ? 1 : 2);
    const double* base = o.kind == 0 ? t.w : (o.kind == 1 ? t.vol : t.volf);
    for (int e = 0; e < planes; ++e)
      for (int a = 0; a < rm::PC; ++a) {
        const int si = o.x + a;
        dst[e * rm::PC + a] = (si < g.ni() && o.y < g.nj()) ? base[e * g.sc + si + (long long)o.y * g.ldc] : 0.0;
      }
  }
}

extern "C" int rm_host_residual(double* res, const double* w, const double* nx, const double* ny, const double* vol, const double* volf,
                                int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref,
                                double tref, double s_suth, double k2, double k4, int im, int jm, int wall, int ioff, int img, int edges,
                                int seglen) {
  if (gh != rm::H || seglen % rm::RB) return 1;
  GridDesc g = make_grid(im, jm, gh);
  if (img > 0) {
    g.ioff = ioff;
    g.img = img;
    g.edges = edges;
  }
  std::vector<double> sm(rm::NSM);
  const SchemeConsts sc = make_consts(cp, cv, prandtl, gam, rgaz, cs, muref, tref, s_suth, k2, k4);
  rm::MCtx t(g, sc);
  t.sm = sm.data();
  t.sqgr = std::sqrt(gam * rgaz);
  t.wall = wall != 0;
  t.w = w; t.nx = nx; t.ny = ny; t.vol = vol; t.volf = volf; t.res = res;
  const int nstrips = (im + rm::W - 1) / rm::W, nseg = (jm + seglen - 1) / seglen;
#define ALL(expr) for (int tid = 0; tid < rm::NT; ++tid) { expr; }
  for (int seg = 0; seg < nseg; ++seg)
    for (int strip = 0; strip < nstrips; ++strip) {
      t.it.i0 = 1 + strip * rm::W;
      t.it.j0 = 1 + seg * seglen;
      t.it.j1 = std::min(jm, (seg + 1) * seglen);
      const int nsteps = (t.it.j1 - t.it.j0 + rm::RB) / rm::RB;
      std::fill(sm.begin(), sm.end(), std::nan(""));
      issue_rows_host(t, 0, rm::PRO_CELL_ROWS, rm::PRO_MET_Q0, rm::PRO_MET_ROWS);
      ALL(rm::prims_rows(t, tid, rm::NT, 0, rm::PRO_CELL_ROWS))
      for (int s = -1; s < nsteps; ++s) {
        const int qJ = 3 + rm::RB * s;
        const bool more = s >= 0 && s + 1 < nsteps;
        // the kernel issues these rows here and consumes them at the end of the step: emulate the asynchronous arrival by writing
        // them right away (a ring slot that is still live would be corrupted and show up as a wrong result)
        if (more) issue_rows_host(t, qJ + 7, rm::RB, qJ + 6, rm::RB);
        int sq0 = qJ + 1, sn = rm::RB, iq0 = qJ + 2, in = rm::RB, jq0 = qJ + 1, jn = rm::RB;
        if (s < 0) { sq0 = 2; sn = 3; iq0 = 1; in = 4; jq0 = 3; jn = 1; }
        ALL(rm::phase_sens_r(t, tid, sq0, sn, iq0, in, jq0, jn))
        if (rm::item_has_ghost_sensor(t, sq0, sn)) ALL(rm::phase_sens_ghost(t, tid, sq0, sn))
        ALL(rm::phase_faces(t, tid, qJ))
        ALL(if (tid < rm::NT_BAL) rm::phase_balance(t, tid, qJ); else if (more) rm::prims_rows(t, tid - rm::NT_BAL, rm::NT - rm::NT_BAL, qJ + 7, rm::RB))
      }
    }
  return 0;
}
