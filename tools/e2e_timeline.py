#!/usr/bin/env python3
"""GPU probe: per-slab timeline of StreamedBlock.step_from_host (H2D done / kernels done / D2H done, ms from the step's start). Scratch tool."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import broadcast_b200 as bb
from broadcast_b200 import cases, _lib
from broadcast_b200.resident import StreamedBlock, _p

im, jm = 8192, 2048
nslab = int(sys.argv[1]) if len(sys.argv) > 1 else 8
taper = len(sys.argv) > 2 and sys.argv[2] == "1"
c = cases.make_bl_case(im, jm, f_geom=bb.f_geom)
bounds = StreamedBlock.tapered_bounds(im, nslab) if taper else None
sb = StreamedBlock(c, nslab=nslab, bounds=bounds)
shape = (5, jm + 2 * c.gh, im + 2 * c.gh)
wp = torch.empty(shape, dtype=torch.float64).pin_memory(); rp = torch.empty(shape, dtype=torch.float64).pin_memory()
wp.copy_(torch.from_numpy(np.ascontiguousarray(c.w.T)))
for _ in range(3):
    sb.step_from_host(wp, rp)
# instrumented copy of step_from_host
E = lambda: torch.cuda.Event(enable_timing=True)
gh, nj, ni = sb.gh, sb.jm + 2 * sb.gh, int(wp.shape[2])
rows = 5 * nj
LL, VP = ctypes.c_longlong, ctypes.c_void_p
main = torch.cuda.current_stream(sb.device)
t0 = E(); t0.record(main)
for s in (sb.s_in, sb.s_k, sb.s_out):
    s.wait_stream(main)
ein, ek, eout = [], [], []
for k, (b, lo, hi) in enumerate(sb.slabs):
    nl = b.im + 2 * gh
    src = wp.data_ptr() + (lo - 1) * 8
    _lib.check(sb.lib.bcd_memcpy2d(_p(b.w), LL(nl * 8), VP(src), LL(ni * 8), LL(nl * 8), LL(rows), 1, VP(sb.s_in.cuda_stream)), "m")
    e = E(); e.record(sb.s_in); ein.append(e)
for k, (b, lo, hi) in enumerate(sb.slabs):
    sb.s_k.wait_event(ein[k])
    with torch.cuda.stream(sb.s_k):
        b.apply_bcs(); b.residual()
    e = E(); e.record(sb.s_k); ek.append(e)
    nl = b.im + 2 * gh
    sb.s_out.wait_event(e)
    dst = rp.data_ptr() + (lo - 1 + gh) * 8
    _lib.check(sb.lib.bcd_memcpy2d(VP(dst), LL(ni * 8), VP(b.res.data_ptr() + gh * 8), LL(nl * 8), LL(b.im * 8), LL(rows), 2, VP(sb.s_out.cuda_stream)), "m")
    e2 = E(); e2.record(sb.s_out); eout.append(e2)
main.wait_stream(sb.s_out); main.synchronize()
out = {"nslab": nslab, "taper": taper, "widths": [b.im for b, _, _ in sb.slabs],
       "h2d_done": [round(t0.elapsed_time(e), 2) for e in ein], "kernels_done": [round(t0.elapsed_time(e), 2) for e in ek],
       "d2h_done": [round(t0.elapsed_time(e), 2) for e in eout]}
print(json.dumps(out))
