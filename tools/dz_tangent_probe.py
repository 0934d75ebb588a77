#!/usr/bin/env python3
"""GPU probe of the fused tangent-of-Dz pass (k_dz_tangent): device time per launch (CUDA events, after warm-up) of both operators in
one pass, of each alone, at the given grid (default C5 8192 x 2048), against the algorithmic traffic (15 + 5 doubles read, 5 written per
operator).  Usage: dz_tangent_probe.py [im jm [reps]]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import broadcast_b200 as bb
from broadcast_b200 import cases
from broadcast_b200.resident import Block

im, jm = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 2048)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
c = cases.make_bl_case(im, jm, f_geom=bb.f_geom, with_w=True)
blk = Block(c)
blk.apply_bcs()
gen = torch.Generator(device=blk.device).manual_seed(4)
wd = torch.randn(blk.w.shape, dtype=torch.float64, device=blk.device, generator=gen)
wd0 = torch.randn(blk.w.shape, dtype=torch.float64, device=blk.device, generator=gen) * blk.w.abs().amax(dim=(1, 2), keepdim=True)
o1, o2 = torch.zeros_like(blk.w), torch.zeros_like(blk.w)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for label, a, b, nbytes in (("both", o1, o2, 240), ("dz", o1, None, 200), ("dz2", None, o2, 160)):
    ms = timed(lambda: blk.dz_tangent(wd0, wd, a, b))
    print(json.dumps({"kernel": "k_dz_tangent", "pass": label, "grid": [im, jm], "ms": round(ms, 4), "algorithmic_bytes_per_cell": nbytes,
                      "GBps": round(nbytes * im * jm / ms / 1e6, 1), "Gcells_per_s": round(im * jm / ms / 1e6, 3)}))
