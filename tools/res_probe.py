#!/usr/bin/env python3
"""GPU probe: fused residual kernel vs the generic (reference-shaped) kernels at several sizes: time and max relative difference."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import broadcast_b200 as bb
from broadcast_b200 import cases
from broadcast_b200.resident import Block

def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

sizes = [tuple(int(x) for x in s.split("x")) for s in (sys.argv[1:] or ["500x150", "2048x512", "8192x2048"])]
for im, jm in sizes:
    c = cases.make_bl_case(im, jm, f_geom=bb.f_geom)
    blk = Block(c); blk.apply_bcs()
    gh = c.gh
    a = blk.residual().clone()
    b = blk.residual(generic=True).clone()
    ai, bi = a[:, gh:-gh, gh:-gh], b[:, gh:-gh, gh:-gh]
    scale = bi.abs().amax(dim=(1, 2))
    err = ((ai - bi).abs().amax(dim=(1, 2)) / scale).cpu().numpy().tolist()
    out = {"im": im, "jm": jm, "fused_ms": timed(lambda: blk.residual()), "generic_ms": timed(lambda: blk.residual(generic=True), 3),
           "rel_diff_fused_vs_generic": err, "nan": bool(torch.isnan(ai).any().item())}
    out["GBs_136"] = 136.0 * im * jm / (out["fused_ms"] * 1e-3) / 1e9
    for name, v in (("march_ms", 5), ("tma_ms", 2), ("tile_v1_ms", 3)):
        c2 = blk.residual(variant=v).clone()
        out[name] = timed(lambda: blk.residual(variant=v))
        out[name + "_maxdiff_vs_default"] = float(((c2 - a)[:, gh:-gh, gh:-gh].abs().amax(dim=(1, 2)) / scale).max().item())
    print(json.dumps(out), flush=True)
    del blk; torch.cuda.empty_cache()
