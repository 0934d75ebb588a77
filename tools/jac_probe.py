#!/usr/bin/env python3
"""GPU probe: time the Jacobian assembly variants at a few grid sizes (CUDA events). Scratch tool."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import broadcast_b200 as bb
from broadcast_b200 import cases
from broadcast_b200.resident import Block, jacobian_coo, jacobian_hybrid

def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

sizes = [tuple(int(x) for x in s.split("x")) for s in (sys.argv[1:] or ["500x150", "630x300", "2048x512"])]
for im, jm in sizes:
    c = cases.make_bl_case(im, jm, f_geom=bb.f_geom)
    blk = Block(c); blk.apply_bcs()
    out = {"im": im, "jm": jm}
    out["residual_ms"] = timed(lambda: blk.residual(), 10)
    blocks = torch.zeros((29, 5, 5, jm, im), dtype=torch.float64, device=blk.device)
    out["hybrid_ms"] = timed(lambda: jacobian_hybrid(blk, blocks=blocks), 2)
    import ctypes
    from broadcast_b200.resident import _p
    out["interior_ms"] = timed(lambda: blk.call("bcd_jacobian_interior", _p(blocks), _p(blk.w), _p(blk.nx), _p(blk.ny), _p(blk.vol), _p(blk.volf),
                                                 blk.gh, *blk._phys, im, jm, ctypes.c_void_p(None), ctypes.c_void_p(None), blk._stream()), 5)
    out["interior_GBs_5904"] = 5904.0 * im * jm / (out["interior_ms"] * 1e-3) / 1e9
    if 25 * 49 * im * jm < 2**31 and 25 * 49 * im * jm * 16 < 60e9:
        nb = 25 * 49 * im * jm
        bufs = (torch.zeros(nb, dtype=torch.float64, device=blk.device), torch.zeros(nb, dtype=torch.int32, device=blk.device),
                torch.zeros(nb, dtype=torch.int32, device=blk.device))
        out["coo_loop_ms"] = timed(lambda: jacobian_coo(blk, out=bufs), 1)
        del bufs
    out["jac_GBs_5904"] = 5904.0 * im * jm / (out["hybrid_ms"] * 1e-3) / 1e9
    print(json.dumps(out), flush=True)
    del blk, blocks; torch.cuda.empty_cache()
