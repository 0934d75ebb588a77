#!/usr/bin/env python3
"""GPU probe: hybrid Jacobian -> CSR, torch-op path (to_csr) vs hand-written kernels (to_csr_device)."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import broadcast_b200 as bb
from broadcast_b200 import cases
from broadcast_b200.resident import Block, jacobian_hybrid
for im, jm in [tuple(int(x) for x in s.split("x")) for s in (sys.argv[1:] or ["500x150", "630x300", "2048x512"])]:
    c = cases.make_bl_case(im, jm, f_geom=bb.f_geom)
    blk = Block(c); blk.apply_bcs()
    Hj = jacobian_hybrid(blk)
    def timed(fn, n=3):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
    out = {"im": im, "jm": jm, "csr_kernels_ms": timed(lambda: Hj.to_csr_device(divide_by_vol=True))}
    if im * jm <= 200000:
        out["csr_torch_ops_ms"] = timed(lambda: Hj.to_csr_torch(divide_by_vol=True), 1)
    out["nnz"] = int(Hj.to_csr_device()[0][-1].item())
    print(json.dumps(out), flush=True)
    del Hj, blk; torch.cuda.empty_cache()
