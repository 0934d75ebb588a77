#!/usr/bin/env python3
"""GPU probe: hybrid Jacobian of one i-slab (rank 0 of 2) of C5 vs the whole block, per call."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import broadcast_b200 as bb
from broadcast_b200 import cases, sharding
from broadcast_b200.resident import Block, jacobian_hybrid
im, jm = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8192x2048").split("x"))
c = cases.make_bl_case(im, jm, f_geom=bb.f_geom)
for world, rank in ((2, 0), (2, 1), (1, 0)):
    cs, desc = sharding.slab_of(c, rank, world)
    blk = Block(cs, slab=desc if world > 1 else None); blk.apply_bcs()
    blocks = torch.empty((29, 5, 5, cs.jm, cs.im), dtype=torch.float64, device=blk.device)
    cd = torch.zeros((cs.jm, cs.im), dtype=torch.float64, device=blk.device)
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        jacobian_hybrid(blk, coefdiag=cd, blocks=blocks)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"world": world, "rank": rank, "im_local": cs.im, "hybrid_ms_per_call": ts}), flush=True)
    del blk, blocks; torch.cuda.empty_cache()
