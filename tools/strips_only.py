#!/usr/bin/env python3
"""GPU probe: the boundary-strip colour loop alone (for an ncu launch list). Scratch tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import broadcast_b200 as bb
from broadcast_b200 import cases
from broadcast_b200.resident import Block, jacobian_strips
im, jm = (int(x) for x in sys.argv[1].split("x"))
c = cases.make_bl_case(im, jm, f_geom=bb.f_geom)
blk = Block(c); blk.apply_bcs()
gh = c.gh
rects = [(1, im, 1, gh), (1, im, jm - gh + 1, jm), (1, gh, gh + 1, jm - gh), (im - gh + 1, im, gh + 1, jm - gh)]
out = jacobian_strips(blk, rects)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); jacobian_strips(blk, rects, out=out); e1.record(); torch.cuda.synchronize()
print("strips ms", e0.elapsed_time(e1))
