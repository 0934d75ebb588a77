#!/usr/bin/env python3
"""Launch the residual kernel of one variant a few times at one size (for ncu captures): res_one.py IMxJM [variant] [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import broadcast_b200 as bb
from broadcast_b200 import cases
from broadcast_b200.resident import Block
im, jm = (int(x) for x in sys.argv[1].split("x"))
v = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
c = cases.make_bl_case(im, jm, f_geom=bb.f_geom)
blk = Block(c); blk.apply_bcs()
for _ in range(n):
    blk.residual(variant=v)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    blk.residual(variant=v)
e1.record(); torch.cuda.synchronize()
print("variant", v, "ms", e0.elapsed_time(e1) / n)
