#!/usr/bin/env python3
"""GPU probe: history of resident.newton_loop on a small boundary-layer block. Scratch tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import broadcast_b200 as bb
from broadcast_b200 import cases
from broadcast_b200.resident import Block, newton_loop
im, jm = (int(x) for x in sys.argv[1].split("x"))
for cfl in [float(x) for x in sys.argv[2:]]:
    c = cases.make_bl_case(im, jm, f_geom=bb.f_geom, with_w=True)
    blk = Block(c)
    h = newton_loop(blk, cfl=cfl, nit=8, rtol=1e-9, maxit=6000)
    print("cfl", cfl)
    for it, n, ni, cm, mv in h:
        print("  it", it, "norm", np.array2string(n, precision=3), "1/cfl %.3e" % cm, "matvecs", mv)
