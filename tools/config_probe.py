#!/usr/bin/env python3
"""GPU probe over the reference's own configurations (BASELINE.json configs C1-C4): device times of the residual step, the
Jacobian assembly, its CSR conversion and the Dz / Dz2 operators, next to the CPU reference (oracle/_ref) where it is quick."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import broadcast_b200 as bb
from broadcast_b200 import cases
from broadcast_b200.resident import Block, jacobian_hybrid, jacobian_coo, dz_coo

def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def cpu_step_ms(c):
    from oracle import refmods
    R = refmods.make(fast=True)
    w = c.w.copy(order="F"); res = c.zeros_state()
    cases.apply_bcs(c, w, R["f_bnd"]); R["f_sch"].flux_num_dnc5_2d(res, w, *c.scheme_args())
    t0 = time.perf_counter()
    for _ in range(3):
        cases.apply_bcs(c, w, R["f_bnd"]); R["f_sch"].flux_num_dnc5_2d(res, w, *c.scheme_args())
    t_res = (time.perf_counter() - t0) / 3 * 1e3
    wd = c.zeros_state(); resd = c.zeros_state()
    t0 = time.perf_counter()
    for n in range(2):
        wd *= 0.0
        R["f_misc"].testvector(wd, n, 2 * n, 3, c.gh, c.im, c.jm)
        cases.apply_bcs_lin(c, w, wd, R["f_bnd"], R["f_lin"])
        R["f_lin"].flux_num_dnc5_2d_d(res, resd, w, wd, *c.scheme_args())
    t_col = (time.perf_counter() - t0) / 2 * 1e3
    return t_res, t_col

cfgs = [("C1 boundary layer 500x150 (card_bl2d_fv_npz.py)", "bl", 500, 150), ("C2/C3 cylinder O-mesh 630x300 (card_cyl2d.py, biglobal_cyl.py)", "cyl", 630, 300)]
for name, kind, im, jm in cfgs:
    c = cases.make_bl_case(im, jm, f_geom=bb.f_geom) if kind == "bl" else cases.make_cyl_case(im, jm, f_geom=bb.f_geom, f_bnd=bb.f_bnd)
    blk = Block(c); blk.apply_bcs()
    out = {"config": name, "cells": im * jm}
    out["gpu_step_ms"] = timed(lambda: blk.step(), 20)
    out["gpu_jacobian_hybrid_ms"] = timed(lambda: jacobian_hybrid(blk), 3)
    Hj = jacobian_hybrid(blk)
    out["gpu_csr_divvol_ms"] = timed(lambda: Hj.to_csr(divide_by_vol=True), 3)
    out["nnz"] = int(Hj.to_csr()[0][-1].item())
    out["gpu_reference_colour_loop_ms"] = timed(lambda: jacobian_coo(blk), 1)
    if kind == "bl":
        out["gpu_dz_dz2_colour_loop_ms (C4)"] = timed(lambda: dz_coo(blk), 1)
    t_res, t_col = cpu_step_ms(c)
    out["cpu_step_ms_1core"] = t_res
    out["cpu_jacobian_s_1core_extrapolated"] = 245 * t_col / 1e3
    out["cpu_note"] = "oracle/_ref (reference Fortran machine-translated to C, gcc -O3), one core; Jacobian = 245 x (seed + linearised fills + tangent), 2 passes timed, scatter / COO->CSR not included"
    print(json.dumps(out), flush=True)
    del Hj, blk; torch.cuda.empty_cache()
