#!/usr/bin/env python3
"""GPU probe: split of the Jacobian -> CSR path (block assembly, row counts + scan, fill) with CUDA events. Scratch tool."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import broadcast_b200 as bb
from broadcast_b200 import cases
from broadcast_b200.resident import Block, jacobian_hybrid, _p


def ev():
    return torch.cuda.Event(enable_timing=True)


sizes = [tuple(int(x) for x in s.split("x")) for s in (sys.argv[1:] or ["1024x2048"])]
for im, jm in sizes:
    c = cases.make_bl_case(im, jm, f_geom=bb.f_geom)
    blk = Block(c); blk.apply_bcs()
    blocks = torch.zeros((29, 5, 5, jm, im), dtype=torch.float64, device=blk.device)
    H = jacobian_hybrid(blk, blocks=blocks)
    H.to_csr(divide_by_vol=True)
    torch.cuda.synchronize()
    out = {"im": im, "jm": jm}
    e = [ev() for _ in range(6)]
    e[0].record()
    H = jacobian_hybrid(blk, blocks=blocks)
    e[1].record()
    dev = blk.device
    n = 5 * im * jm
    indptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    counts = torch.empty(n + 1, dtype=torch.int32, device=dev)
    bsum = torch.empty(n // 2048 + 2, dtype=torch.int64, device=dev)
    region = np.asarray(H.region, dtype=np.int32)
    ns = len(H.strips)
    PP = ctypes.c_void_p * max(ns, 1)
    sj = PP(*[t[0].data_ptr() for t in H.strips]); si = PP(*[t[1].data_ptr() for t in H.strips]); sk = PP(*[t[2].data_ptr() for t in H.strips])
    slen = (ctypes.c_longlong * max(ns, 1))(*[t[0].numel() for t in H.strips])
    srect = np.asarray(H.strip_rects, dtype=np.int32).reshape(-1)
    VP = ctypes.c_void_p
    e[2].record()
    blk.call("bcd_hybrid_csr_indptr", _p(indptr), _p(counts), _p(bsum), _p(H.blocks), region.ctypes.data_as(VP), ns, sj, si, slen,
             ctypes.c_double(2e-16), blk.gh, blk.im, blk.jm, blk._stream())
    e[3].record()
    nnz = int(indptr[-1].item())
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    data = torch.empty(nnz, dtype=torch.float64, device=dev)
    e[4].record()
    blk.call("bcd_hybrid_csr_fill", _p(indices), _p(data), _p(counts), _p(indptr), _p(H.blocks), region.ctypes.data_as(VP), ns,
             srect.ctypes.data_as(VP), sj, si, sk, slen, ctypes.c_double(2e-16), _p(blk.vol), blk.gh, blk.im, blk.jm, blk._stream())
    e[5].record()
    torch.cuda.synchronize()
    out.update(assembly_ms=e[0].elapsed_time(e[1]), count_scan_ms=e[2].elapsed_time(e[3]), fill_ms=e[4].elapsed_time(e[5]), nnz=nnz,
               blocks_GB=blocks.numel() * 8 / 1e9, csr_GB=nnz * 12 / 1e9)
    # counting fused into the assembly kernel
    del indices, data
    f0, f1, f2 = ev(), ev(), ev()
    f0.record()
    H2 = jacobian_hybrid(blk, blocks=blocks, count_thresh=2e-16)
    f1.record()
    ip2, idx2, dat2 = H2.to_csr(divide_by_vol=True)
    f2.record()
    torch.cuda.synchronize()
    out.update(assembly_counted_ms=f0.elapsed_time(f1), to_csr_counted_ms=f1.elapsed_time(f2), same_indptr=bool(torch.equal(ip2, indptr)))
    del H2, ip2, idx2, dat2
    indices = data = None
    print(json.dumps(out), flush=True)
    del blk, blocks, H, indices, data; torch.cuda.empty_cache()
