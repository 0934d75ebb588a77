#!/usr/bin/env python3
"""GPU probe: host<->device copy rates that bound the e2e number (contiguous vs pitched, one direction vs both)."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from broadcast_b200 import _lib
lib = _lib.lib()
nj, ni = 2054, 8198
n = 5 * nj * ni
h_in = torch.empty(n, dtype=torch.float64).pin_memory(); h_in.fill_(1.0)
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda"); d_out = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(fn, rep=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(rep): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / rep
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
out = {"bytes": n * 8}
out["h2d_GBs"] = n * 8 / timeit(h2d) / 1e9
out["d2h_GBs"] = n * 8 / timeit(d2h) / 1e9
t = timeit(both); out["duplex_ms"] = t * 1e3; out["duplex_GBs_each"] = n * 8 / t / 1e9
for nslab in (4, 8, 16):
    w = ni // nslab
    def pitched():
        for k in range(nslab):
            st = (s1, s2)[k % 2]
            with torch.cuda.stream(st):
                sp = ctypes.c_void_p(st.cuda_stream)
                lib.bcd_memcpy2d(ctypes.c_void_p(d_in.data_ptr() + k * w * 5 * nj * 8), ctypes.c_longlong(w * 8), ctypes.c_void_p(h_in.data_ptr() + k * w * 8),
                                 ctypes.c_longlong(ni * 8), ctypes.c_longlong(w * 8), ctypes.c_longlong(5 * nj), 1, sp)
    out[f"h2d_pitched_{nslab}slabs_GBs"] = nslab * w * 5 * nj * 8 / timeit(pitched) / 1e9
for nslab in (8,):
    w = ni // nslab
    def d2h_pitched():
        for k in range(nslab):
            st = (s1, s2)[k % 2]
            with torch.cuda.stream(st):
                sp = ctypes.c_void_p(st.cuda_stream)
                lib.bcd_memcpy2d(ctypes.c_void_p(h_out.data_ptr() + k * w * 8), ctypes.c_longlong(ni * 8), ctypes.c_void_p(d_out.data_ptr() + k * w * 5 * nj * 8),
                                 ctypes.c_longlong(w * 8), ctypes.c_longlong(w * 8), ctypes.c_longlong(5 * nj), 2, sp)
    out[f"d2h_pitched_{nslab}slabs_GBs"] = nslab * w * 5 * nj * 8 / timeit(d2h_pitched) / 1e9
    def duplex_pitched():
        for k in range(nslab):
            lib.bcd_memcpy2d(ctypes.c_void_p(d_in.data_ptr() + k * w * 5 * nj * 8), ctypes.c_longlong(w * 8), ctypes.c_void_p(h_in.data_ptr() + k * w * 8),
                             ctypes.c_longlong(ni * 8), ctypes.c_longlong(w * 8), ctypes.c_longlong(5 * nj), 1, ctypes.c_void_p(s1.cuda_stream))
            lib.bcd_memcpy2d(ctypes.c_void_p(h_out.data_ptr() + k * w * 8), ctypes.c_longlong(ni * 8), ctypes.c_void_p(d_out.data_ptr() + k * w * 5 * nj * 8),
                             ctypes.c_longlong(w * 8), ctypes.c_longlong(w * 8), ctypes.c_longlong(5 * nj), 2, ctypes.c_void_p(s2.cuda_stream))
    t = timeit(duplex_pitched); out["duplex_pitched_ms"] = t * 1e3
print(json.dumps(out))
