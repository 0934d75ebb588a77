#!/usr/bin/env python3
"""GPU probe: one device Newton correction (resident.newton_step) at a named config size, timed, against scipy's sparse LU of the
same matrix on the host (the role MUMPS plays in the reference).  Scratch tool."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import broadcast_b200 as bb
from broadcast_b200 import cases
from broadcast_b200.resident import Block, jacobian_hybrid, gmres, newton_step

im, jm = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "500x150").split("x"))
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
with_lu = len(sys.argv) > 3 and sys.argv[3] == "lu"
c = cases.make_bl_case(im, jm, f_geom=bb.f_geom, with_w=True)
blk = Block(c); blk.apply_bcs(); blk.residual()
gh = c.gh
ip, idx, dat = jacobian_hybrid(blk).to_csr()
n = 5 * im * jm
# diagonal of the un-relaxed Jacobian -> pseudo-time term coefdiag = cflm1 * vol (BROADCAST_npz.py:1067)
rows = torch.repeat_interleave(torch.arange(n, device=dat.device), (ip[1:] - ip[:-1]))
diag = dat[rows == idx.long()].abs()
vol = torch.as_tensor(np.ascontiguousarray(c.vol[gh:-gh, gh:-gh].T), device=blk.device)
cflm1 = scale * float(diag.median()) / float(vol.median())
coef = (cflm1 * vol).contiguous()
out = {"im": im, "jm": jm, "n": n, "cflm1": cflm1}
w0 = blk.w.clone()
for rep in range(2):
    blk.w.copy_(w0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dw, info = newton_step(blk, coefdiag=coef, restart=40, maxit=20000, rtol=1e-8)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
out.update(newton_step_s=dt, **info)
# the solve alone
blk.w.copy_(w0); blk.apply_bcs(); blk.residual()
H = jacobian_hybrid(blk, coefdiag=coef)
ip, idx, dat = H.to_csr()
rhs = blk.res[:, gh:gh + jm, gh:gh + im].permute(2, 1, 0).contiguous().view(-1)
torch.cuda.synchronize(); t0 = time.perf_counter()
x, info2 = gmres(ip, idx, dat, rhs, restart=40, maxit=20000, rtol=1e-8)
torch.cuda.synchronize(); out["gmres_s"] = time.perf_counter() - t0
out["gmres_matvecs"] = info2["matvecs"]
if with_lu:
    import scipy.sparse as sp, scipy.sparse.linalg as spla
    A = sp.csr_matrix((dat.cpu().numpy(), idx.cpu().numpy(), ip.cpu().numpy()), shape=(n, n)).tocsc()
    b = rhs.cpu().numpy()
    t0 = time.perf_counter(); xr = spla.spsolve(A, b); out["scipy_lu_s"] = time.perf_counter() - t0
    out["err_vs_lu"] = float(np.abs(x.cpu().numpy() - xr).max() / np.abs(xr).max())
print(json.dumps(out))
