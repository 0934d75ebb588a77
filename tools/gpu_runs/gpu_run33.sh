mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r33_pytest.log; cat gpurun_out/r33_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r33_int_launches.csv python - > gpurun_out/r33_int.log 2>&1 <<'PY'
import sys, ctypes; sys.path.insert(0, '.')
import torch, broadcast_b200 as bb
from broadcast_b200 import cases
from broadcast_b200.resident import Block, _p
im, jm = 4096, 1024
c = cases.make_bl_case(im, jm, f_geom=bb.f_geom)
blk = Block(c); blk.apply_bcs()
blocks = torch.zeros((29, 5, 5, jm, im), dtype=torch.float64, device=blk.device)
for _ in range(2):
    blk.call("bcd_jacobian_interior", _p(blocks), _p(blk.w), _p(blk.nx), _p(blk.ny), _p(blk.vol), _p(blk.volf), blk.gh, *blk._phys, im, jm, ctypes.c_void_p(None), ctypes.c_void_p(None), blk._stream())
torch.cuda.synchronize()
PY
python tools/ncu_summary.py gpurun_out/r33_int_launches.csv 8 | cut -c1-150
