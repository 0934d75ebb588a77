#!/usr/bin/env python3
"""Multi-GPU probe (torchrun): host-link bandwidth of pinned copies with 1 .. N ranks active at once, H2D alone, D2H alone, both.
Explains what bounds the end-to-end (host buffers in, residual out) number at N > 1.  Scratch tool."""
import json, os, sys, time
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
MB = 256
h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
d_out = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
flag = torch.zeros(1, device="cuda")


def run(active, mode, reps=6):
    dist.all_reduce(flag)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rank < active:
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt if rank < active else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_dir = MB / 1024 * reps / float(t[0])
    return per_dir   # GB/s per rank and direction (slowest rank)


out = {}
run(world, "both", 2)
for active in [a for a in (1, 2, 4, 8) if a <= world]:
    for mode in ("h2d", "d2h", "both"):
        out[f"{mode}_{active}"] = round(run(active, mode), 2)
if rank == 0:
    print(json.dumps({"world": world, "GBps_per_rank_per_direction": out}))
dist.destroy_process_group()
