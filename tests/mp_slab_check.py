"""Run under torchrun by tests/test_multigpu_gpu.py (one process per GPU): slab-sharded step over the peer-store halo exchange
and the captured step graph vs the single-GPU step; colour-sharded device colour loop vs the unsharded one.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist


def checksum(res, gh, im, jm):
    own = res[:, gh:gh + jm, gh:gh + im].contiguous()
    return own.view(torch.int64).sum().reshape(1)


def main():
    mode = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import broadcast_b200 as bb
    import helpers as H
    from broadcast_b200 import sharding
    from broadcast_b200.resident import Block, jacobian_coo
    mods = dict(f_geom=bb.f_geom, f_bnd=bb.f_bnd)
    out = {}
    if mode in ("bl", "cyl"):
        im, jm = (192, 64) if mode == "bl" else (168, 48)
        g = H.make_case(mode, im, jm, mods, with_w=True)
        gh = g.gh
        sl, desc = sharding.slab_of(g, rank, world)
        blk = Block(sl, dev, slab=desc)
        periodic = bool(sl.slab_periodic)

        def run(halo):
            blk.upload_state(sl.w)
            blk.w[:, :, :gh] = float("nan")
            blk.w[:, :, -gh:] = float("nan")
            if periodic:
                blk.apply_bcs(); halo(blk.w)
            else:
                halo(blk.w); blk.apply_bcs()
            blk.residual()
            c = checksum(blk.res, gh, blk.im, jm)
            dist.all_reduce(c)
            return int(c.item())
        ph = sharding.PeerHalo(gh, rank, world, blk.w, periodic=periodic)
        for _ in range(3):                       # several exchanges: both mailbox parities
            cs = run(ph)
        out["checksum_sharded"] = cs
        out["nccl_checksum"] = run(sharding.HaloExchange(gh, rank, world, periodic=periodic))
        if not periodic:                          # the captured step graph (exchange + fills + residual)
            blk.upload_state(sl.w)
            sg = sharding.StepGraph(blk, ph)
            blk.upload_state(sl.w)
            sg()
            c = checksum(blk.res, gh, blk.im, jm)
            dist.all_reduce(c)
            out["checksum_graph"] = int(c.item())
            assert out["checksum_graph"] == cs, (out["checksum_graph"], cs)
            sg.close()
        err = torch.tensor([ph.error()], device=dev)
        dist.all_reduce(err)
        out["halo_error"] = int(err.item())
        ph.close()
        if rank == 0:
            G = Block(g, dev)
            G.apply_bcs()
            G.residual()
            out["checksum_single"] = int(checksum(G.res, gh, im, jm).item())
    elif mode == "colours":
        g = H.make_case("cyl", 42, 30, mods, with_w=True)
        blk = Block(g, dev)
        blk.apply_bcs()
        s = 2 * g.gh + 1
        c0, c1 = sharding.colour_range(s * s, rank, world)
        mine = tuple(t.cpu().numpy() for t in jacobian_coo(blk, colours=(c0, c1)))
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            n = 5 * g.im * g.jm
            A = sharding.merge_colour_shards(parts, n)
            B = H.coo_to_dict(*(t.cpu().numpy() for t in jacobian_coo(blk)))
            B.resize((n, n))
            D = (A - B).tocoo()
            out = {"nnz_merged": int(A.nnz), "nnz_full": int(B.nnz), "max_abs_diff": float(np.abs(D.data).max()) if D.nnz else 0.0}
    dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
