"""GPU tests of the multi-GPU step: the peer-store halo exchange (csrc/halo.cu), the CUDA-graph step,
i-slabs of the i-periodic O-mesh, and -- when the box has two or more GPUs -- the real one-process-per-GPU path under torchrun:
slab-sharded residual over NVLink == the single-GPU residual BIT FOR BIT (order-independent checksum), colour sharding on two
physical GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _slabs(g, world, poison=True):
    import torch
    from broadcast_b200 import sharding
    from broadcast_b200.resident import Block
    out = []
    for r in range(world):
        sl, desc = sharding.slab_of(g, r, world)
        b = Block(sl, slab=desc)
        if poison:
            if desc[2] & 1:
                b.w[:, :, :g.gh] = float("nan")
            if desc[2] & 2:
                b.w[:, :, -g.gh:] = float("nan")
        out.append(b)
    return out


@pytest.mark.parametrize("world", [2, 3])
def test_peer_halo_equals_copy_exchange(gpu, world):
    """three exchanges in a row (both mailbox parities, changing data) through k_halo_push / k_halo_unpack, every slab on its own
    stream, against plain device copies"""
    import torch
    from broadcast_b200 import sharding
    from broadcast_b200.resident import local_halo_exchange
    g = H.make_case("bl", 66, 28, gpu, with_w=True)
    A, B = _slabs(g, world), _slabs(g, world)
    halos = sharding.PeerHalo.local([b.w for b in A], g.gh)
    streams = [torch.cuda.Stream() for _ in range(world)]
    for it in range(3):
        for a, b in zip(A, B):
            a.w[:, :, g.gh:-g.gh] *= 1.0 + 0.01 * (it + 1)
            b.w[:, :, g.gh:-g.gh] *= 1.0 + 0.01 * (it + 1)
        torch.cuda.synchronize()
        for h, a, st in zip(halos, A, streams):
            with torch.cuda.stream(st):
                h(a.w)
        local_halo_exchange(B)
        torch.cuda.synchronize()
        for h, a, b in zip(halos, A, B):
            assert h.error() == 0
            assert torch.equal(a.w.nan_to_num(nan=-7.0), b.w.nan_to_num(nan=-7.0)), it
    for h in halos:
        h.close()


@pytest.mark.parametrize("world", [2, 3])
def test_periodic_slabs_match_single_block(gpu, world):
    """O-mesh (i-periodic) in i-slabs: j-side fills, then the exchange that replaces the join across the cut (first <-> last slab),
    then the residual == the single-block residual, bit for bit"""
    import torch
    from broadcast_b200 import sharding
    from broadcast_b200.resident import Block
    im, jm = 84, 30
    g = H.make_case("cyl", im, jm, gpu, with_w=True)
    G = Block(g)
    G.apply_bcs()
    resG = G.residual().clone()
    blocks = _slabs(g, world)
    halos = sharding.PeerHalo.local([b.w for b in blocks], g.gh, periodic=True)
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for b, h, st in zip(blocks, halos, streams):
        with torch.cuda.stream(st):
            b.apply_bcs()
            h(b.w)
            b.residual()
    torch.cuda.synchronize()
    gh = g.gh
    for r, b in enumerate(blocks):
        lo, hi = sharding.slab_range(im, r, world)
        assert halos[r].error() == 0 and not torch.isnan(b.w).any()
        assert torch.equal(b.res[:, gh:gh + jm, gh:gh + b.im], resG[:, gh:gh + jm, gh + lo - 1:gh + hi]), r
    for h in halos:
        h.close()


def test_step_graph_replays_the_step(gpu):
    """[fills + residual] captured once (bcd_graph_begin / _end) and replayed == the same calls issued one by one; the graph
    follows the state (pointers are captured, not values)"""
    import torch
    from broadcast_b200 import sharding
    from broadcast_b200.resident import Block
    g = H.make_case("bl", 96, 40, gpu, with_w=True)
    blk = Block(g)
    sg = sharding.StepGraph(blk)
    for it in range(2):
        blk.w[:, :, g.gh:-g.gh] *= 1.0 + 0.02 * (it + 1)
        w0 = blk.w.clone()
        sg()
        r1, w1 = blk.res.clone(), blk.w.clone()
        blk.w.copy_(w0)
        blk.res.zero_()
        blk.apply_bcs()
        blk.residual()
        assert torch.equal(blk.w, w1) and torch.equal(blk.res, r1)
    sg.close()


def _torchrun(nproc, *args, timeout=600):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "mp_slab_check.py"), *args]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


@pytest.mark.parametrize("kind", ["bl", "cyl"])
def test_two_gpus_torchrun_slabs_bit_identical(gpu, kind):
    """two processes, two GPUs: PeerHalo over NVLink + StepGraph; the all-reduced checksum of the owned residual cells equals the
    checksum of the single-GPU residual computed by rank 0 (wrap-around int64 sum of bit patterns: order independent)"""
    rec = _torchrun(2, kind)
    assert rec["halo_error"] == 0
    assert rec["checksum_sharded"] == rec["checksum_single"], rec
    assert rec["nccl_checksum"] == rec["checksum_single"], rec


def test_two_gpus_torchrun_colour_sharding(gpu):
    """colour sharding under real torchrun on two physical GPUs: merged filtered COO == the unsharded device loop"""
    rec = _torchrun(2, "colours")
    assert rec["nnz_merged"] == rec["nnz_full"] and rec["max_abs_diff"] == 0.0, rec
