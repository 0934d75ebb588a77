"""world_size-2 (and 3) gloo tests of the i-slab host logic on CPU: slab extraction, halo exchange, global numbering.
The compute on each rank is done by the checker (oracle/_ref) -- this file tests plumbing, not kernels: the slab
results after one halo exchange must reproduce the single-block result wherever a slab has complete stencil data."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from broadcast_b200 import cases, sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, im, jm, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import refmods
        R = refmods.make()
        g = cases.make_bl_case(im, jm, f_geom=R["f_geom"])
        # reference: single block
        wg, rg = H.residual_sequence(R, g)
        sl, (ioff, img, edges) = sharding.slab_of(g, rank, world)
        gh = g.gh
        lo, hi = sharding.slab_range(im, rank, world)
        assert ioff == lo - 1 and img == im and sl.im == hi - lo + 1
        # each rank starts from ITS columns only: poison the halo, then exchange
        w = sl.w.copy(order="F")
        if edges & 1:
            w[:gh] = np.nan
        if edges & 2:
            w[-gh:] = np.nan
        t = torch.from_numpy(np.ascontiguousarray(w.T))            # (5, nj, ni) image of the Fortran array
        halo = sharding.HaloExchange(gh, rank, world)
        halo(t)
        w = np.asfortranarray(t.numpy().T)
        assert not np.isnan(w).any()
        assert np.array_equal(w[gh:-gh, gh:-gh], g.w[lo - 1 + gh:hi + gh, gh:-gh])   # owned cells
        assert np.array_equal(w[:gh, gh:-gh], g.w[lo - 1:lo - 1 + gh, gh:-gh])       # left halo = neighbour's cells (or ghosts)
        cases.apply_bcs(sl, w, R["f_bnd"])
        res = sl.zeros_state()
        R["f_sch"].flux_num_dnc5_2d(res, w, *sl.scheme_args())
        # rows whose stencil (gh + 1 cells for the extrapolated-gradient layer) stays inside real data agree with the global run
        a = gh + 1 if edges & 1 else 0
        b = sl.im - (gh + 1 if edges & 2 else 0)
        mine = res[gh + a:gh + b, gh:-gh]
        ref = rg[lo - 1 + gh + a:lo - 1 + gh + b, gh:-gh]
        err = np.abs(mine - ref).max() / np.abs(rg).max()
        # boundary tables are slices of the global ones
        cnt = torch.tensor([float(sl.im)])
        dist.all_reduce(cnt)
        out[rank] = (float(err), float(cnt[0]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_and_halo_exchange_gloo(world, ref):
    im, jm = 45, 20
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), im, jm, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        err, total = out[r]
        assert total == im
        assert err < 1e-13, (r, err)


def test_slab_partition_and_row_gather():
    for im, world in [(8192, 8), (45, 3), (500, 7)]:
        rng = [sharding.slab_range(im, r, world) for r in range(world)]
        assert rng[0][0] == 1 and rng[-1][1] == im
        assert all(rng[k][1] + 1 == rng[k + 1][0] for k in range(world - 1))
        assert max(b - a for a, b in rng) - min(b - a for a, b in rng) <= 1
    import scipy.sparse as sp
    A = sp.random(60, 60, density=0.2, format="csr", random_state=0)
    parts = []
    for a, b in [(0, 25), (25, 40), (40, 60)]:
        B = A[a:b]
        parts.append((B.indptr, B.indices, B.data))
    ip, idx, dat = sharding.gather_row_blocks(parts)
    C = sp.csr_matrix((dat, idx, ip), shape=A.shape)
    assert (C != A).nnz == 0


# ---- colour sharding (small grids and the i-periodic O-mesh) ------------------------------------------------
def test_colour_ranges_partition_the_passes():
    for world in (1, 2, 3, 4, 8, 49, 60):
        got = []
        for r in range(world):
            c0, c1 = sharding.colour_range(49, r, world)
            assert 0 <= c0 <= c1 <= 49
            got += list(range(c0, c1))
        assert got == list(range(49))
        sizes = [sharding.colour_range(49, r, world)[1] - sharding.colour_range(49, r, world)[0] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def _colour_worker(rank, world, port, im, jm, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import refmods
        R = refmods.make()
        c = H.make_case("cyl", im, jm, R, with_w=True)          # periodic in i: cannot be slab-sharded
        w, _ = H.residual_sequence(R, c)
        s = 2 * c.gh + 1
        c0, c1 = sharding.colour_range(s * s, rank, world)
        mine = [(m, l, k) for m in range(5) for l in range(s) for k in range(s) if c0 <= l * s + k < c1]
        jac, ia, ja = H.jacobian_sequence(R, c, w, mine, None)   # the checker computes this rank's passes
        parts = [None] * world
        dist.all_gather_object(parts, (jac, ia, ja))
        if rank == 0:
            n = 5 * im * jm
            A = sharding.merge_colour_shards(parts, n)
            jf, iaf, jaf = H.jacobian_sequence(R, c, w, None, None)
            B = H.coo_to_dict(jf, iaf, jaf)
            B.resize((n, n))
            D = (A - B).tocoo()
            out.put((A.nnz, B.nnz, float(np.abs(D.data).max()) if D.nnz else 0.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_colour_sharded_jacobian_merges_to_the_full_one_gloo(world):
    """every rank runs a contiguous range of the 49 colour passes on the whole (i-periodic) grid; rank 0 merges the filtered
    COO lists: identical to the single-process colour loop (each matrix entry comes from exactly one colour)"""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_colour_worker, args=(r, world, port, 21, 14, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    nnzA, nnzB, err = out.get(timeout=10)
    assert nnzA == nnzB and err == 0.0


# ---- i-slabs of the i-periodic O-mesh: the join across the cut becomes the exchange between the last and the first slab ----------
def _periodic_worker(rank, world, port, im, jm, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import refmods
        R = refmods.make()
        g = H.make_case("cyl", im, jm, R, with_w=True)
        wg, rg = H.residual_sequence(R, g)                 # single block: fills (noref, wall, join twice) + residual
        sl, (ioff, img, edges) = sharding.slab_of(g, rank, world)
        gh = g.gh
        lo, hi = sharding.slab_range(im, rank, world)
        assert edges == 3 and sl.slab_periodic and not sl.periodic_i and all(b[0] != "jn" for b in sl.bcs)
        w = sl.w.copy(order="F")
        w[:gh] = np.nan
        w[-gh:] = np.nan
        cases.apply_bcs(sl, w, R["f_bnd"])                 # j-side fills of the slab's own columns first ...
        t = torch.from_numpy(np.ascontiguousarray(w.T))
        sharding.HaloExchange(gh, rank, world, periodic=True)(t)   # ... then the exchange (all rows, ghost rows included)
        w = np.asfortranarray(t.numpy().T)
        assert not np.isnan(w).any()
        # the slab's padded state is the window of the single block's filled state (the cut included)
        ref = np.concatenate([wg[-2 * gh:-gh], wg[gh:-gh], wg[gh:2 * gh]], axis=0)[lo - 1:hi + 2 * gh]
        assert np.array_equal(w, ref), np.abs(w - ref).max()
        res = sl.zeros_state()
        R["f_sch"].flux_num_dnc5_2d(res, w, *sl.scheme_args())
        a, b = gh + 1, sl.im - (gh + 1)                    # away from the slab edges (the checker extrapolates gradients there)
        err = np.abs(res[gh + a:gh + b, gh:-gh] - rg[lo - 1 + gh + a:lo - 1 + gh + b, gh:-gh]).max() / np.abs(rg).max()
        out[rank] = float(err)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_periodic_slabs_and_exchange_gloo(world, ref):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_periodic_worker, args=(world, _free_port(), 42, 16, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        assert out[r] < 1e-13, (r, out[r])


def test_row_windows_reproduce_the_block_residual_on_the_oracle(ref):
    """sharding.row_window_of (the windows of resident.RowStreamedBlock): every window run as a block of its own on the ORACLE -- fills
    of its clipped boundary list, `_nowall` scheme away from the wall -- gives, inside a margin of two rows, exactly the rows of the
    whole-block residual (and a margin of zero rows does not: the cut is visible in the outermost rows)"""
    im, jm = 36, 64
    c = H.make_case("bl", im, jm, ref, with_w=True)
    gh = c.gh
    w, res = H.residual_sequence(ref, c)
    M = 2
    seen = np.zeros(jm, dtype=bool)
    for a, b in [(1, 20), (21, 45), (46, 64)]:
        la, lb = max(1, a - M), min(jm, b + M)
        wc = sharding.row_window_of(c, la, lb)
        assert wc.jm == lb - la + 1 and ("nowall" in wc.scheme) == (la != 1)
        kinds = [bc[0] for bc in wc.bcs]
        assert ("wall" in kinds) == (la == 1) and ("noref" in kinds) == (lb == jm) and "inflow" in kinds and "outflow" in kinds
        # the window's state = the parent's rows (ghost rows of a cut side hold real rows), as the host copy delivers them
        wc.w[:] = c.w[:, la - 1:lb + 2 * gh]
        ww, rw = H.residual_sequence(ref, wc, wc.scheme)
        own = slice(gh + (a - la), gh + (a - la) + (b - a + 1))
        assert np.array_equal(rw[gh:-gh, own], res[gh:-gh, gh + a - 1:gh + b])
        seen[a - 1:b] = True
        if la != 1:   # without the margin the first row of a cut window differs
            assert not np.array_equal(rw[gh:-gh, gh], res[gh:-gh, gh + la - 1])
    assert seen.all()


def test_slab_bounds_and_tapered_widths():
    from broadcast_b200.resident import StreamedBlock
    b = StreamedBlock.tapered_bounds(8192, 8)
    assert b[0] == 0 and b[-1] == 8192 and len(b) == 9
    w = np.diff(b)
    assert all(x % 32 == 0 for x in w) and w[0] < w[3] and w[-1] < w[4] and list(w[:4]) == sorted(w[:4])
    for k in range(8):
        lo, hi = sharding.slab_range(8192, k, 8, b)
        assert (lo, hi) == (b[k] + 1, b[k + 1])
    with pytest.raises(ValueError):
        sharding.slab_range(100, 0, 2, [0, 60, 90])
