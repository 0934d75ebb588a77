"""The boundary strips of the hybrid Jacobian on compact SUB-BLOCKS (csrc/jacobian.cu: bcd_jacobian_strips cuts every strip out of the
block with a margin of gh + 2, clips the boundary list to the window, numbers rows / columns through ioff / joff and runs the colour
loop with up to 16 chains) against the same colour loop on the whole grid (BROADCAST_B200_STRIPS_FULL=1): same slots, same integers,
same values, on the boundary-layer case (inlet / non-reflecting / outflow / wall lists), with a coefdiag, on i-slabs, and with the
isothermal wall + pressure outlet.  The full-grid loop itself is checked against oracle/_ref in tests/test_parity_gpu.py."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _strips(blk, coef, full):
    import torch
    from broadcast_b200.resident import jacobian_hybrid
    if full:
        os.environ["BROADCAST_B200_STRIPS_FULL"] = "1"
    else:
        os.environ.pop("BROADCAST_B200_STRIPS_FULL", None)
    try:
        Hj = jacobian_hybrid(blk, coefdiag=coef, strip_buffers="fresh")
        torch.cuda.synchronize()
        return [(j.clone(), a.clone(), b.clone()) for j, a, b in Hj.strips], Hj.strip_rects
    finally:
        os.environ.pop("BROADCAST_B200_STRIPS_FULL", None)


@pytest.mark.parametrize("im,jm", [(66, 28), (130, 40), (300, 70)])
def test_windowed_strips_equal_full_grid_loop(gpu, im, jm):
    import torch
    from broadcast_b200.resident import Block
    c = H.make_case("bl", im, jm, gpu, with_w=True)
    coef = np.asfortranarray(np.random.default_rng(3).uniform(0.5, 1.5, size=(im, jm)))
    blk = Block(c)
    blk.apply_bcs()
    a, ra = _strips(blk, coef, full=True)
    b, rb = _strips(blk, coef, full=False)
    assert ra == rb and len(a) == 4
    for q, ((ja, iaa, jaa), (jb, iab, jab)) in enumerate(zip(a, b)):
        assert torch.equal(iaa, iab) and torch.equal(jaa, jab), (q, ra[q])
        assert torch.equal(ja, jb), (q, ra[q], (ja - jb).abs().max().item())
    # second assembly on a changed state: the cached graphs follow the data (windows are re-cut every call)
    blk.w[:, c.gh:-c.gh, c.gh:-c.gh] *= 1.003
    blk.apply_bcs()
    a, _ = _strips(blk, coef, full=True)
    b, _ = _strips(blk, coef, full=False)
    for (ja, iaa, jaa), (jb, iab, jab) in zip(a, b):
        assert torch.equal(ja, jb) and torch.equal(iaa, iab) and torch.equal(jaa, jab)


def test_windowed_strips_on_slabs(gpu):
    import torch
    from broadcast_b200 import sharding
    from broadcast_b200.resident import Block, local_halo_exchange
    g = H.make_case("bl", 150, 40, gpu, with_w=True)
    blocks = []
    for r in range(3):
        sl, desc = sharding.slab_of(g, r, 3)
        blocks.append(Block(sl, slab=desc))
    local_halo_exchange(blocks)
    for r, b in enumerate(blocks):
        b.apply_bcs()
        lo, hi = sharding.slab_range(150, r, 3)
        coef = np.asfortranarray(np.random.default_rng(r).uniform(0.5, 1.5, size=(b.im, b.jm)))
        x, rx = _strips(b, coef, full=True)
        y, ry = _strips(b, coef, full=False)
        assert rx == ry
        for (ja, iaa, jaa), (jb, iab, jab) in zip(x, y):
            assert torch.equal(iaa, iab) and torch.equal(jaa, jab) and torch.equal(ja, jb)


def test_banded_assembly_equals_whole_block_csr(gpu):
    """BandedAssembly (the C5-on-one-GPU path: i-bands through one reused band buffer) == the CSR of the whole-block assembly,
    pattern and values, with the division by the cell volume"""
    import torch
    from broadcast_b200.resident import Block, BandedAssembly, jacobian_hybrid
    c = H.make_case("bl", 132, 36, gpu, with_w=True)
    blk = Block(c)
    blk.apply_bcs()
    ip, idx, dat = jacobian_hybrid(blk).to_csr(divide_by_vol=True)
    for nband in (1, 3, 4):
        ba = BandedAssembly(c, nband)
        parts = ba.assemble_csr(blk.w)
        assert len(parts) == nband
        ip2, idx2, dat2 = BandedAssembly.gather(parts)
        assert torch.equal(ip, ip2) and torch.equal(idx, idx2)
        assert (dat - dat2).abs().max().item() <= 1e-13 * dat.abs().max().item()
        # a second assembly on a changed state overwrites the index / value arrays of the first (reuse) and must equal the
        # whole-block CSR of that state; the row counts come from the assembly kernel itself (jacobian_hybrid(count_thresh=...))
        w2 = blk.w.clone()
        blk.w[:, c.gh:-c.gh, c.gh:-c.gh] *= 1.0 + 1e-3 * torch.rand_like(blk.w[:, c.gh:-c.gh, c.gh:-c.gh])
        blk.apply_bcs()
        ipb, idxb, datb = jacobian_hybrid(blk).to_csr(divide_by_vol=True)
        store = [t[1].untyped_storage().data_ptr() for t in parts]
        parts = ba.assemble_csr(blk.w)
        ip3, idx3, dat3 = BandedAssembly.gather(parts)
        assert torch.equal(ipb, ip3) and torch.equal(idxb, idx3)
        assert (datb - dat3).abs().max().item() <= 1e-13 * datb.abs().max().item()
        assert any(t[1].untyped_storage().data_ptr() == s_ for t, s_ in zip(parts, store))
        blk.w.copy_(w2)
