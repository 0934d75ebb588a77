"""k_residual_fast_bulk (csrc/residual_bulk.cu: every input of a tile delivered by TMA / bulk copies, metrics read from shared
memory) against the LDG tile kernel k_residual_fast: the same phase functions on the same numbers, so bit for bit the same residual,
on grids that exercise ragged tiles, both parities of the node leading dimension (shifted rows), the O-mesh, the nowall scheme and
i-slabs.  Parity of the tile kernel itself with the oracle: tests/test_parity_gpu.py, tests/test_configs_gpu.py."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,im,jm", [("bl", 96, 48), ("bl", 70, 21), ("bl", 33, 9), ("bl", 32, 8), ("bl", 301, 77), ("cyl", 45, 19),
                                        ("cyl", 126, 60), ("bl", 300, 70)])
def test_bulk_kernel_bit_identical_to_tile_kernel(gpu, kind, im, jm):
    import torch
    from broadcast_b200.resident import Block
    c = H.make_case(kind, im, jm, gpu, with_w=True)
    blk = Block(c)
    blk.apply_bcs()
    a = blk.residual(variant=4).clone()
    blk.res.fill_(float("nan"))
    b = blk.residual(variant=6).clone()
    gh = c.gh
    assert torch.equal(a[:, gh:-gh, gh:-gh], b[:, gh:-gh, gh:-gh])
    assert torch.isfinite(b[:, gh:-gh, gh:-gh]).all()
    # nowall scheme
    blk.wall = 0
    a = blk.residual(variant=4).clone()
    b = blk.residual(variant=6).clone()
    assert torch.equal(a[:, gh:-gh, gh:-gh], b[:, gh:-gh, gh:-gh])


def test_bulk_kernel_on_slabs(gpu):
    import torch
    from broadcast_b200 import sharding
    from broadcast_b200.resident import Block, local_halo_exchange
    g = H.make_case("bl", 130, 40, gpu, with_w=True)
    blocks = []
    for r in range(3):
        sl, desc = sharding.slab_of(g, r, 3)
        blocks.append(Block(sl, slab=desc))
    local_halo_exchange(blocks)
    gh = g.gh
    for b in blocks:
        b.apply_bcs()
        x = b.residual(variant=4).clone()
        y = b.residual(variant=6).clone()
        assert torch.equal(x[:, gh:-gh, gh:-gh], y[:, gh:-gh, gh:-gh])
