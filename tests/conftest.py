import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """f2py-shaped modules over oracle/_ref (the reference Fortran machine-translated to C)."""
    from oracle import refmods
    if not refmods.available():
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import build_ref
        if not build_ref.build(verbose=False):
            pytest.skip("oracle/_ref is not built and /root/reference is absent")
    return refmods.make()


@pytest.fixture(scope="session")
def gpu():
    """product drop-in modules; fails loudly if the CUDA library or device is missing"""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import broadcast_b200 as bb
    assert bb._lib.device_count() > 0
    return dict(f_sch=bb.f_sch, f_lin=bb.f_lin, f_bnd=bb.f_bnd, f_geom=bb.f_geom, f_norm=bb.f_norm, f_misc=bb.f_misc, f_dz=bb.f_dz, f_lindz=bb.f_lindz, f_init=bb.f_init)
