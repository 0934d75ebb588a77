"""bc_general_2d / bc_general_2d_d (srcfv/borders/bc_general.F90, srcfv/tangent/bc_general_d.f90): the Dirichlet fill from a table the
cards name as the alternative inlet routine (card_bl2d_fv.py:107), on every side of a block and for the ghost depths of the scheme
family, against the reference routines on oracle/_ref -- bit-exact (the fill copies, the tangent zeroes)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", [3, 5, 9])
@pytest.mark.parametrize("loc", ["Ilo", "Ihi", "Jlo", "Jhi"])
def test_bc_general_every_side(gpu, ref, loc, order):
    c = H.make_case("bl", 23, 17, ref, with_w=True, order=order)
    gh, im, jm = c.gh, c.im, c.jm
    rng = np.random.default_rng(3)
    if loc[0] == "I":
        lm = jm
        i = 1 if loc == "Ilo" else im
        itf = np.array([[i, 1], [i, jm]], dtype=float)
    else:
        lm = im
        j = 1 if loc == "Jlo" else jm
        itf = np.array([[1, j], [im, j]], dtype=float)
    field = np.asfortranarray(rng.standard_normal((lm, gh, 5)))
    wa, wb = c.w.copy(order="F"), c.w.copy(order="F")
    gpu["f_bnd"].bc_general_2d(wa, loc, itf, field, gh, im, jm)
    ref["f_bnd"].bc_general_2d(wb, loc, itf, field, gh, im, jm)
    assert np.array_equal(wa, wb) and not np.array_equal(wa, c.w)
    wd = np.asfortranarray(rng.standard_normal(c.w.shape))
    da, db = wd.copy(order="F"), wd.copy(order="F")
    gpu["f_lin"].bc_general_2d_d(wa, da, loc, itf, field, gh, im, jm)
    ref["f_lin"].bc_general_2d_d(wb, db, loc, itf, field, gh, im, jm)
    assert np.array_equal(da, db) and np.array_equal(wa, wb) and not np.array_equal(da, wd)
