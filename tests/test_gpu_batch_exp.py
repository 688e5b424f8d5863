"""GPU: fixed-base get_window_table / batch_exp / batch_exp_with_coeff and batch_to_special
through the C-ABI against the reference fixtures and the oracle."""
import numpy as np
import pytest

from oracle.binding import R_ORDER, ints_to_mont
from tests import inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("grp", ["g1", "g2"])
def test_golden(engine, golden, grp):
    g = golden(f"batch_exp_{grp}")
    t = engine.get_window_table(grp, 254, int(g["window"][0]), g["base"], expected_scalars=40)
    assert (engine.batch_exp(254, 0, t, g["scalars"]) == g["batch_exp"]).all()
    assert (engine.batch_exp_with_coeff(254, 0, t, g["coeff"], g["scalars"]) == g["batch_exp_with_coeff"]).all()
    t.close()
    assert (engine.batch_exp_once(grp, g["base"], g["scalars"]) == g["batch_exp"]).all()
    assert engine.batch_exp_once(grp, g["base"], np.zeros((0, 4), dtype=np.uint64)).shape[0] == 0


@pytest.mark.parametrize("grp,n", [("g1", 20000), ("g2", 3000)])
def test_vs_oracle(engine, orc, grp, n):
    base = inputs.bases(orc, grp, 2, seed=501, affine=False)[0][1]
    s = inputs.fr_uniform(orc, n, seed=502).copy()
    s[:3] = ints_to_mont([0, 1, R_ORDER - 1], R_ORDER)
    coeff = inputs.fr_uniform(orc, 1, seed=503)[0]
    want = orc.batch_exp(grp, base, s)
    assert (engine.batch_exp_once(grp, base, s) == want).all()
    assert (engine.batch_exp_once(grp, base, s, coeff=coeff) == orc.batch_exp(grp, base, s, coeff=coeff)).all()
    # zero base: every output is zero
    z = engine.batch_exp_once(grp, inputs.zero_point(grp), s[:50])
    assert (z == np.tile(inputs.zero_point(grp), (50, 1))).all()


@pytest.mark.parametrize("grp", ["g1", "g2"])
def test_batch_to_special(engine, orc, golden, grp):
    g = golden(f"group_{grp}")
    assert (engine.batch_to_special(grp, g["P"]) == g["batch_to_special"]).all()
    P, _ = inputs.bases(orc, grp, 2500, seed=511, affine=False)
    P[::7] = inputs.zero_point(grp)
    assert (engine.batch_to_special(grp, P) == orc.batch_to_special(grp, P)).all()
