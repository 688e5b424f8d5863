"""GPU: the device field / point arithmetic (PTX path) through the C-ABI parity hooks,
bit-exact against the oracle and the reference fixtures."""
import numpy as np
import pytest

from oracle.binding import Q, R_ORDER, ints_to_mont
from tests import inputs

pytestmark = pytest.mark.gpu


def _rand_mont(rng, n, mod):
    raw = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= np.uint64((1 << 61) - 1)  # < 2^253 < p: any residue below p is a valid Montgomery image
    edge = ints_to_mont([0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2], mod)
    raw[: edge.shape[0]] = edge
    return raw


@pytest.mark.parametrize("field,mod,name", [(0, Q, "fq"), (1, R_ORDER, "fr")])
def test_prime_field_elementwise_2pow16(engine, orc, field, mod, name):
    rng = np.random.default_rng(100 + field)
    n = 1 << 16
    a, b = _rand_mont(rng, n, mod), _rand_mont(rng, n, mod)[::-1].copy()
    for op in (0, 1, 2, 3, 5):
        assert (engine.test_field_op(field, op, a, b) == orc.field_op(name, op, a, b)).all(), op
    # op 8 = a b - b (a + b): the fused two-product Montgomery pass (mul_sub)
    want8 = orc.field_op(name, 3, orc.field_op(name, 0, a, b), orc.field_op(name, 0, b, orc.field_op(name, 2, a, b)))
    assert (engine.test_field_op(field, 8, a, b) == want8).all()
    nz = a[(a != 0).any(axis=1)][:2000]
    assert (engine.test_field_op(field, 4, nz) == orc.field_op(name, 4, nz)).all()
    if name == "fr":
        assert (engine.test_field_op(1, 6, a) == orc.fr_as_bigint(a)).all()
        assert (engine.test_field_op(1, 7, a) == orc.fr_from_bigint(a)).all()


def test_fq2_elementwise(engine, orc):
    rng = np.random.default_rng(102)
    n = 1 << 14
    a = np.concatenate([_rand_mont(rng, n, Q), _rand_mont(rng, n, Q)[::-1]], axis=1)
    b = np.concatenate([_rand_mont(rng, n, Q)[::-1], _rand_mont(rng, n, Q)], axis=1)
    for op in (0, 1, 2, 3, 5):
        assert (engine.test_field_op(2, op, a, b) == orc.field_op("fq2", op, a, b)).all(), op
    want8 = orc.field_op("fq2", 3, orc.field_op("fq2", 0, a, b), orc.field_op("fq2", 0, b, orc.field_op("fq2", 2, a, b)))
    assert (engine.test_field_op(2, 8, a, b) == want8).all()
    nz = a[(a != 0).any(axis=1)][:500]
    assert (engine.test_field_op(2, 4, nz) == orc.field_op("fq2", 4, nz)).all()


def test_field_golden(engine, golden):
    g = golden("fields")
    for field, name in ((0, "fq"), (1, "fr"), (2, "fq2")):
        for op, opn in enumerate(("mul", "sqr", "add", "sub", "inv", "neg")):
            x = g[f"{name}_inv_in"] if opn == "inv" else g[f"{name}_a"]
            assert (engine.test_field_op(field, op, x, g[f"{name}_b"]) == g[f"{name}_{opn}"]).all(), (name, opn)


@pytest.mark.parametrize("gi,grp", [(0, "g1"), (1, "g2")])
def test_group_formulas(engine, orc, golden, gi, grp):
    g = golden(f"group_{grp}")
    P, Qj, Qa = g["P"], g["Q"], g["Q_affine"]
    norm = lambda x: orc.group_op(grp, 3, x)
    assert (norm(engine.test_group_op(gi, 0, P, Qj)) == norm(g["add"])).all()
    assert (norm(engine.test_group_op(gi, 1, P, Qa)) == norm(g["mixed_add"])).all()
    assert (norm(engine.test_group_op(gi, 2, P)) == norm(g["dbl"])).all()
    assert (norm(engine.test_group_op(gi, 6, P, Qa)) == norm(orc.group_op(grp, 1, P, orc.group_op(grp, 4, Qa)))).all()
    assert (norm(engine.test_group_op(gi, 8, P)) == norm(P)).all()
    for k in (0, 1, 2, 3, 255, 32767, 40000):
        want = orc.scalar_mul(grp, P, np.tile(ints_to_mont([k], R_ORDER), (P.shape[0], 1)), stride_base=True)
        assert (norm(engine.test_group_op(gi, 7, P, None, k)) == want).all(), k


@pytest.mark.parametrize("gi,grp,n", [(0, "g1", 4096), (1, "g2", 1024)])
def test_group_random_vs_oracle(engine, orc, gi, grp, n):
    P, _ = inputs.bases(orc, grp, n, seed=201, affine=False)
    Qa, _ = inputs.bases(orc, grp, n, seed=202, affine=True)
    norm = lambda x: orc.group_op(grp, 3, x)
    assert (norm(engine.test_group_op(gi, 1, P, Qa)) == norm(orc.group_op(grp, 1, P, Qa))).all()
    assert (norm(engine.test_group_op(gi, 0, P, Qa)) == norm(orc.group_op(grp, 0, P, Qa))).all()
    assert (engine.batch_to_special(grp, P) == orc.batch_to_special(grp, P)).all()
