"""CPU: the device arithmetic headers (legosnark_b200/csrc/{ptx_ops,field,curve}.cuh)
compiled for the host with an emulated carry flag, checked bit-for-bit against
the oracle.  This validates the even/odd-accumulator Montgomery product, the
conditional-subtract add/sub chains and the XYZZ point formulas (including the
doubling / inverse / infinity branches) before any GPU time is spent; the same
entry points run as CUDA kernels in tests/test_gpu_arith.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.binding import Q, R_ORDER, ints_to_mont, int_to_limbs
from tests import inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_u64p = ctypes.POINTER(ctypes.c_uint64)


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "tests", "emu", "emu_lib.cpp")
    so = os.path.join(ROOT, "tests", "emu", "libb200emu.so")
    csrc = os.path.join(ROOT, "legosnark_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("ptx_ops.cuh", "field.cuh", "curve.cuh", "test_ops.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I" + csrc, "-o", so, src], check=True)
    return ctypes.CDLL(so)


def _p(a):
    return None if a is None else a.ctypes.data_as(_u64p)


def field_op(emu, field, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = None if b is None else np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros_like(a)
    assert emu.emu_field_op(field, op, _p(a), _p(b), ctypes.c_size_t(a.shape[0]), _p(out)) == 0
    return out


def group_op(emu, group, op, a, b=None, k=0):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = None if b is None else np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros_like(a)
    assert emu.emu_group_op(group, op, _p(a), _p(b), ctypes.c_size_t(a.shape[0]), ctypes.c_uint32(k), _p(out)) == 0
    return out


def edge_and_random(rng, n, mod):
    xs = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << 253) % mod, (1 << 32) - 1, 1 << 32, (1 << 224) - 1]
    xs += [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(n - len(xs))]
    return xs


@pytest.mark.parametrize("field,mod,name", [(0, Q, "fq"), (1, R_ORDER, "fr")])
def test_prime_field_ops(emu, orc, field, mod, name):
    rng = np.random.default_rng(7 + field)
    xs = edge_and_random(rng, 3000, mod)
    ys = edge_and_random(rng, 3000, mod)[::-1]
    # Montgomery images (any residue is a valid image) plus raw edge residues
    a = ints_to_mont(xs, mod)
    b = ints_to_mont(ys, mod)
    a[:10] = np.array([int_to_limbs(v) for v in xs[:10]])
    for op in (0, 1, 2, 3, 5):
        assert (field_op(emu, field, op, a, b) == orc.field_op(name, op, a, b)).all(), op
    # op 8 = a b - b (a + b) through the fused two-product Montgomery pass
    want8 = orc.field_op(name, 3, orc.field_op(name, 0, a, b), orc.field_op(name, 0, b, orc.field_op(name, 2, a, b)))
    assert (field_op(emu, field, 8, a, b) == want8).all()
    nz = a[(a != 0).any(axis=1)][:200]
    assert (field_op(emu, field, 4, nz) == orc.field_op(name, 4, nz)).all()
    if name == "fr":
        assert (field_op(emu, field, 6, a) == orc.fr_as_bigint(a)).all()
        assert (field_op(emu, field, 7, a) == orc.fr_from_bigint(a)).all()


def test_fq2_ops(emu, orc):
    rng = np.random.default_rng(9)
    a = np.concatenate([ints_to_mont(edge_and_random(rng, 1000, Q), Q), ints_to_mont(edge_and_random(rng, 1000, Q)[::-1], Q)], axis=1)
    b = np.concatenate([ints_to_mont(edge_and_random(rng, 1000, Q)[::-1], Q), ints_to_mont(edge_and_random(rng, 1000, Q), Q)], axis=1)
    for op in (0, 1, 2, 3, 5):
        assert (field_op(emu, 2, op, a, b) == orc.field_op("fq2", op, a, b)).all(), op
    want8 = orc.field_op("fq2", 3, orc.field_op("fq2", 0, a, b), orc.field_op("fq2", 0, b, orc.field_op("fq2", 2, a, b)))
    assert (field_op(emu, 2, 8, a, b) == want8).all()
    nz = a[(a != 0).any(axis=1)][:100]
    assert (field_op(emu, 2, 4, nz) == orc.field_op("fq2", 4, nz)).all()


@pytest.mark.parametrize("gi,grp", [(0, "g1"), (1, "g2")])
def test_group_formulas(emu, orc, golden, gi, grp):
    g = golden(f"group_{grp}")
    P, Qj, Qa = g["P"], g["Q"], g["Q_affine"]
    norm = lambda x: orc.group_op(grp, 3, x)
    assert (norm(group_op(emu, gi, 0, P, Qj)) == norm(g["add"])).all()
    assert (norm(group_op(emu, gi, 1, P, Qa)) == norm(g["mixed_add"])).all()
    assert (norm(group_op(emu, gi, 2, P)) == norm(g["dbl"])).all()
    # a - affine(b) == a + (-b)
    negQ = orc.group_op(grp, 4, Qa)
    assert (norm(group_op(emu, gi, 6, P, Qa)) == norm(orc.group_op(grp, 1, P, negQ))).all()
    # round trip Jacobian -> XYZZ -> Jacobian keeps the point
    assert (norm(group_op(emu, gi, 8, P)) == norm(P)).all()
    # small scalar multiples
    for k in (0, 1, 2, 3, 5, 16, 255, 32767, 40000):
        want = orc.scalar_mul(grp, P, np.tile(ints_to_mont([k], R_ORDER), (P.shape[0], 1)), stride_base=True)
        assert (norm(group_op(emu, gi, 7, P, None, k)) == want).all(), k


@pytest.mark.parametrize("field,mod,name", [(0, Q, "fq"), (1, R_ORDER, "fr")])
def test_paired_products(emu, orc, field, mod, name):
    """Fp::mul2 (two independent Montgomery products with alternating rows) and Fp::mul_add2 (two fused two-product passes
    with alternating rows): measured options of the accumulation kernels; same limbs as the plain forms."""
    rng = np.random.default_rng(21 + field)
    n = 2000
    a, b, c, d = (ints_to_mont(edge_and_random(rng, n, mod), mod) for _ in range(4))
    b, d = b[::-1].copy(), d[::-1].copy()
    a[:10] = np.array([int_to_limbs(v) for v in edge_and_random(rng, 10, mod)])  # raw residues are valid images too
    out0, out1 = np.zeros_like(a), np.zeros_like(a)
    args = lambda: (_p(a), _p(b), _p(c), _p(d), ctypes.c_size_t(n), _p(out0), _p(out1))
    assert emu.emu_pair_field_op(field, 0, *args()) == 0
    assert (out0 == orc.field_op(name, 0, a, b)).all() and (out1 == orc.field_op(name, 0, c, d)).all()
    assert emu.emu_pair_field_op(field, 1, *args()) == 0
    want0 = orc.field_op(name, 2, orc.field_op(name, 0, a, b), orc.field_op(name, 0, c, d))
    want1 = orc.field_op(name, 2, orc.field_op(name, 0, a, d), orc.field_op(name, 0, c, b))
    assert (out0 == want0).all() and (out1 == want1).all()


def test_paired_mixed_addition(emu, orc, golden):
    """xyzz_madd_paired (the mixed addition with its independent products issued in pairs, knob g1_paired) on the reference
    fixtures, including the same-point (tangent) and opposite-point rows they contain."""
    g = golden("group_g1")
    P, Qa = g["P"], g["Q_affine"]
    norm = lambda x: orc.group_op("g1", 3, x)
    out = np.zeros_like(P)
    assert emu.emu_madd_paired_g1(_p(np.ascontiguousarray(P)), _p(np.ascontiguousarray(Qa)), ctypes.c_size_t(P.shape[0]), 0, _p(out)) == 0
    assert (norm(out) == norm(g["mixed_add"])).all()
    negQ = orc.group_op("g1", 4, Qa)
    assert emu.emu_madd_paired_g1(_p(np.ascontiguousarray(P)), _p(np.ascontiguousarray(Qa)), ctypes.c_size_t(P.shape[0]), 1, _p(out)) == 0
    assert (norm(out) == norm(orc.group_op("g1", 1, P, negQ))).all()
    # P + P (tangent branch) and P - P (infinity) through the paired adder
    Pa = norm(P)
    assert emu.emu_madd_paired_g1(_p(np.ascontiguousarray(Pa)), _p(np.ascontiguousarray(Pa)), ctypes.c_size_t(Pa.shape[0]), 0, _p(out)) == 0
    assert (norm(out) == norm(orc.group_op("g1", 2, Pa))).all()
    assert emu.emu_madd_paired_g1(_p(np.ascontiguousarray(Pa)), _p(np.ascontiguousarray(Pa)), ctypes.c_size_t(Pa.shape[0]), 1, _p(out)) == 0
    zero = norm(out)
    assert (zero[:, 8:12] == 0).all()  # Z == 0: the point at infinity
