"""Deterministic synthetic inputs for the parity tests, golden fixtures and bench.

Bases are k_i * G built with a checker's fixed-base batch_exp (the reference's
own way of making commitment keys, LS/prototools/interp.h:36-59); scalars follow
SURVEY.md §8(d): SHA512_rng uniform (LFF/common/rng.tcc:26-72), 32-bit
(legogrothmatrix.cc:29-32), 0/1-heavy witnesses, plus the adversarial corpus.
Everything is numpy uint64 Montgomery limbs.
"""
from __future__ import annotations

import numpy as np

from oracle.binding import Q, R_ORDER, MONT_R, int_to_limbs, ints_to_mont

LIMBS = {"g1": 12, "g2": 24}


def fr_uniform(chk, n, seed=0):
    """SHA512_rng<Fr>(seed * 2^40 + i), Montgomery limbs."""
    return chk.sha512_rng_fr(seed << 40, n)


def fr_small(n, bits=32, seed=0):
    rng = np.random.default_rng(seed + 77)
    xs = [int(v) for v in rng.integers(0, 1 << bits, size=n, dtype=np.uint64)]
    return ints_to_mont(xs, R_ORDER)


def fr_zero_one_heavy(chk, n, seed=0, frac=0.9):
    rng = np.random.default_rng(seed + 99)
    s = fr_uniform(chk, n, seed + 5).copy()
    pick = rng.random(n)
    one = ints_to_mont([1], R_ORDER)[0]
    s[pick < frac / 2] = 0
    s[(pick >= frac / 2) & (pick < frac)] = one
    return s


def fr_const(n, value):
    return np.tile(ints_to_mont([value], R_ORDER), (n, 1))


def bases(chk, group, n, seed=1, affine=True):
    """P_i = k_i * G with k_i = SHA512_rng(seed*2^40 + i); affine (Z = 1) or raw Jacobian."""
    k = fr_uniform(chk, n, seed)
    return chk.batch_exp(group, chk.one(group), k, normalise=affine), k


def zero_point(group, curve=0):
    """The curve's own zero: alt_bn128 (0,1,0) (alt_bn128_init.cpp:145-147), bn128 (1,1,0) (bn128_init.cpp:101-103)."""
    L = LIMBS[group]
    w = L // 3
    one = int_to_limbs(MONT_R % Q)
    z = np.zeros(L, dtype=np.uint64)
    z[w:w + 4] = one
    if curve == 1:
        z[0:4] = one
    return z


def negate(chk, group, pts):
    return chk.group_op(group, 4, pts)


def scalar_sum_check(chk, group, k, s):
    """Independent oracle of SURVEY.md §8(c): sum s_i (k_i G) == (sum s_i k_i mod r) G, affine."""
    from oracle.binding import mont_to_ints
    ks = mont_to_ints(k, R_ORDER)
    ss = mont_to_ints(s, R_ORDER)
    tot = sum(a * b for a, b in zip(ks, ss)) % R_ORDER
    return chk.scalar_mul(group, chk.one(group), ints_to_mont([tot], R_ORDER), normalise=True)[0]


def fr_fast_uniform(n, seed=0):
    """Uniform 253-bit integers read as Montgomery images (every residue below r is one): numpy's generator, for
    the 2^22..2^26 cases where SHA512_rng's one hash per element would take minutes."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 61) - 1)
    return a


def scalar_sum_fast(k, s):
    """sum k_i s_i mod r for Montgomery-form limb arrays, exact and vectorised: both vectors are cut into 16-bit
    pieces, the 16 x 16 piece-by-piece dot products run as float64 matrix products over chunks of 2^19 rows
    (every partial sum stays below 2^51, so the floating-point sums are exact), and the pieces are recombined
    with Python integers.  Returns the plain (non-Montgomery) integer."""
    k = np.ascontiguousarray(k, dtype=np.uint64).view(np.uint16).reshape(len(k), 16)
    s = np.ascontiguousarray(s, dtype=np.uint64).view(np.uint16).reshape(len(s), 16)
    acc = [[0] * 16 for _ in range(16)]
    step = 1 << 19
    for lo in range(0, len(k), step):
        m = k[lo:lo + step].astype(np.float64).T @ s[lo:lo + step].astype(np.float64)
        for a in range(16):
            for b in range(16):
                acc[a][b] += int(m[a, b])
    tot = sum(acc[a][b] << (16 * (a + b)) for a in range(16) for b in range(16))
    rinv = pow(MONT_R, -1, R_ORDER)
    return tot * rinv * rinv % R_ORDER


def scalar_sum_point(chk, group, k, s):
    """(sum s_i k_i mod r) * G as a normalised point: SURVEY.md 8(c)'s independent oracle at any n."""
    return chk.scalar_mul(group, chk.one(group), ints_to_mont([scalar_sum_fast(k, s)], R_ORDER), normalise=True)[0]


def msm_cases(chk, group, sizes=(0, 1, 2, 3, 17, 64, 257)):
    """Named (bases, scalars) cases covering SURVEY.md §8(d)'s distributions and edge corpus."""
    nmax = max(max(sizes), 64)
    P, _ = bases(chk, group, nmax, seed=1, affine=True)
    PJ, _ = bases(chk, group, 64, seed=2, affine=False)
    cases = {}
    for n in sizes:
        cases[f"uniform_n{n}"] = (P[:n], fr_uniform(chk, n, seed=3))
    n = 64
    cases["jacobian_bases"] = (PJ, fr_uniform(chk, n, seed=4))
    cases["scalars_32bit"] = (P[:n], fr_small(n, 32))
    cases["zero_one_heavy"] = (P[:n], fr_zero_one_heavy(chk, n))
    cases["all_zero_scalars"] = (P[:n], fr_const(n, 0))
    cases["all_one_scalars"] = (P[:n], fr_const(n, 1))
    cases["all_rminus1"] = (P[:n], fr_const(n, R_ORDER - 1))
    cases["all_equal_bases"] = (np.tile(P[5], (n, 1)), fr_uniform(chk, n, seed=6))
    cases["all_equal_bases_equal_scalars"] = (np.tile(P[5], (n, 1)), fr_const(n, 123456789))
    # P_i = -P_j pairs with equal scalars: cancels to zero
    half = P[: n // 2]
    pm = np.concatenate([half, negate(chk, group, half)])
    s_half = fr_uniform(chk, n // 2, seed=7)
    cases["cancelling_pairs"] = (pm, np.concatenate([s_half, s_half]))
    # infinity bases sprinkled in
    Pi = P[:n].copy()
    Pi[::5] = zero_point(group)
    cases["infinity_bases"] = (Pi, fr_uniform(chk, n, seed=8))
    # one hot bucket: same scalar everywhere, distinct bases
    cases["single_hot_bucket"] = (P[:n], fr_const(n, 0x1234))
    return cases
