#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libffref.so).

The reference ships no golden vectors for multi_exp / batch_exp,
so the pins are outputs of the reference itself, run in the build container where
/root/reference exists.  Every input array is stored next to the reference's
output, so the fixtures are self-contained on the GPU box.

    python tools/make_golden.py          # rewrites tests/golden/
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.binding import Checker, Q, R_ORDER, ints_to_mont  # noqa: E402
from tests import inputs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def rand_field(rng, n, mod):
    xs = [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(n)]
    # edge values first
    xs[:6] = [0, 1, 2, mod - 1, mod - 2, (1 << 253) % mod]
    return ints_to_mont(xs, mod)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = Checker("ref")
    rng = np.random.default_rng(20261017)

    # ---- field KATs: Fp_model / Fp2_model operators -----------------------------
    f = {}
    for name, mod in (("fq", Q), ("fr", R_ORDER)):
        a, b = rand_field(rng, 64, mod), rand_field(rng, 64, mod)[::-1].copy()
        f[f"{name}_a"], f[f"{name}_b"] = a, b
        for op, opn in enumerate(("mul", "sqr", "add", "sub", "inv", "neg")):
            if opn == "inv":
                nz = a.copy()
                nz[(nz == 0).all(axis=1)] = ints_to_mont([7], mod)[0]
                f[f"{name}_inv_in"] = nz
                f[f"{name}_inv"] = ref.field_op(name, op, nz, b)
            else:
                f[f"{name}_{opn}"] = ref.field_op(name, op, a, b)
    a2 = np.concatenate([rand_field(rng, 64, Q), rand_field(rng, 64, Q)[::-1]], axis=1)
    b2 = np.concatenate([rand_field(rng, 64, Q)[::-1], rand_field(rng, 64, Q)], axis=1)
    f["fq2_a"], f["fq2_b"] = a2, b2
    for op, opn in enumerate(("mul", "sqr", "add", "sub", "inv", "neg")):
        if opn == "inv":
            nz = a2.copy()
            nz[(nz == 0).all(axis=1)] = 3
            f["fq2_inv_in"] = nz
            f["fq2_inv"] = ref.field_op("fq2", op, nz, b2)
        else:
            f[f"fq2_{opn}"] = ref.field_op("fq2", op, a2, b2)
    raw = np.array([[int(x) for x in rng.integers(0, 1 << 62, 4)] for _ in range(16)], dtype=np.uint64)
    f["fr_bigint_in"] = raw
    f["fr_from_bigint"] = ref.fr_from_bigint(raw)
    f["fr_as_bigint"] = ref.fr_as_bigint(f["fr_a"])
    f["sha512_rng_fr_idx0"] = ref.sha512_rng_fr(0, 32)
    f["sha512_rng_fr_idx_1e12"] = ref.sha512_rng_fr(10**12, 8)
    np.savez_compressed(os.path.join(OUT, "fields.npz"), **f)

    # ---- group-law KATs (test_groups.cpp identities incl. doubling / inverse / zero cases) ----
    for grp in ("g1", "g2"):
        g = {}
        n = 24
        PJ, _ = inputs.bases(ref, grp, n, seed=11, affine=False)
        QJ, _ = inputs.bases(ref, grp, n, seed=12, affine=False)
        QA = ref.batch_to_special(grp, QJ)
        zero = inputs.zero_point(grp)
        # edge rows: 0: P+P (doubling, different Z), 1: P+(-P), 2: 0+Q, 3: P+0, 4: 0+0, 5: affine P + same affine P
        PA = ref.batch_to_special(grp, PJ)
        QJ[0], QA[0] = PJ[0], PA[0]
        QJ[1] = ref.group_op(grp, 4, PJ[1:2])[0]
        QA[1] = ref.group_op(grp, 4, PA[1:2])[0]
        PJ[2] = zero
        QJ[3], QA[3] = zero, zero
        PJ[4], QJ[4], QA[4] = zero, zero, zero
        PJ[5], QJ[5], QA[5] = PA[5], PA[5], PA[5]
        g["P"], g["Q"], g["Q_affine"] = PJ, QJ, QA
        g["add"] = ref.group_op(grp, 0, PJ, QJ)
        g["add_explicit"] = ref.group_op(grp, 5, PJ, QJ)
        g["mixed_add"] = ref.group_op(grp, 1, PJ, QA)
        g["dbl"] = ref.group_op(grp, 2, PJ)
        g["neg"] = ref.group_op(grp, 4, PJ)
        g["to_affine"] = ref.group_op(grp, 3, PJ)
        g["batch_to_special"] = ref.batch_to_special(grp, PJ)
        g["one"] = ref.one(grp)
        sc = inputs.fr_uniform(ref, 8, seed=13)
        g["scalar_mul_scalars"] = sc
        g["scalar_mul_one"] = ref.scalar_mul(grp, ref.one(grp), sc, normalise=True)
        # (r1 a) + (r2 a) == (r1 + r2) a with test_groups.cpp:94-95's fixed r1, r2
        r12 = ints_to_mont([76749407, 44410867, 76749407 + 44410867], R_ORDER)
        g["r1r2_scalars"] = r12
        g["r1r2_mul"] = ref.scalar_mul(grp, PA[7], r12, normalise=True)
        np.savez_compressed(os.path.join(OUT, f"group_{grp}.npz"), **g)

    # ---- MSM KATs: every distribution of SURVEY §8(d), both entry points, chunked and not ----
    for grp, sizes in (("g1", (0, 1, 2, 3, 17, 64, 257, 1026)), ("g2", (0, 1, 2, 3, 17, 64, 130))):
        m = {}
        names = []
        for name, (B, S) in inputs.msm_cases(ref, grp, sizes).items():
            names.append(name)
            m[f"{name}__bases"], m[f"{name}__scalars"] = B, S
            r = ref.msm(grp, B, S, chunks=1, variant=0)
            for variant in (0, 1):
                for chunks in (1, 3, 8):
                    r2 = ref.msm(grp, B, S, chunks=chunks, variant=variant)
                    assert (r == r2).all(), (grp, name, variant, chunks)
            m[f"{name}__result"] = r
            if ref.lib.ref_has_bn128():
                rb = ref.msm(grp, B, S, chunks=1, variant=1, curve=1)
                zero = (r.reshape(3, -1)[2] == 0).all()
                if zero:  # bn128 zero is (1,1,0)
                    assert (rb.reshape(3, -1)[2] == 0).all()
                else:
                    assert (rb == r).all(), (grp, name)
        m["names"] = np.array(names)
        np.savez_compressed(os.path.join(OUT, f"msm_{grp}.npz"), **m)

    # ---- batch_exp KATs -------------------------------------------------------------
    for grp in ("g1", "g2"):
        b = {}
        base = inputs.bases(ref, grp, 3, seed=21, affine=False)[0][2]
        sc = inputs.fr_uniform(ref, 40, seed=22).copy()
        sc[0] = 0
        sc[1] = ints_to_mont([1], R_ORDER)[0]
        sc[2] = ints_to_mont([R_ORDER - 1], R_ORDER)[0]
        coeff = inputs.fr_uniform(ref, 1, seed=23)[0]
        b["base"], b["scalars"], b["coeff"] = base, sc, coeff
        b["window"] = np.array([ref.exp_window_size(grp, 40)])
        b["batch_exp"] = ref.batch_exp(grp, base, sc, normalise=True)
        b["batch_exp_with_coeff"] = ref.batch_exp(grp, base, sc, coeff=coeff, normalise=True)
        b["window_sizes_n"] = np.array([1, 4, 5, 100, 7122, 57818, 1 << 16, 1 << 20, 1 << 22, 1 << 24, 1 << 26])
        b["window_sizes"] = np.array([ref.exp_window_size(grp, int(n)) for n in b["window_sizes_n"]])
        np.savez_compressed(os.path.join(OUT, f"batch_exp_{grp}.npz"), **b)

    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {sorted(os.listdir(OUT))} ({tot / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
