#!/usr/bin/env python3
"""Generate tests/golden/sumcheck.npz from the UNMODIFIED reference (oracle/_ref/liblsref.so): DPBeta::compute_eq_tbl,
DPMatrixMle's constructor, the round polynomials of CPSumcheck::prove's loop (make_new_h_poly + pushRandomness) with
DPBetaDummy (the matrix sum-check) and with a real DPBeta, and DPBeta's round-0 suffix table.  Inputs are stored next
to the outputs: the file is self-contained on the GPU box.

    python tools/make_golden_sumcheck.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.binding import Checker, R_ORDER, ints_to_mont  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "sumcheck.npz")


def main():
    ref = Checker("ref")
    f = {}
    edge = ints_to_mont([0, 1, R_ORDER - 1, 2], R_ORDER)
    dims = [1, 2, 3, 5, 8]
    f["dims"] = np.array(dims)
    for d in dims:
        n = 1 << d
        rho = ref.sha512_rng_fr(9000 + d, d)
        r = ref.sha512_rng_fr(9100 + d, d)
        if d == 3:
            r[0], r[1], r[2] = edge[0], edge[1], edge[2]  # challenge coordinates 0, 1 and r - 1 (DPBeta inverts rho: it stays random)
        a = ref.sha512_rng_fr(9200 + d, n)
        b = ref.sha512_rng_fr(9300 + d, n)
        a[: min(n, 4)] = edge[: min(n, 4)]
        f[f"rho_{d}"], f[f"r_{d}"], f[f"a_{d}"], f[f"b_{d}"] = rho, r, a, b
        f[f"eq_{d}"] = ref.fr_eq_table(rho)
        f[f"eq_r_{d}"] = ref.fr_eq_table(r)
        h, nc = ref.sumcheck_h_polys(a, b, None, r)
        assert nc == 3
        f[f"h_dummy_{d}"] = h[:, :3]
        if d >= 2:
            h, nc = ref.sumcheck_h_polys(a, b, rho, r)
            assert nc == 4
            f[f"h_beta_{d}"] = h
            f[f"beta_suffix_{d}"] = ref.fr_beta_suffix(rho)
    mdims = [1, 2, 4, 6]
    f["matrix_dims"] = np.array(mdims)
    for d in mdims:
        n = 1 << d
        A = ref.sha512_rng_fr(9400 + d, n * n)
        A[: min(n * n, 4)] = edge[: min(n * n, 4)]
        rho = ref.sha512_rng_fr(9500 + d, d)
        f[f"A_{d}"], f[f"mrho_{d}"] = A, rho
        f[f"matrix_mle_{d}"] = ref.fr_matrix_mle(A, rho)
    # libfqfft's step_radix2_domain (2^k + 2^r points): the four transforms and divide_by_Z_on_coset
    g5 = ints_to_mont([5], R_ORDER)  # Fr::multiplicative_generator: the coset of r1cs_to_qap_witness_map
    shapes = [(2, 0), (3, 1), (5, 0), (7, 4), (11, 0)]
    f["step_shapes"] = np.array(shapes)
    for lb, ls in shapes:
        m = (1 << lb) + (1 << ls)
        a = ref.sha512_rng_fr(9600 + lb, m)
        a[: min(m, 4)] = edge[: min(m, 4)]
        f[f"step_a_{lb}_{ls}"] = a
        for mode in range(4):
            f[f"step_{lb}_{ls}_m{mode}"] = ref.fr_step_fft(a, lb, ls, mode, g5)
        f[f"step_{lb}_{ls}_divz"] = ref.fr_step_divide_z(a, lb, ls)
    np.savez_compressed(OUT, **f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
