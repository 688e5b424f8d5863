#!/usr/bin/env python3
"""Generate tests/golden/fr_vectors.npz from the UNMODIFIED reference (oracle/_ref/liblsref.so =
LegoSNARK's poly.h / polytools.h / mle.h and libfqfft's basic_radix2_domain behind
oracle/ref_wrap_ls.cpp).  The reference ships no fixtures for these routines, so the pins are its
own outputs, produced in the build container where /root/reference exists.  Inputs are stored next
to the outputs: the file is self-contained on the GPU box.

    python tools/make_golden_fr.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.binding import Checker, R_ORDER, ints_to_mont  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fr_vectors.npz")


def main():
    ref = Checker("ref")
    f = {}
    g = ints_to_mont([5], R_ORDER)  # Fr::multiplicative_generator: the coset of r1cs_to_qap_witness_map
    g2 = ref.sha512_rng_fr(4242, 1)
    f["coset_g"], f["coset_g2"] = g, g2
    edge = ints_to_mont([0, 1, R_ORDER - 1, 2, R_ORDER - 2, 1 << 253], R_ORDER)
    dims = [1, 2, 3, 7, 10, 11]
    f["dims"] = np.array(dims)
    for d in dims:
        n = 1 << d
        v = ref.sha512_rng_fr(1000 * d, n)
        v[: min(n, 6)] = edge[: min(n, 6)]
        r = ref.sha512_rng_fr(77 * d, d)
        if d == 3:
            r[0] = edge[0]  # r_i = 0 and r_i = 1 challenge coordinates
            r[1] = edge[1]
        f[f"v_{d}"], f[f"r_{d}"] = v, r
        f[f"eval_mle_{d}"] = ref.fr_eval_mle(v, r)
        f[f"mle_bind_{d}"] = ref.fr_mle_bind(v, r[:1])
        for mode, gg in ((0, None), (1, None), (2, g), (3, g), (2, g2), (3, g2)):
            tag = f"fft_{d}_m{mode}" + ("" if gg is None or gg is g else "_g2")
            f[tag] = ref.fr_fft(v, mode, gg)
    # CPPoly::prove over an installed key of distinct bases (affine and raw Jacobian)
    for d, affine in ((1, True), (4, True), (8, False)):
        n = 1 << d
        k = ref.sha512_rng_fr(31337, n)
        P = ref.batch_exp("g1", ref.one("g1"), k, normalise=affine)
        v = ref.sha512_rng_fr(555 + d, n)
        r = ref.sha512_rng_fr(666 + d, d)
        f[f"prove_bases_{d}"], f[f"prove_v_{d}"], f[f"prove_r_{d}"] = P, v, r
        f[f"prove_witness_{d}"] = ref.cppoly_prove_g1(P, v, r)
    f["prove_dims"] = np.array([1, 4, 8])
    np.savez_compressed(OUT, **f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
