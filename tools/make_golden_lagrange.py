#!/usr/bin/env python3
"""Generate tests/golden/lagrange.npz from the UNMODIFIED reference (oracle/_ref/liblsref.so): libfqfft's
evaluate_all_lagrange_polynomials(t) for basic_radix2_domain (basic_radix2_domain_aux.tcc:183-236) and step_radix2_domain
(step_radix2_domain.tcc:161-186) at random points and at points of the domain itself (the unit-vector branch).  Inputs are
stored next to the outputs: the file is self-contained on the GPU box.

    python tools/make_golden_lagrange.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.binding import Checker, R_ORDER, ints_to_mont  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "lagrange.npz")
ROOT_2_28 = 19103219067921713944291392827692070036145651957329286315305642004821462161904  # Fr::root_of_unity


def main():
    ref = Checker("ref")
    f = {}
    shapes = [(1, -1), (2, -1), (5, -1), (6, -1), (9, -1), (12, -1), (1, 0), (3, 0), (4, 2), (6, 5), (9, 0), (11, 7), (12, 0)]
    f["shapes"] = np.array(shapes)
    for lb, ls in shapes:
        lsn = None if ls < 0 else ls
        t = ref.sha512_rng_fr(9700 + 16 * lb + ls + 1, 1)
        f[f"t_{lb}_{ls}"] = t
        f[f"u_{lb}_{ls}"] = ref.fr_lagrange(lb, lsn, t)
        # t on the domain: a point of the 2^lb-th roots of unity (big half / basic domain) ...
        order = lb if ls < 0 else lb  # big_omega = omega^2 has order 2^lb
        w = pow(ROOT_2_28, 1 << (28 - order), R_ORDER)
        k = (1 << lb) // 3
        tin = ints_to_mont([pow(w, k, R_ORDER)], R_ORDER)
        f[f"tin_{lb}_{ls}"] = tin
        f[f"uin_{lb}_{ls}"] = ref.fr_lagrange(lb, lsn, tin)
        if ls >= 1:  # ... and a point of the small half: omega * small_omega^j
            omega = pow(ROOT_2_28, 1 << (28 - (lb + 1)), R_ORDER)
            so = pow(ROOT_2_28, 1 << (28 - ls), R_ORDER)
            ts = ints_to_mont([omega * pow(so, (1 << ls) // 3, R_ORDER) % R_ORDER], R_ORDER)
            f[f"tsm_{lb}_{ls}"] = ts
            f[f"usm_{lb}_{ls}"] = ref.fr_lagrange(lb, lsn, ts)
    np.savez_compressed(OUT, **f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
