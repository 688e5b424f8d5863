#!/usr/bin/env python3
"""Generate tests/golden/wire.npz from the UNMODIFIED reference's own stream operators (operator<< / operator>> of
alt_bn128_G1/G2 and bn128_G1/G2 with point compression, through oracle/_ref/libffref.so): the compressed image
(X bytes + flag bits) of a set of points and what reading that image back gives.

    python tools/make_golden_wire.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.binding import Checker  # noqa: E402
from tests import inputs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "wire.npz")


def main():
    ref = Checker("ref")
    f = {}
    for grp, n in (("g1", 96), ("g2", 40)):
        P, _ = inputs.bases(ref, grp, n, seed=77, affine=False)  # raw Jacobian: operator<< normalises
        P[5] = ref.batch_to_special(grp, P[5:6])[0]              # an already-affine point
        for fl, curve in ((0, 0), (2, 1)):
            Pin = P.copy()
            Pin[3] = inputs.zero_point(grp, curve=curve)         # the curve's own zero
            Pin[17] = inputs.zero_point(grp, curve=curve)
            x, flags = ref.compress(grp, Pin, fl)
            f[f"{grp}_f{fl}_points"] = Pin
            f[f"{grp}_f{fl}_x"], f[f"{grp}_f{fl}_flags"] = x, flags
            f[f"{grp}_f{fl}_read_back"] = ref.decompress(grp, x, flags, fl)
    np.savez_compressed(OUT, **f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
