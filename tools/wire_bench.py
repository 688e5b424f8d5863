"""Point (de)compression throughput through host buffers on one GPU, with the reference's own stream operators
(oracle/_ref) timed on a sample beside it.  One JSON line per (group, direction)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import legosnark_b200 as lb
from bench import generator, random_scalars
from oracle.binding import Checker

lb.init(1)
ref = Checker("ref") if Checker.available("ref") else None
for grp, d in (("g1", 20), ("g2", 18)):
    n = 1 << d
    P = lb.batch_exp_once(grp, generator(grp), random_scalars(n, 3))
    for fl in (0, 2):
        lb.compress_points(grp, P[:1024], fl)
        t0 = time.perf_counter(); x, flags = lb.compress_points(grp, P, fl); tc = (time.perf_counter() - t0) * 1e3
        lb.decompress_points(grp, x[:1024], flags[:1024], fl)
        t0 = time.perf_counter(); back = lb.decompress_points(grp, x, flags, fl); td = (time.perf_counter() - t0) * 1e3
        line = {"group": grp, "log2n": d, "flavour": fl, "compress_ms": tc, "decompress_ms": td, "round_trip": bool((back == P).all()),
                "compress_points_per_s": n / (tc * 1e-3), "decompress_points_per_s": n / (td * 1e-3)}
        if ref is not None:
            m = 1 << 13
            t0 = time.perf_counter(); rx, rf = ref.compress(grp, P[:m], fl); rc = (time.perf_counter() - t0)
            t0 = time.perf_counter(); rb = ref.decompress(grp, rx, rf, fl); rd = (time.perf_counter() - t0)
            line.update(reference_sample=m, reference_compress_points_per_s=m / rc, reference_decompress_points_per_s=m / rd,
                        bit_identical=bool((rx == x[:m]).all() and (rf == flags[:m]).all()))
        print(json.dumps(line), flush=True)
lb.shutdown()
