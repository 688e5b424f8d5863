#!/usr/bin/env python3
"""tools/sweep.py — MSM time vs n (and optional (c, L) tuning grid) on one GPU.

    python tools/sweep.py [--group g1] [--log2n 10,12,...] [--tune]

Per size: wall time of the host-buffer call b200_msm_* (cold bases, Z = 1 and Z != 1), of the
resident-key call with host scalars, and the device pipeline time from b200_last_stats.
Bases are made on the device (fixed-base kernel); results are cross-checked between the paths.
Prints one JSON line per size."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import legosnark_b200 as lb
from bench import generator, random_scalars


def timeit(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return r, float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default="g1")
    ap.add_argument("--log2n", default="10,12,14,16,18,20")
    ap.add_argument("--tune", action="store_true")
    ap.add_argument("--reps", type=int, default=7)
    a = ap.parse_args()
    g = a.group
    L = 12 if g == "g1" else 24
    lb.init(1)
    for l2 in [int(x) for x in a.log2n.split(",")]:
        n = (1 << l2) + (2 if l2 == 10 else 0)   # 1026 = cplink's prove size
        s = random_scalars(n, 1)
        k = random_scalars(n, 2)
        P = lb.batch_exp_once(g, generator(g), k)            # normalised Jacobian (Z = 1)
        # a non-normalised representative of the same points: (X l^2, Y l^3, l) with l = 3 -> via group op 2a - a? keep simple: use a + a - a path
        key = lb.CommitmentKey(g, bases=P)
        row = {"group": g, "n": n}
        r0, med, mn = timeit(lambda: lb.multi_exp(g, P, s), a.reps)
        st = lb.last_stats()
        row.update(cold_ms=med, cold_min_ms=mn, c=st["window_bits"], W=st["num_windows"], L=st["chunk_len"],
                   device_ms=st["device_ms"], accumulate_ms=st["accumulate_ms"], finalize_us=st["host_finalize_us"],
                   launches=st["kernel_launches"])
        r1, med, mn = timeit(lambda: key.multi_exp(s), a.reps)
        st = lb.last_stats()
        row.update(resident_ms=med, resident_min_ms=mn, resident_device_ms=st["device_ms"])
        assert (r0 == r1).all()
        if a.tune:
            best = None
            for c in range(max(4, st["window_bits"] - 3), st["window_bits"] + 3):
                for Lc in (8, 16, 32, 64):
                    lb.set_tuning(c, Lc)
                    r2, med, mn = timeit(lambda: key.multi_exp(s), 5)
                    assert (r2 == r0).all()
                    if best is None or med < best[0]:
                        best = (med, c, Lc, lb.last_stats()["device_ms"])
            lb.set_tuning(0, 0)
            row.update(tuned_resident_ms=best[0], tuned_c=best[1], tuned_L=best[2], tuned_device_ms=best[3])
        key.close()
        print(json.dumps(row), flush=True)
    lb.shutdown()


if __name__ == "__main__":
    main()
