#!/usr/bin/env python3
"""tools/identity_check.py — the scalar-sum identity  sum s_i (k_i G) == (sum s_i k_i mod r) G   at sizes
that are too slow for the driver's pytest run (2^26 G1, 2^22 / 2^24 G2): bases made by the GPU fixed-base path and
spot-checked against the oracle, the right-hand side computed exactly on the host (tests/inputs.py), host-buffer
path, plain resident key and precomputed key.  One JSON line per size.

    python tools/identity_check.py g1:26 g2:22 g2:24
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import legosnark_b200 as lb
from oracle.binding import Checker
from tests import inputs


def main():
    orc = Checker("orc")
    lb.init(1)
    for spec in sys.argv[1:]:
        grp, l2 = spec.split(":")
        n = 1 << int(l2)
        k = inputs.fr_fast_uniform(n, seed=9000 + int(l2))
        s = inputs.fr_fast_uniform(n, seed=9100 + int(l2))
        table = lb.get_window_table(grp, 254, 0, orc.one(grp), expected_scalars=n)
        P = lb.batch_exp(254, 0, table, k)
        table.close()
        idx = np.r_[0:32, n - 32:n, np.random.default_rng(1).integers(0, n, 192)]
        assert (P[idx] == orc.batch_exp(grp, orc.one(grp), k[idx])).all()
        want = inputs.scalar_sum_point(orc, grp, k, s)
        row = {"group": grp, "log2n": int(l2)}
        t0 = time.perf_counter()
        ok_host = bool((lb.multi_exp(grp, P, s) == want).all())
        row["host_buffer_ms"] = (time.perf_counter() - t0) * 1e3
        key = lb.CommitmentKey(grp, P)
        del P
        t0 = time.perf_counter()
        ok_plain = bool((key.multi_exp(s) == want).all())
        row["plain_key_ms"] = (time.perf_counter() - t0) * 1e3
        key.precompute()
        t0 = time.perf_counter()
        ok_pre = bool((key.multi_exp(s) == want).all())
        row["precomputed_key_ms"] = (time.perf_counter() - t0) * 1e3
        st = lb.last_stats()
        row.update(window_bits=st["window_bits"], windows=st["num_windows"], device_ms=st["device_ms"],
                   identity_holds={"host_buffer": ok_host, "plain_key": ok_plain, "precomputed_key": ok_pre})
        key.close()
        print(json.dumps(row), flush=True)
        assert ok_host and ok_plain and ok_pre, spec
    lb.shutdown()


if __name__ == "__main__":
    main()
