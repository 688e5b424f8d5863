#!/usr/bin/env python3
"""tools/e2e_probe.py — the host-buffer call b200_msm_g1 / g2 (pinned Jacobian bases + scalars: the bench's `e2e`) for a grid of
tuning knobs, window sizes and upload-chunk counts.  One JSON line per setting; every result is compared with the first.

    python tools/e2e_probe.py [log2n = 20] [--group g1] [--knobs dense_direct=1,0 window_bits=0,15 chunks=0,3]"""
import argparse, itertools, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import legosnark_b200 as lb
from bench import generator, random_scalars

ap = argparse.ArgumentParser()
ap.add_argument("log2n", nargs="?", type=int, default=20)
ap.add_argument("--group", default="g1")
ap.add_argument("--resident", action="store_true", help="resident precomputed key + pinned host scalars (b200_msm_pinned_*) instead of the cold-key call")
ap.add_argument("--knobs", nargs="*", default=[])
a = ap.parse_args()
n = 1 << a.log2n
lb.init(1)
k = random_scalars(n, 2)
P = torch.from_numpy(lb.batch_exp_once(a.group, generator(a.group), k).view(np.int64)).pin_memory().numpy().view(np.uint64)
s = torch.from_numpy(random_scalars(n, 1).view(np.int64)).pin_memory().numpy().view(np.uint64)
key = None
if a.resident:
    key = lb.CommitmentKey(a.group, bases=P)
    key.precompute(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names, grids = [], []
for kv in a.knobs:
    name, vals = kv.split("=")
    names.append(name)
    grids.append([int(v) for v in vals.split(",")])
ref = None
for combo in itertools.product(*grids) if grids else [()]:
    for name, v in zip(names, combo):
        if name == "window_bits":
            lb.set_tuning(v, 0)
        elif name == "chunks":
            lb.set_pipeline_chunks(v)
        else:
            lb.set_tuning_ex(name, v)
    ts, dev = [], []
    for it in range(10):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = key.multi_exp(s) if key is not None else lb.multi_exp(a.group, P, s)
        ts.append((time.perf_counter() - t0) * 1e3)
        dev.append(lb.last_stats()["device_ms"])
    if ref is None:
        ref = r
    st = lb.last_stats()
    print(json.dumps({"group": a.group, "log2n": a.log2n, "call": "resident key" if a.resident else "cold key", "knobs": dict(zip(names, combo)), "c": st["window_bits"], "W": st["num_windows"],
                      "e2e_ms_median": float(np.median(ts[2:])), "e2e_ms_min": float(np.min(ts[2:])),
                      "device_ms_median": float(np.median(dev[2:])), "launches": st["kernel_launches"], "same": bool((r == ref).all())}),
          flush=True)
lb.set_tuning(0, 0)
lb.set_pipeline_chunks(0)
lb.shutdown()
