#!/usr/bin/env python3
"""tools/stage_times.py — A/B of the engine's tuning knobs on one GPU: device time of one MSM on a resident key
(CUDA events inside the engine: whole pipeline, sort phase, k_accumulate) for a grid of knob settings.

    python tools/stage_times.py --log2n 20 [--group g1] [--plain] [--knobs partition_sort=0,1 reduce_marginals=0,1 reduce_log_segment=-1,2,3,4]

Every setting is checked against the first one's result (same group element).  One JSON line per setting."""
import argparse, itertools, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import legosnark_b200 as lb
from bench import generator, random_scalars


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default="g1")
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--plain", action="store_true", help="plain resident key (default: precomputed)")
    ap.add_argument("--precompute-bits", type=int, default=0)
    ap.add_argument("--reps", type=int, default=9)
    ap.add_argument("--knobs", nargs="*", default=[])
    a = ap.parse_args()
    n = 1 << a.log2n
    lb.init(1)
    dev = torch.device("cuda", 0)
    s = torch.from_numpy(random_scalars(n, 1).view(np.int64)).to(dev)
    k = random_scalars(n, 2)
    P = lb.batch_exp_once(a.group, generator(a.group), k)
    key = lb.CommitmentKey(a.group, bases=P)
    del P
    if not a.plain:
        key.precompute(a.precompute_bits)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    names, grids = [], []
    for kv in a.knobs:
        name, vals = kv.split("=")
        names.append(name)
        grids.append([int(v) for v in vals.split(",")])
    ref = None
    stream = torch.cuda.current_stream().cuda_stream
    for combo in itertools.product(*grids) if grids else [()]:
        for name, v in zip(names, combo):
            if name == "window_bits":
                lb.set_tuning(v, 0)
            elif name == "task_len":
                lb.set_tuning(0, v)
            else:
                lb.set_tuning_ex(name, v)
        rows = []
        for it in range(a.reps + 2):
            flush.zero_()
            torch.cuda.synchronize()
            r = key.multi_exp_device(s.data_ptr(), n, 0, stream)
            st = lb.last_stats()
            if it >= 2:
                rows.append((st["device_ms"], st["sort_ms"], st["accumulate_ms"]))
        if ref is None:
            ref = r
        ok = bool((r == ref).all())
        rows = np.array(rows)
        med = np.median(rows, axis=0)
        print(json.dumps({"group": a.group, "log2n": a.log2n, "key": "plain" if a.plain else "precomputed",
                          "knobs": dict(zip(names, combo)), "device_ms": float(med[0]), "sort_ms": float(med[1]),
                          "accumulate_ms": float(med[2]), "rest_ms": float(med[0] - med[1] - med[2]),
                          "c": st["window_bits"], "W": st["num_windows"], "L": st["chunk_len"], "launches": st["kernel_launches"],
                          "points_per_s": n / (med[0] * 1e-3), "same_result": ok}), flush=True)
    key.close()
    lb.shutdown()


if __name__ == "__main__":
    main()
