#!/usr/bin/env python3
"""tools/e2e_probe.py — the host-buffer call b200_msm_g1 (pinned Jacobian bases + scalars, the bench's `e2e`) under forced
window sizes and upload-chunk counts: is the pipelined plain-key path at its best geometry?  One JSON line per setting."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import legosnark_b200 as lb
from bench import generator, random_scalars

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log2n
lb.init(1)
k = random_scalars(n, 2)
P = torch.from_numpy(lb.batch_exp_once("g1", generator("g1"), k).view(np.int64)).pin_memory().numpy().view(np.uint64)
s = torch.from_numpy(random_scalars(n, 1).view(np.int64)).pin_memory().numpy().view(np.uint64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for even in (1, 0):
  lb.set_tuning_ex("even_chunks", even)
  for chunks in ((0, 3, 4, 5, 6, 8) if not even else (0,)):
    for c in (0,):
        lb.set_tuning(c, 0)
        lb.set_pipeline_chunks(chunks)
        ts = []
        for it in range(8):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = lb.multi_exp("g1", P, s)
            ts.append((time.perf_counter() - t0) * 1e3)
        if ref is None:
            ref = r
        st = lb.last_stats()
        print(json.dumps({"log2n": log2n, "even_chunks": even, "chunks": chunks, "c_forced": c, "c": st["window_bits"], "W": st["num_windows"],
                          "e2e_ms_median": float(np.median(ts[2:])), "e2e_ms_min": float(np.min(ts[2:])), "same": bool((r == ref).all())}), flush=True)
lb.set_tuning(0, 0)
lb.set_pipeline_chunks(0)
lb.shutdown()
