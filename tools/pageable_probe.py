#!/usr/bin/env python3
"""tools/pageable_probe.py — b200_msm_g1 from PAGEABLE host buffers (a libff caller's std::vector) for the staging
thread count in B200_COPY_THREADS (read once per process: run once per value).  One JSON line."""
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
P = np.array(lb.batch_exp_once("g1", generator("g1"), random_scalars(n, 2)))
s = np.array(random_scalars(n, 1))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for it in range(10):
    flush.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = lb.multi_exp("g1", P, s)
    ts.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({"log2n": log2n, "copy_threads": os.environ.get("B200_COPY_THREADS", "default(4)"),
                  "copy_chunk_kb": os.environ.get("B200_COPY_CHUNK_KB", "default(4096)"),
                  "e2e_pageable_ms_median": float(np.median(ts[3:])), "min": float(np.min(ts[3:]))}), flush=True)
lb.shutdown()
