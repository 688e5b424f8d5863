"""Pageable host <-> device transfer probe: time b200_fr_fft through host buffers (in + out = 2 x 32 B x 2^log2n)
and a resident-key pin of 2^log2n bases for B200_COPY_THREADS as given in the environment."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import legosnark_b200 as lb
from bench import random_scalars, generator
d = int(sys.argv[1]) if len(sys.argv) > 1 else 22
lb.init(1)
n = 1 << d
a = random_scalars(n, 1)
lb.fr_fft(a, 0)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); lb.fr_fft(a, 0); ts.append((time.perf_counter() - t0) * 1e3)
P = lb.batch_exp_once("g1", generator("g1"), random_scalars(1 << 20, 2))
tp = []
for _ in range(3):
    t0 = time.perf_counter(); k = lb.CommitmentKey("g1", P); tp.append((time.perf_counter() - t0) * 1e3); k.close()
s = random_scalars(1 << 20, 3)
tm = []
for _ in range(4):
    t0 = time.perf_counter(); lb.multi_exp("g1", P, s); tm.append((time.perf_counter() - t0) * 1e3)
print(f"B200_COPY_THREADS={os.environ.get('B200_COPY_THREADS','(default 4)')}: fft 2^{d} host buffers {min(ts):.2f} ms "
      f"({2*n*32/min(ts)/1e6:.1f} GB/s incl. transform) | pin 2^20 G1 bases (100 MB) {min(tp):.2f} ms | cold host-buffer MSM 2^20 (134 MB up) {min(tm):.2f} ms")
lb.shutdown()
