"""Host-buffer (cold key) G1 MSM at 2^log2n from pinned memory for several upload chunk counts."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import legosnark_b200 as lb
from bench import generator, random_scalars
d = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << d
lb.init(1)
P = lb.batch_exp_once("g1", generator("g1"), random_scalars(n, 2))
s = random_scalars(n, 3)
Pp = torch.from_numpy(P.view(np.int64)).pin_memory(); sp = torch.from_numpy(s.view(np.int64)).pin_memory()
Pn, sn = Pp.numpy().view(np.uint64), sp.numpy().view(np.uint64)
ref = None
for ch in (0, 2, 3, 4, 5, 6, 8, 12):
    lb.set_pipeline_chunks(ch)
    ts = []
    for it in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = lb.multi_exp("g1", Pn, sn); ts.append((time.perf_counter() - t0) * 1e3)
    if ref is None: ref = out
    print(f"chunks={ch}: {np.mean(ts[2:]):.3f} ms (min {min(ts):.3f}) ok={bool((out == ref).all())}", flush=True)
lb.shutdown()
