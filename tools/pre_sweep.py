"""Sweep of the precomputed-key geometry on one GPU: window bits c, reduction segment length,
stage-2 split.  Prints one JSON line per configuration (device ms over `reps` runs, L2 flushed)."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import legosnark_b200 as lb
from bench import generator, random_scalars

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--group", default="g1")
ap.add_argument("--cs", default="0,16,17,18,19,20,21")
ap.add_argument("--logS", default="-1")
ap.add_argument("--splits", default="0")
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lb.init_devices([0])
n = 1 << a.log2n
A = 8 if a.group == "g1" else 16
stream = torch.cuda.current_stream().cuda_stream
d_s = torch.from_numpy(random_scalars(n, 1000).view(np.int64)).to(dev)
d_k = torch.from_numpy(random_scalars(n, 2000).view(np.int64)).to(dev)
table = lb.get_window_table(a.group, 254, 0, generator(a.group), expected_scalars=n)
d_aff = torch.empty((n, A), dtype=torch.int64, device=dev)
lb.batch_exp_device(table, d_k.data_ptr(), n, d_aff.data_ptr(), stream)
torch.cuda.synchronize()
table.close()
key = lb.CommitmentKey(a.group, device_affine_ptr=d_aff.data_ptr(), n=n)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def run():
    ms, acc = [], []
    for it in range(a.reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        e0.record()
        out = key.multi_exp_device(d_s.data_ptr(), n, 0, stream)
        e1.record(); e1.synchronize()
        if it >= 2:
            ms.append(e0.elapsed_time(e1)); acc.append(lb.last_stats()["accumulate_ms"])
    return out, float(np.mean(ms)), float(np.min(ms)), float(np.mean(acc)), lb.last_stats()

lb.set_tuning_ex("use_precomputed", 0)
ref, m, mn, acc, st = run()
print(json.dumps({"mode": "plain", "c": st["window_bits"], "W": st["num_windows"], "ms": m, "min_ms": mn, "acc_ms": acc,
                  "device_ms": st["device_ms"], "entries": st["num_entries"]}), flush=True)
lb.set_tuning_ex("use_precomputed", 1)
for c in [int(x) for x in a.cs.split(",")]:
    t0 = time.perf_counter()
    key.precompute(c)
    pre_s = time.perf_counter() - t0
    for ls in [int(x) for x in a.logS.split(",")]:
        for sp in [int(x) for x in a.splits.split(",")]:
            lb.set_tuning_ex("reduce_log_segment", ls)
            lb.set_tuning_ex("reduce_split", sp)
            out, m, mn, acc, st = run()
            print(json.dumps({"mode": "pre", "c_req": c, "c": st["window_bits"], "W": st["num_windows"], "logS": ls, "split": sp,
                              "ms": m, "min_ms": mn, "acc_ms": acc, "device_ms": st["device_ms"], "entries": st["num_entries"],
                              "L": st["chunk_len"], "tasks": st["num_tasks"], "precompute_s": pre_s, "ok": bool((out == ref).all())}), flush=True)
key.close()
lb.shutdown()
