"""G1 MSM time by scalar distribution (uniform, 32-bit, 0/1-heavy, all-equal scalars)
on one GPU at n = 2^log2n, resident key: plain / precomputed key, ones filter on / off.  The result of
every variant must agree with the first one.  One JSON line per (distribution, variant)."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import legosnark_b200 as lb
from bench import generator, random_scalars, R_ORDER, MONT_R, limbs

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--reps", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lb.init_devices([0])
n = 1 << a.log2n
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); stream = ts.cuda_stream
d_k = torch.from_numpy(random_scalars(n, 2000).view(np.int64)).to(dev)
table = lb.get_window_table("g1", 254, 0, generator("g1"), expected_scalars=n)
d_aff = torch.empty((n, 8), dtype=torch.int64, device=dev)
lb.batch_exp_device(table, d_k.data_ptr(), n, d_aff.data_ptr(), stream)
torch.cuda.synchronize(); table.close()
key = lb.CommitmentKey("g1", device_affine_ptr=d_aff.data_ptr(), n=n)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ONE = limbs(MONT_R % R_ORDER)
rng = np.random.default_rng(7)

def dist(name):
    if name == "uniform":
        return random_scalars(n, 1)
    if name == "32-bit":   # rand32b, legogrothmatrix.cc:29-32: small integers, stored in Montgomery form
        v = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
        out = np.zeros((n, 4), dtype=np.uint64)
        out[:, 0] = v  # standard-form small integers ...
        return lb.test_field_op(1, 7, out)  # ... to Montgomery form on the device (op 7 = from bigint)
    if name == "zero-one-heavy":  # 90 % of the scalars in {0, 1}
        s = random_scalars(n, 3)
        u = rng.random(n)
        s[u < 0.45] = 0
        s[(u >= 0.45) & (u < 0.9)] = ONE
        return s
    if name == "all-equal":
        return np.tile(random_scalars(1, 5), (n, 1))
    raise ValueError(name)

def run(d_s):
    ms = []
    for it in range(a.reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_(); e0.record()
        out = key.multi_exp_device(d_s.data_ptr(), n, 0, stream)
        e1.record(); e1.synchronize()
        if it >= 2: ms.append(e0.elapsed_time(e1))
    return out, float(np.mean(ms)), lb.last_stats()

pre_done = False
for name in ("uniform", "32-bit", "zero-one-heavy", "all-equal"):
    d_s = torch.from_numpy(dist(name).view(np.int64)).to(dev)
    ref = None
    for pre in (0, 1):
        if pre and not pre_done:
            key.precompute(); pre_done = True
        lb.set_tuning_ex("use_precomputed", pre)
        for ones in (1, 0):
            lb.set_tuning_ex("ones_filter", ones)
            out, ms, st = run(d_s)
            if ref is None: ref = out
            print(json.dumps({"dist": name, "log2n": a.log2n, "precomputed_key": bool(pre), "ones_filter": bool(ones), "ms": ms,
                              "points_per_s": n / (ms * 1e-3), "c": st["window_bits"], "W": st["num_windows"], "entries": st["num_entries"],
                              "acc_ms": st["accumulate_ms"], "agrees": bool((out == ref).all())}), flush=True)
    del d_s
lb.set_tuning_ex("ones_filter", 1); lb.set_tuning_ex("use_precomputed", 1)
key.close(); lb.shutdown()
