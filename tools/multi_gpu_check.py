"""In-process multi-GPU check (b200_init(G): one host thread per device, index-range shards, host
sum of partials — what the C++ shim uses): MSM, batch_exp and the knowledge-commitment pair must
give bit-identical results on G devices and on 1.  Run on a box with >= 2 GPUs."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import legosnark_b200 as lb
from bench import generator, random_scalars

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 18)
res = {}
for g in (G, 1):
    lb.init(g)
    assert lb.device_count() == g
    k, s = random_scalars(n, 7), random_scalars(n, 8)
    out = {}
    for grp, m in (("g1", n), ("g2", n // 4)):
        P = lb.batch_exp_once(grp, generator(grp), k[:m])
        t0 = time.perf_counter()
        out[grp] = lb.multi_exp(grp, P, s[:m])
        out[grp + "_ms"] = (time.perf_counter() - t0) * 1e3
        out[grp + "_P"] = P
    o2, o1 = lb.kc_multi_exp(out["g2_P"], out["g1_P"][: n // 4], s[: n // 4])
    out["kc"] = np.concatenate([o2, o1])
    # resident key sharded over the devices, plain and with the precomputed window multiples, full range and a sub-range
    key = lb.CommitmentKey("g1", out["g1_P"])
    out["pinned"] = key.multi_exp(s)
    out["pinned_sub"] = key.multi_exp(s[: n // 2 + 5], offset=n // 4)
    key.precompute()
    t0 = time.perf_counter()
    out["pinned_pre"] = key.multi_exp(s)
    out["pre_ms"] = (time.perf_counter() - t0) * 1e3
    out["pinned_pre_sub"] = key.multi_exp(s[: n // 2 + 5], offset=n // 4)
    key.close()
    assert (out["pinned"] == out["g1"]).all() and (out["pinned_pre"] == out["g1"]).all()
    assert (out["pinned_sub"] == out["pinned_pre_sub"]).all()
    res[g] = out
    lb.shutdown()
for key in ("g1", "g2", "kc", "g1_P", "g2_P", "pinned", "pinned_sub", "pinned_pre", "pinned_pre_sub"):
    assert (res[G][key] == res[1][key]).all(), key
print(f"multi-GPU in-process check ok: {G} devices == 1 device on MSM g1 2^{int(np.log2(n))} / g2 / kc / batch_exp; "
      f"host-buffer g1 MSM {res[G]['g1_ms']:.2f} ms on {G} GPUs vs {res[1]['g1_ms']:.2f} ms on 1; "
      f"precomputed resident key (host scalars) {res[G]['pre_ms']:.2f} ms vs {res[1]['pre_ms']:.2f} ms")
