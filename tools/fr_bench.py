"""Secondary measurements for the Fr vector kernels on one GPU:
radix-2 FFT over Fr (device-resident and host-buffer), evalMLE / witness folding, and CPPoly::prove
end to end against a resident key; the reference's own routines (oracle/_ref/liblsref.so) timed beside
them on the host cores.  One JSON line per measurement.  Not the headline bench (bench.py)."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import legosnark_b200 as lb
from bench import generator, random_scalars, measured_peaks

ap = argparse.ArgumentParser()
ap.add_argument("--fft", default="16,20,22,24")
ap.add_argument("--fold", default="20,24")
ap.add_argument("--prove", default="16,20")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--no-cpu", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
lb.init_devices([0])
# a non-default torch stream: the engine enqueues on the stream it is handed (NULL would mean its own),
# and torch's events must sit on that same stream to see the kernels
tstream = torch.cuda.Stream()
torch.cuda.set_stream(tstream)
stream = tstream.cuda_stream
assert stream != 0
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
imad_peak, _ = lb.imad_peak(0, 1 << 14)
hbm_gbs, hbm_src = measured_peaks()
G5 = np.array([0x1b0d0ef99fffffe6, 0xeaba68a3a32a913f, 0x47d8eb76d8dd0689, 0x15d0085520f5bbc3], dtype=np.uint64)  # 5 in Montgomery form
ref = None
if not a.no_cpu:
    from oracle.binding import Checker
    if Checker.available("ref"):
        ref = Checker("ref")

def dev_time(fn, reps):
    ms = []
    for it in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        if it >= 3: ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms)), float(np.min(ms))

def host_time(fn, reps):
    ts = []
    for it in range(reps + 2):
        t0 = time.perf_counter(); r = fn(); dt = time.perf_counter() - t0
        if it >= 2: ts.append(dt * 1e3)
    return float(np.mean(ts)), float(np.min(ts)), r

for d in [int(x) for x in a.fft.split(",") if x]:
    n = 1 << d
    h = random_scalars(n, 10 + d)
    d_a = torch.from_numpy(h.view(np.int64)).to(dev)
    for mode, name in ((0, "FFT"), (3, "icosetFFT")):
        ms, mn = dev_time(lambda: lb.fr_fft_device(d_a.data_ptr(), d, mode, G5, stream), a.reps)
        st = lb.last_stats()
        passes = st["num_windows"]
        modmul = n / 2 * d + (n if mode == 3 else 0)
        bytes_ = passes * 2 * n * 32 + n / 2 * d * 32 * 0  # twiddles are L2-resident tables (n/2 x 32 B)
        out = {"what": "fr_fft_device", "mode": name, "log2n": d, "ms": ms, "min_ms": mn, "passes": passes, "launches": st["kernel_launches"],
               "elements_per_s": n / (ms * 1e-3),
               "roofline": {"bound": "imad", "achieved": modmul * 136 / (ms * 1e-3) / 1e12, "peak": imad_peak / 1e12,
                            "frac": modmul * 136 / (ms * 1e-3) / imad_peak, "unit": "T multiply-add/s",
                            "algorithmic": "n/2 log n butterflies x 1 modmul x 136 multiply-adds",
                            "hbm": {"bytes": bytes_, "achieved_gbs": bytes_ / (ms * 1e-3) / 1e9, "peak_gbs": hbm_gbs, "frac": bytes_ / (ms * 1e-3) / 1e9 / hbm_gbs}}}
        print(json.dumps(out), flush=True)
    if d <= 22:
        ms, mn, res = host_time(lambda: lb.fr_fft(h, 0), max(3, a.reps // 2))
        line = {"what": "fr_fft_host_buffers", "mode": "FFT", "log2n": d, "ms": ms, "min_ms": mn, "h2d_bytes": n * 32, "d2h_bytes": n * 32}
        if ref is not None and d <= 22:
            t0 = time.perf_counter(); want = ref.fr_fft(h, 0); dt = (time.perf_counter() - t0) * 1e3
            line["cpu_reference_ms"] = dt
            line["cpu_reference"] = f"libfqfft basic_radix2_domain::FFT, MULTICORE, {ref.max_threads()} threads (includes the wrapper's vector copies)"
            line["bit_identical"] = bool((want == res).all())
        print(json.dumps(line), flush=True)
    del d_a

for d in [int(x) for x in a.fold.split(",") if x]:
    n = 1 << d
    v = random_scalars(n, 20 + d); r = random_scalars(d, 30 + d)
    ms, mn, ev = host_time(lambda: lb.evalMLE(v, r), max(3, a.reps // 2))
    line = {"what": "evalMLE_host_buffers", "log2n": d, "ms": ms, "min_ms": mn, "h2d_bytes": n * 32, "launches": lb.last_stats()["kernel_launches"]}
    ms2, mn2, _ = host_time(lambda: lb.fold_witness(v, r), 3)
    line["fold_witness_ms"] = ms2
    if ref is not None and d <= 22:
        t0 = time.perf_counter(); want = ref.fr_eval_mle(v, r); dt = (time.perf_counter() - t0) * 1e3
        line["cpu_reference_ms"] = dt
        line["cpu_reference"] = "MultiVPolyT::evalMLE (single thread as shipped)"
        line["bit_identical"] = bool((want == ev).all())
    print(json.dumps(line), flush=True)

for d in [int(x) for x in a.prove.split(",") if x]:
    n = 1 << d
    m = n // 2
    d_k = torch.from_numpy(random_scalars(m, 40 + d).view(np.int64)).to(dev)
    table = lb.get_window_table("g1", 254, 0, generator("g1"), expected_scalars=m)
    d_aff = torch.empty((m, 8), dtype=torch.int64, device=dev)
    lb.batch_exp_device(table, d_k.data_ptr(), m, d_aff.data_ptr(), stream)
    torch.cuda.synchronize(); table.close()
    key = lb.CommitmentKey("g1", device_affine_ptr=d_aff.data_ptr(), n=m)
    v = random_scalars(n, 50 + d); r = random_scalars(d, 60 + d)
    ms, mn, (wit, ev) = host_time(lambda: lb.cppoly_prove(key, v, r), 5)
    st = lb.last_stats()
    # the same proof through the per-call boundary: fold on the host side of the C-ABI, one pinned MSM per level
    def per_level():
        w, _ = lb.fold_witness(v, r)
        out = []
        for i in range(d):
            mm = 1 << (d - i - 1); s0 = n - 2 * mm
            out.append(key.multi_exp(w[s0:s0 + mm]))
        return np.stack(out)
    ms3, mn3, wit3 = host_time(per_level, 3)
    t0 = time.perf_counter(); key.precompute(); pre_ms = (time.perf_counter() - t0) * 1e3
    ms2, mn2, (wit2, _) = host_time(lambda: lb.cppoly_prove(key, v, r), 5)
    print(json.dumps({"what": "cppoly_prove_g1", "l": d, "ms": ms, "min_ms": mn, "launches": st["kernel_launches"], "msm_device_ms": st["device_ms"],
                      "ms_precomputed_key": ms2, "key_precompute_ms_one_off": pre_ms, "ms_per_level_calls": ms3,
                      "same_witness": bool((wit == wit2).all() and (wit == wit3).all()), "h2d_bytes": (n + d) * 32}), flush=True)
    key.close(); del d_aff, d_k
lb.shutdown()
