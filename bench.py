#!/usr/bin/env python3
"""bench.py — G1 MSM points/s on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log2n 20] [--group g1]
    python bench.py --impl reference ...      # the reference's CPU multi_exp on the host cores

A "step" is one multi-scalar multiplication of n = 2^log2n synthetic points per GPU
(config 2 of BASELINE.json at the default 2^20): uniform random scalars, bases k_i*G made on
the device by the fixed-base kernel.  Numbers on the JSON line:

  value   whole-job points/s with bases resident in HBM (CommitmentKey) and scalars already
          on the device; timed with CUDA events on the launching stream around the C-ABI
          call, which includes the D2H of the W window sums and the host Horner tail.
  e2e     the same MSM through the reference-facing host call b200_msm_g1 (== libff::multi_exp
          signature: host Jacobian bases + host scalars, here in pinned memory): H2D of
          bases and scalars, Jacobian->affine ingest, MSM, D2H, host tail, all inside the
          timed region.
  roofline  k_accumulate (the bucket mixed-addition kernel, >80 % of the step) against the
          MEASURED IMAD.WIDE.U32 peak of this GPU (b200_imad_peak), canonical count of
          SURVEY.md §8(d): 11 modmul per mixed add x 136 multiply-adds per modmul; plus
          the gather traffic against the measured HBM copy bandwidth.
  cpu_baseline  libff's multi_exp<BDLO12> (oracle/_ref, built from the unmodified reference)
          on this box's host cores over a bounded sample of the same workload.

  e2e_pageable  as e2e, but from plain (pageable) numpy buffers — what a libff caller's std::vector is.

With N > 1 (torchrun) every rank owns the index range of an N * 2^log2n-point MSM (weak scaling, the
default) or of ONE 2^log2n-point MSM (--scaling strong: the reference's own split, multiexp.tcc:417-438):
it runs the whole pipeline on its slice, the 96-byte partials go through a page of host shared memory
(legosnark_b200.multi.HostMailbox) and are summed on the host — no collective, no NCCL in the data path
(torch.distributed only lines the ranks up before a step and takes the max of the timings afterwards).

--impl reference times libff's multi_exp<BDLO12> (oracle/_ref = the unmodified reference sources) with
chunks = all host threads on the SAME workload (2^log2n points per step by default), trying both of the
reference's curve builds (alt_bn128 + USE_ASM, and bn128 = ate-pairing's Xbyak JIT, LegoSNARK's default
CURVE) and the -march=native build when it runs on this host; the faster combination is the one timed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MONT_R = 1 << 256
METRIC = "g1_msm_points_per_s"
UNIT = "points/s"


def limbs(x, n=4):
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)], dtype=np.uint64)


def generator(group):
    """G1_one = (1, 2, 1) / G2_one (alt_bn128_init.cpp:148-150, 209-213), Montgomery limbs."""
    m = lambda v: limbs(v * MONT_R % Q)
    if group == "g1":
        return np.concatenate([m(1), m(2), m(1)])
    xc0 = 10857046999023057135944570762232829481370756359578518086990519993285655852781
    xc1 = 11559732032986387107991004021392285783925812861821192530917403151452391805634
    yc0 = 8495653923123431417604973247489272438418190587263600148770280649306958101930
    yc1 = 4082367875863433681332203403145435568316851327593401208105741076214120093531
    return np.concatenate([m(xc0), m(xc1), m(yc0), m(yc1), m(1), m(0)])


def random_scalars(n, seed):
    """Uniform 253-bit integers read as Montgomery images (every residue < r is one)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 61) - 1)
    return a


def workload_config(args, world):
    """The `config` object: identical for the GPU arm and the reference arm of one (group, log2n, scaling, N)."""
    per = f"2^{args.log2n} points per GPU" if args.scaling == "weak" else f"one MSM of 2^{args.log2n} points split over the GPUs by index range"
    return {"workload": f"{args.group} MSM, {per} (BASELINE.json configs[1] at 2^20; configs[4] for the other sizes / G2)",
            "group": args.group, "log2n": args.log2n, "n_gpus": world, "scaling": args.scaling,
            "scalars": "uniform 253-bit", "bases": "k_i*G, distinct",
            "key": "GPU arm `value`: precomputed resident key (window multiples 2^(c k) P_i kept in HBM, one-off cost reported in "
                   "plain_key.key_precompute_ms_one_off); `plain_key`: plain resident affine key; `e2e`: cold host bases + scalars "
                   "through b200_msm_*; reference arm: libff multi_exp<BDLO12> on host vectors",
            "l2": "GPU arm: flushed between timed iterations (512 MiB write, completed before the timed region starts)",
            "sharding": "index range per rank, partials summed on the host (no collective)"}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if not (t0 <= ts <= t1):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(group, log2n, pre=False):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    try:
        e = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[f"{group}_2p{log2n}" + ("_pre" if pre else "")]
        return float(e["dram_read_bytes"]) + float(e["dram_write_bytes"])
    except Exception:
        return None


def checker(native=False):
    from oracle.binding import Checker
    if Checker.available("ref"):
        try:
            return Checker("ref", native=native), "reference"
        except Exception:
            pass
    return Checker("orc"), "port"


def run_reference(args, rank, world):
    """--impl reference: libff multi_exp<BDLO12> with chunks = host threads on the arm's workload."""
    if rank != 0:
        return
    from oracle.binding import Checker
    chk, kind = checker(native=True)
    group = args.group
    m = 1 << min(args.log2n, args.ref_log2n)
    k = chk.sha512_rng_fr(1 << 40, m)
    P = chk.batch_exp(group, chk.one(group), k, normalise=True)
    s = chk.sha512_rng_fr(0, m)
    threads = chk.max_threads()
    # which of the reference's builds is faster here: alt_bn128 (x86-64 asm Fp) or bn128 (Xbyak JIT Fp, LegoSNARK's default
    # CURVE); one probe each on a 2^16 prefix, then every timed step runs the winner
    probes = {}
    if kind == "reference":
        mp_ = min(m, 1 << 16)
        for name, curve in (("alt_bn128+USE_ASM", 0), ("bn128 (Xbyak JIT)", 1)):
            chk.msm(group, P[:mp_], s[:mp_], chunks=threads, variant=0, curve=curve)
            t0 = time.perf_counter()
            chk.msm(group, P[:mp_], s[:mp_], chunks=threads, variant=0, curve=curve)
            probes[name] = mp_ / (time.perf_counter() - t0)
        best = max(probes, key=probes.get)
        curve = 0 if best.startswith("alt") else 1
    else:
        best, curve = "C restatement of alt_bn128", 0
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if kind == "reference":
            chk.msm(group, P, s, chunks=threads, variant=0, curve=curve)
        else:
            chk.msm(group, P, s, chunks=threads, variant=0)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = m / (ms * 1e-3)
    sample = (f"{group} multi_exp<BDLO12> of 2^{int(np.log2(m))} points per step, chunks={threads}, build: {best}"
              f"{', -march=native' if getattr(chk, 'native', False) else ''}; probes (points/s at 2^16): "
              + ", ".join(f"{a}: {b:.3g}" for a, b in probes.items()))
    print(json.dumps({
        "impl": "reference", "metric": METRIC.replace("g1", group), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u64 limbs (254-bit modular integer)", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=20, help="points per GPU = 2^log2n")
    ap.add_argument("--group", default="g1", choices=["g1", "g2"])
    ap.add_argument("--ref-log2n", type=int, default=20, help="cap on the size of one step of the reference arm (2^20 = 1.5 s per step)")
    ap.add_argument("--cpu-sample-log2n", type=int, default=18, help="sample of the in-arm cpu_baseline / parity leg")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 2^log2n points per GPU; strong: one 2^log2n-point MSM split over the GPUs")
    ap.add_argument("--transport", default="mailbox", choices=["mailbox", "nccl"], help="how the 96-byte partials travel")
    ap.add_argument("--no-cplink", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-precompute", action="store_true",
                    help="headline on the plain resident key (default: key extended by b200_key_precompute_*)")
    ap.add_argument("--precompute-bits", type=int, default=0, help="window bits of the precomputed key (0 = engine's choice)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1 to every rank when N > 1; the reference arm is rank 0 alone
        # with all the host threads it can use, so undo that default before libgomp reads it
        if os.environ.get("OMP_NUM_THREADS") == "1" and (world > 1 or "TORCHELASTIC_RUN_ID" in os.environ):
            os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import legosnark_b200 as lb
    from legosnark_b200 import multi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lb.init_devices([local_rank])

    group = args.group
    if args.scaling == "weak":
        n = 1 << args.log2n
    else:  # strong: this rank's index range of ONE 2^log2n-point MSM (multiexp.tcc:417-431: the last chunk takes the remainder)
        lo_, hi_ = lb.shard_range(1 << args.log2n, rank, world)
        n = hi_ - lo_
    n_total = world * n if args.scaling == "weak" else 1 << args.log2n
    mailbox = None
    if world > 1 and args.transport == "mailbox":
        # one node: the ranks share /dev/shm.  If it cannot be used (read-only, missing) every rank falls back to the
        # torch.distributed transport together (the flag is agreed on with an all_reduce).
        name = f"b200_partials_{os.environ.get('MASTER_PORT', '0')}_{os.environ.get('TORCHELASTIC_RUN_ID', 'run')}"
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            if rank == 0:
                mailbox = multi.HostMailbox(name, rank, world, create=True)
        except OSError:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok[0]) and rank != 0:
            try:
                mailbox = multi.HostMailbox(name, rank, world, create=False)
            except OSError:
                ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok[0]):
            mailbox = None
    L = 12 if group == "g1" else 24
    A = 8 if group == "g1" else 16
    stream = torch.cuda.current_stream().cuda_stream

    # ---- synthetic inputs: scalars on host (pinned) and device, bases k_i * G made on the device ----
    s_host = torch.from_numpy(random_scalars(n, 1000 + rank).view(np.int64)).pin_memory()
    k_host = torch.from_numpy(random_scalars(n, 2000 + rank).view(np.int64))
    d_s = s_host.to(dev)
    d_k = k_host.to(dev)
    table = lb.get_window_table(group, 254, 0, generator(group), expected_scalars=n)
    d_aff = torch.empty((n, A), dtype=torch.int64, device=dev)
    lb.batch_exp_device(table, d_k.data_ptr(), n, d_aff.data_ptr(), stream)
    torch.cuda.synchronize()
    table.close()
    key = lb.CommitmentKey(group, device_affine_ptr=d_aff.data_ptr(), n=n)
    # host image of the same bases in the reference layout (Jacobian, Z = 1), pinned
    jac_host = torch.empty((n, L), dtype=torch.int64).pin_memory()
    jac_host[:, :A] = d_aff.cpu()
    one_z = torch.from_numpy(generator(group)[A:].view(np.int64).copy())
    jac_host[:, A:] = one_z
    del d_k, d_aff
    s_np = s_host.numpy().view(np.uint64)
    jac_np = jac_host.numpy().view(np.uint64)
    # the same inputs in pageable memory (what a libff caller's std::vector is)
    s_page, jac_page = np.array(s_np), np.array(jac_np)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def align():
        """The L2 flush is test hygiene, not part of a step: wait for it (it would otherwise run next to the first
        kernels of the step and take their bandwidth) and line the ranks up after it, so that the gather of the
        partials inside the timed region does not wait for a peer that is still flushing."""
        torch.cuda.synchronize()  # the flush runs on torch's stream, the engine on its own: do not let them overlap
        if world > 1:
            dist.barrier()

    def step_resident():
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        align()
        ev0.record()
        part = key.multi_exp_device(d_s.data_ptr(), n, 0, stream)
        ev1.record()
        ev1.synchronize()
        dev_ms = ev0.elapsed_time(ev1)
        t0 = time.perf_counter()
        res = multi.sharded_multi_exp(group, part, dev, mailbox) if world > 1 else part
        gather_ms = (time.perf_counter() - t0) * 1e3 if world > 1 else 0.0
        return res, dev_ms + gather_ms, lb.last_stats()

    def step_e2e_resident():
        """Host scalars (pinned) against the resident key: what a LegoSNARK commit pays once its key is on the device."""
        flush.zero_()
        torch.cuda.synchronize()
        align()
        t0 = time.perf_counter()
        part = key.multi_exp(s_np)
        res = multi.sharded_multi_exp(group, part, dev, mailbox) if world > 1 else part
        return res, (time.perf_counter() - t0) * 1e3, lb.last_stats()

    def step_e2e(bases=None, scalars=None):
        bases = jac_np if bases is None else bases
        scalars = s_np if scalars is None else scalars
        flush.zero_()
        torch.cuda.synchronize()
        align()
        t0 = time.perf_counter()
        part = lb.multi_exp(group, bases, scalars)
        res = multi.sharded_multi_exp(group, part, dev, mailbox) if world > 1 else part
        return res, (time.perf_counter() - t0) * 1e3, lb.last_stats()

    # ---- measured denominators -------------------------------------------------------
    imad_wide_peak, _ = lb.imad_peak(0, 1 << 14)
    imad32_peak, _ = lb.imad_peak(1, 1 << 14)
    modmul_peak, _ = lb.imad_peak(2, 1 << 9)
    hbm_gbs, hbm_src = measured_peaks()

    # ---- plain resident key first (secondary number), then the key is extended by its window
    # multiples (one-off, timed) and the headline steps run on the precomputed key -------------
    plain = None
    if not args.no_precompute:
        for _ in range(args.warmup):
            step_resident()
        barrier()
        pl_ms, pl_acc = [], []
        for _ in range(args.steps):
            res_plain, ms, stp = step_resident()
            pl_ms.append(ms)
            pl_acc.append(stp["accumulate_ms"])
        barrier()
        tot_pl = float(np.sum(pl_ms))
        if world > 1:
            t = torch.tensor([tot_pl], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot_pl = float(t[0])
        t0 = time.perf_counter()
        key.precompute(args.precompute_bits)
        pre_ms = (time.perf_counter() - t0) * 1e3
        plain = {"value": n_total / (tot_pl / args.steps * 1e-3), "unit": UNIT, "ms_per_step": tot_pl / args.steps,
                 "window_bits": stp["window_bits"], "windows": stp["num_windows"], "k_accumulate_ms": float(np.mean(pl_acc)),
                 "mixed_adds": stp["num_entries"], "key_precompute_ms_one_off": pre_ms}

    # ---- warm-up, then exactly K timed steps of each flavour ----------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    time.sleep(0.25)
    t_start = time.perf_counter()
    res_ms, acc_ms, fin_us, sort_ms = [], [], [], []
    result = None
    for _ in range(args.steps):
        result, ms, st = step_resident()
        res_ms.append(ms)
        acc_ms.append(st["accumulate_ms"])
        fin_us.append(st["host_finalize_us"])
        sort_ms.append(st["sort_ms"])
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop(t_start, t_end)
    stats = st

    for _ in range(min(args.warmup, 3)):
        step_e2e()
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        res2, ms, st2 = step_e2e()
        e2e_ms.append(ms)
    barrier()
    assert (res2 == result).all(), "host-buffer path and resident path disagree"
    for _ in range(min(args.warmup, 3)):
        step_e2e_resident()
    barrier()
    e2r_ms = []
    for _ in range(args.steps):
        res3, ms, st3 = step_e2e_resident()
        e2r_ms.append(ms)
    barrier()
    assert (res3 == result).all(), "resident-key host-scalar path disagrees"
    for _ in range(min(args.warmup, 3)):
        step_e2e(jac_page, s_page)
    barrier()
    e2p_ms = []
    for _ in range(args.steps):
        res4, ms, st4 = step_e2e(jac_page, s_page)
        e2p_ms.append(ms)
    barrier()
    assert (res4 == result).all(), "pageable host-buffer path disagrees"
    if plain is not None:
        assert (res_plain == result).all(), "precomputed key and plain key disagree"

    tot_res, tot_e2e, tot_e2r, tot_e2p = float(np.sum(res_ms)), float(np.sum(e2e_ms)), float(np.sum(e2r_ms)), float(np.sum(e2p_ms))
    if world > 1:
        t = torch.tensor([tot_res, tot_e2e, tot_e2r, tot_e2p], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_res, tot_e2e, tot_e2r, tot_e2p = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    ms_per_step = tot_res / args.steps
    value = n_total / (ms_per_step * 1e-3)
    e2e_value = n_total / (tot_e2e / args.steps * 1e-3)

    # ---- roofline of the dominant kernel ---------------------------------------------
    acc = float(np.mean(acc_ms))
    entries = stats["num_entries"]
    MODMUL_MIXED = {"g1": 11, "g2": 29}[group]      # canonical madd-2007-bl count, SURVEY.md §8(d)
    MODMUL_ACTUAL = {"g1": 10, "g2": 28}[group]     # XYZZ madd-2008-s: 8M+2S (Fq2: 8*3 + 2*2)
    imads = entries * MODMUL_MIXED * 136.0
    achieved = imads / (acc * 1e-3) / 1e12
    gather_bytes = entries * ({"g1": 64, "g2": 128}[group] + 4.0)
    roofline = {
        "bound": "imad", "kernel": f"k_accumulate<{'Fq' if group == 'g1' else 'Fq2'}>",
        "achieved": achieved, "peak": imad_wide_peak / 1e12, "unit": "T multiply-add/s (32x32+64, lane-ops)",
        "frac": achieved / (imad_wide_peak / 1e12), "traffic": ncu_traffic(group, args.log2n, plain is not None),
        "peak_source": "measured live on this GPU: b200_imad_peak(0) = the IMAD.WIDE.U32[.X] carry-row stream of the "
                       "Montgomery product on all SMs (hardware issue limit: 32 wide multiply-adds/clk/SM = "
                       f"{32 * 148 * 1.965e9 / 1e12:.2f} T/s at 1965 MHz, ncu sm__pipe_fmaheavy_cycles_active)",
        "ms_per_launch": acc, "share_of_step": acc / (float(np.mean(res_ms))),
        "algorithmic": {"mixed_adds_per_launch": entries, "modmul_per_mixed_add": MODMUL_MIXED,
                        "imad_per_modmul": 136, "modmul_per_mixed_add_executed": MODMUL_ACTUAL},
        "modmul_per_s": entries * MODMUL_ACTUAL / (acc * 1e-3), "modmul_per_s_peak_measured": modmul_peak,
        "imad32_peak": imad32_peak / 1e12,
        # bucket sort (radix partition staged through shared memory, sort_kernels.cuh): scalars read once, W (entry, bucket)
        # pairs moved once as 8 B and W bucket-ordered entries written as 4 B; north star: "HBM GB/s for the sort and gather phases"
        "hbm_sort": {"bytes_per_step": n * 32.0 + 3.0 * 4.0 * stats["num_windows"] * n, "ms": float(np.mean(sort_ms)),
                     "achieved_gbs": (n * 32.0 + 12.0 * stats["num_windows"] * n) / (float(np.mean(sort_ms)) * 1e-3) / 1e9,
                     "peak_gbs": hbm_gbs, "frac": (n * 32.0 + 12.0 * stats["num_windows"] * n) / (float(np.mean(sort_ms)) * 1e-3) / 1e9 / hbm_gbs,
                     "note": "instruction- and latency-bound (digit recoding, shared-memory ranks: ncu sm throughput 52-65 %), not bandwidth-bound; "
                             "six launches: k_part_hist, k_part_scan, k_part_scatter, k_part_sort, k_part_scan, k_task_emit"},
        "hbm": {"gather_bytes_per_launch": gather_bytes, "achieved_gbs": gather_bytes / (acc * 1e-3) / 1e9,
                "peak_gbs": hbm_gbs, "peak_source": hbm_src, "frac": gather_bytes / (acc * 1e-3) / 1e9 / hbm_gbs},
    }

    # ---- CPU baseline beside it (rank 0, N = 1), with a parity check on the sample ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        chk, kind = checker()
        m = min(n, 1 << args.cpu_sample_log2n)
        threads = chk.max_threads()
        t0 = time.perf_counter()
        want = chk.msm(group, jac_np[:m], s_np[:m], chunks=threads, variant=0)
        dt = time.perf_counter() - t0
        got = key.multi_exp(s_np[:m])
        assert (got == want).all(), "GPU result differs from the CPU reference on the baseline sample"
        cpu = {"value": m / dt, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"first {m} points of the workload, multi_exp<BDLO12> chunks={threads}, {dt:.2f} s; "
                         "GPU result on the same sample bit-identical"}

    # ---- the other half of BASELINE.json's metric: cplink prove ms, from the end-to-end driver over the
    # reference's own CPlink classes (integration/cplink_driver.cc; prebuilt binaries travel with the
    # snapshot).  Run after the engine of this process is shut down; rank 0, N = 1 only.
    key.close()
    lb.shutdown()
    cplink = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_cplink:
        cplink = cplink_prove_ms()
    if mailbox is not None:
        mailbox.close(unlink=rank == 0)

    if rank == 0:
        print(json.dumps({
            "metric": METRIC.replace("g1", group), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u32 limbs (254-bit modular integer, 8x32-bit Montgomery)", "data": "synthetic",
            "config": workload_config(args, world),
            "engine": {"points_this_rank": n, "points_total": n_total, "window_bits": stats["window_bits"], "windows": stats["num_windows"],
                       "key": ("plain affine key" if plain is None else
                               f"precomputed key: {stats['num_windows']} levels x 64 B per base resident in HBM, all windows share one bucket set"),
                       "partials": ("host shared memory (HostMailbox)" if mailbox is not None else "n/a (one rank)" if world == 1 else "NCCL all_gather")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": stats2_bytes(st2, "h2d_bytes") * world,
                    "d2h_bytes_per_step": stats2_bytes(st2, "d2h_bytes") * world, "ms_per_step": tot_e2e / args.steps,
                    "path": "b200_msm_* with pinned host Jacobian bases + scalars (cold key)"},
            "e2e_pageable": {"value": n_total / (tot_e2p / args.steps * 1e-3), "unit": UNIT, "ms_per_step": tot_e2p / args.steps,
                             "h2d_bytes_per_step": stats2_bytes(st4, "h2d_bytes") * world, "d2h_bytes_per_step": stats2_bytes(st4, "d2h_bytes") * world,
                             "path": "b200_msm_* from pageable host buffers (a libff caller's std::vector): staged through pinned chunks inside the call"},
            "e2e_resident_key": {"value": n_total / (tot_e2r / args.steps * 1e-3), "unit": UNIT, "ms_per_step": tot_e2r / args.steps,
                                 "h2d_bytes_per_step": float(st3["h2d_bytes"]) * world, "d2h_bytes_per_step": float(st3["d2h_bytes"]) * world,
                                 "path": "b200_msm_pinned_*: pinned host scalars against the key resident in HBM (a commitment under a fixed key)"},
            "gpu_launches": int(stats["kernel_launches"]) * args.steps,
            "launches_per_step": int(stats["kernel_launches"]),
            "host_finalize_us": float(np.mean(fin_us)),
            "roofline": roofline, "cpu_baseline": cpu, "plain_key": plain, "cplink": cplink,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cplink_prove_ms():
    """src/examples/cplink at its shipped size (2^10, prove = one G1 MSM of 1026 points + the CPlink proof) through
    the shim build and through the reference with OpenMP on; None when the drivers were not built."""
    out = {}
    for tag, exe in (("b200", "cplink_b200"), ("reference_omp", "cplink_cpuomp")):
        path = os.path.join(ROOT, "integration", "_ref", exe)
        if not os.path.exists(path):
            return None
        try:
            env = dict(os.environ, B200_GPUS="1")
            env.pop("OMP_NUM_THREADS", None)
            r = subprocess.run([path, "10", "5"], capture_output=True, text=True, timeout=120, env=env)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
            d = json.loads(line)
            out[tag] = {"prove_ms": d["prove_ms_best"], "verified": d["verified"], "fingerprint": d["fingerprint"]}
        except Exception as e:  # the headline numbers do not depend on this leg
            out[tag] = {"error": str(e)[:200]}
    try:
        out["same_proof"] = out["b200"]["fingerprint"] == out["reference_omp"]["fingerprint"]
    except Exception:
        pass
    out["what"] = "integration/cplink_driver.cc over the reference's CPlink classes, log2N = 10, best of 5 proves"
    return out


def stats2_bytes(st, k):
    return float(st[k])


if __name__ == "__main__":
    main()
