#!/usr/bin/env python3
"""tools/ncu_summary.py — turn the raw CSV page of an `ncu --set full` capture (tools/gpu.sh ncufull step) into the
per-kernel summary that is committed under profiles/ and, with --traffic-key, into the entry of
profiles/ncu_traffic.json that bench.py reads for `roofline.traffic` (so the number on the bench line is generated from
a capture, not typed in).

    python tools/ncu_summary.py gpurun_out/<tag>_ncu_<kernels>_raw.csv --out profiles/<tag>_<what>_summary.json \
        [--traffic-key g1_2p20_pre --traffic-kernel k_accumulate]
"""
import argparse, csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = {
    "gpu__time_duration.sum": "time",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fmaheavy_pct",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active": "inst_fmaheavy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__inst_executed_op_local_ld.sum": "local_loads",
    "smsp__inst_executed_op_local_st.sum": "local_stores",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("--out")
    ap.add_argument("--traffic-key")
    ap.add_argument("--traffic-kernel", default="k_accumulate")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw_csv)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        e = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", "")}
        for h, name in KEEP.items():
            if h not in idx:
                continue
            try:
                v = float(r[idx[h]].replace(",", ""))
            except ValueError:
                continue
            u = units[idx[h]]
            if name in ("dram_read", "dram_write"):
                e[name + "_bytes"] = v * UNIT_SCALE.get(u, 1.0)
            elif name == "time":
                e["time_us"] = v * UNIT_SCALE.get(u, 1e-9) * 1e6
            else:
                e[name] = v
        out.append(e)
    doc = {"source": os.path.basename(a.raw_csv), "how": "ncu --set full --clock-control none (tools/gpu.sh ncufull); one entry per captured launch", "launches": out}
    if a.out:
        json.dump(doc, open(a.out, "w"), indent=1)
        print("wrote", a.out)
    else:
        print(json.dumps(doc, indent=1))
    if a.traffic_key:
        sel = [e for e in out if e["kernel"].startswith(a.traffic_kernel)]
        if not sel:
            sys.exit(f"no launch of {a.traffic_kernel} in the capture")
        e = sel[-1]
        p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        t = json.load(open(p)) if os.path.exists(p) else {}
        t[a.traffic_key] = {"kernel": e["kernel"], "dram_read_bytes": e.get("dram_read_bytes"), "dram_write_bytes": e.get("dram_write_bytes"),
                            "time_us_under_ncu": e.get("time_us"), "source": os.path.basename(a.raw_csv), "generated_by": "tools/ncu_summary.py"}
        json.dump(t, open(p, "w"), indent=1)
        print("updated", p, a.traffic_key)


if __name__ == "__main__":
    main()
