#!/usr/bin/env python
"""Turns the .ncu-rep / launch-list files brought back in gpurun_out/ into the small text summaries
committed under profiles/ (the .ncu-rep files themselves are scratch).

    python profiles/summarize.py gpurun_out/prof_r01_gated.ncu-rep > profiles/r01_kernels_gated.txt
    python profiles/summarize.py --launches gpurun_out/launches_r01_gated.csv > profiles/r01_launches.txt
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.max",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
]


def kernels(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    ki = head.index("Kernel Name")
    for r in rows[2:]:
        print(f"== {r[ki]}")
        for m in METRICS:
            if m in head:
                i = head.index(m)
                print(f"   {m:62s} {r[i]:>16s} {units[i]}")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    head = rows[0]
    ki, vi = head.index("Kernel Name"), head.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        if not name.startswith("b200::"):
            continue
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    print(f"{'kernel':48s} {'launches':>8s} {'mean us':>10s} {'share':>7s}   (ncu serialised, cold cache: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:48s} {len(v):8d} {sum(v) / len(v) / 1e3:10.2f} {100 * sum(v) / total:6.1f}%")


def traffic(path, kernel, out_json):
    """dram bytes of one launch of `kernel` from a `--set full` capture -> the JSON bench.py reads for roofline.traffic"""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    ki = head.index("Kernel Name")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        if kernel in r[ki]:
            rd, wr = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
            d = {"kernel": r[ki], "dram_bytes_read": int(float(r[rd].replace(",", "")) * scale[units[rd]]),
                 "dram_bytes_write": int(float(r[wr].replace(",", "")) * scale[units[wr]]),
                 "duration_us": float(r[head.index("gpu__time_duration.sum")].replace(",", "")), "source": path.split("/")[-1],
                 "note": "one launch, `ncu --set full --clock-control none` (replayed, cold cache)"}
            json.dump(d, open(out_json, "w"), indent=1)
            print(d)
            return
    raise SystemExit(f"kernel {kernel} not in {path}")


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        kernels(sys.argv[1])
