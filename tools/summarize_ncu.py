#!/usr/bin/env python3
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).

    python tools/summarize_ncu.py launches gpurun_out/<tag>/launches.csv profiles/<name>.md
    python tools/summarize_ncu.py full gpurun_out/<tag>/<rep>.ncu-rep profiles/<name>.md
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum.per_cycle_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "lts__t_bytes.sum", "gpc__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "sm__icc_request_hit_rate", "local_load", "local_store", "l1tex__t_bytes_pipe_lsu_mem_local", "smsp__inst_executed_op_local"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        k = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), source `{src}`\n\n")
        f.write("Per-launch times are cold-cache and serialised: read the SHARES, not the absolutes.\n\n| total ms | launches | ms/launch | share | kernel |\n|---:|---:|---:|---:|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {t / 1e3:.3f} | {c} | {t / 1e3 / c:.3f} | {100 * t / tot:.1f}% | `{k[:100]}` |\n")


def full(src, dst):
    if src.endswith(".csv"):     # the raw page exported on the GPU box (ncu -i rep --page raw --csv)
        out = open(src).read()
    else:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"ncu `--set full --clock-control none` capture, source `{src}`\n\n")
        for V in rows[2:]:
            name = V[H.index("Kernel Name")] if "Kernel Name" in H else "?"
            f.write(f"## `{name[:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for h_, u, v in zip(H, U, V):
                if any(k in h_ for k in KEEP):
                    f.write(f"| {h_} | {v} | {u} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
