"""Summarise an `ncu -i x.ncu-rep --page raw --csv` dump (one kernel launch) into a Markdown table under profiles/.
    python tools/ncu_raw_summary.py gpurun_out/<tag>/prof_<kernel>.raw.csv profiles/<name>.md ["what was captured"]"""
import csv, sys
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "lts__t_bytes.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
H, U, V = rows[0], rows[1], rows[2]
name = V[H.index("Kernel Name")] if "Kernel Name" in H else "?"
with open(sys.argv[2], "w") as f:
    f.write(f"ncu `--set full --clock-control none` capture, source `{sys.argv[1]}`" + (f" - {sys.argv[3]}" if len(sys.argv) > 3 else "") + "\n\n")
    f.write(f"## `{name[:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
    for k in KEEP:
        if k in H:
            i = H.index(k)
            f.write(f"| {k} | {V[i]} | {U[i]} |\n")
print(open(sys.argv[2]).read()[:300])
