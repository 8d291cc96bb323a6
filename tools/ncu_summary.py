"""Condense `ncu -i X.ncu-rep --page raw --csv` into one column per kernel: python tools/ncu_summary.py raw.csv > summary.csv"""
import csv, sys
KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fmaheavy.sum",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Kernel Name" in r)
units = rows[rows.index(hdr) + 1]
data = [r for r in rows[rows.index(hdr) + 2:] if len(r) == len(hdr)]
ki = hdr.index("Kernel Name")
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [r[ki][:48] for r in data])
for m in KEEP:
    if m in hdr:
        j = hdr.index(m)
        w.writerow([m, units[j]] + [r[j] for r in data])
