#!/usr/bin/env python3
"""Print the headline metrics of an .ncu-rep (raw page csv on stdin: ncu -i X --page raw --csv)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, vals = rows[0], rows[-1]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__inst_executed.sum",
        "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for h, v in zip(hdr, vals):
    if h in keys or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        print(f"{h} = {v}")
