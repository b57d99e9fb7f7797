#!/usr/bin/env python
"""Text summary of one .ncu-rep capture (the metrics DESIGN.md cites).  Usage: summarize.py file.ncu-rep"""
import csv, io, subprocess, sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, vals = rows[0], rows[1], rows[2]
keep = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit", "launch__registers_per_thread", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "launch__shared_mem_per_block_dynamic")
print("kernel", vals[head.index("Kernel Name")] if "Kernel Name" in head else "?")
for k, u, v in sorted(zip(head, units, vals)):
    if k.startswith(keep) and "not_issued" not in k and "per_warp_active" not in k and "peak_sustained_active.pct" not in k:
        print("  %s %s %s" % (k, u, v))
