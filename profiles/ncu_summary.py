#!/usr/bin/env python
"""Compact per-launch summary of an .ncu-rep (`ncu --set full`): durations, DRAM bytes, issue/pipe utilisation, stall
reasons. Usage: ncu_summary.py report.ncu-rep out.csv"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]
keep += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"))
idx = [hdr.index(k) for k in keep if k in hdr]
w = csv.writer(open(sys.argv[2], "w", newline=""))
w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
for r in data:
    w.writerow([r[i] for i in idx])
print(f"{len(data)} launches -> {sys.argv[2]}")
