#!/usr/bin/env python
"""Summarise an .ncu-rep (details + raw pages) into the few numbers the tuning notes quote."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
keep = ('Duration', 'Registers Per Thread', 'Executed Ipc Active', 'Issue Slots Busy', 'DRAM Throughput', 'Memory Throughput', 'L1/TEX Hit Rate',
        'L2 Hit Rate', 'Achieved Occupancy', 'Avg. Active Threads Per Warp', 'Executed Instructions', 'Theoretical Occupancy', 'No Eligible',
        'Warp Cycles Per Issued Instruction', 'Grid Size', 'Mem Busy', 'Max Bandwidth', 'L1/TEX Cache Throughput', 'L2 Cache Throughput',
        'Local Memory Spilling Requests', 'Branch Efficiency')
out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
cur = None
for r in csv.reader(io.StringIO(out)):
    if len(r) > 14 and r[0] != 'ID':
        if r[0] != cur:
            cur = r[0]
            print(f"--- launch {r[0]}: {r[4][:70]}")
        if r[12] in keep:
            print(f"   {r[12]:38s} {r[14]:>14s} {r[13]}")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'lts__t_bytes.sum', 'lts__t_bytes.sum.per_second', 'lts__t_sectors.sum', 'l1tex__t_bytes.sum', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'gpu__time_duration.sum']
stall = [c for c in h if 'pcsamp_warps_issue_stalled' in c and 'not_issued' not in c]
for k in want + stall:
    if k in h:
        i = h.index(k)
        print(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', 'stall_'):40s} unit {rows[1][i]:8s} " + " ".join(r[i] for r in rows[2:]))
