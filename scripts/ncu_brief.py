#!/usr/bin/env python
"""Key metrics of one ncu report (raw page) in a compact table.  usage: ncu_brief.py X.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg']
for i, h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h and float(vals[i] or 0) > 0.05):
        print(f"{h.replace('smsp__average_warps_issue_stalled_','stall:').replace('_per_issue_active.ratio',''):78s} {vals[i]:>16s} {units[i]}")
