#!/usr/bin/env python
"""Print the handful of ncu raw metrics the profile summaries quote: ncu -i X.ncu-rep --page raw --csv | python ncu_keys.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
H, U, V = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'lts__t_bytes.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_ld.sum',
        'smsp__inst_executed_op_global_st.sum', 'launch__occupancy_limit_registers', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for i, h in enumerate(H):
    if h in keys or ('issue_stalled' in h and 'per_issue_active' in h):
        print(f"{h} = {V[i]} {U[i]}")
