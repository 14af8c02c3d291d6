"""Prints the key metrics of an .ncu-rep (read on the CPU box with `ncu -i`)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum ',
        'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_lsu.sum.pct', 'smsp__inst_executed_op_shared', 'local_op',
        'smsp__average_warps_issue_stalled', 'sm__inst_executed_pipe_xu', 'sm__inst_executed_pipe_alu.sum.pct',
        'sm__inst_executed_pipe_fma.sum.pct', 'smsp__inst_executed_pipe_fp64', 'launch__shared_mem_per_block',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'launch__grid_size', 'launch__block_size',
        'smsp__pcsamp_warps_issue_stalled']
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
    for i, h in enumerate(hdr):
        if any(k in h + ' ' for k in KEYS):
            print(f'{h:90s} {units[i]:14s} {r[i]}')
