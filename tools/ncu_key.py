"""Key metrics of an .ncu-rep for the FP64-bound kernels (read with `ncu -i` on the CPU box)."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_inst_executed_op_shared.sum',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'sm__cycles_active.avg',
        'smsp__inst_executed_op_branch.sum', 'sm__inst_executed_pipe_fp64.sum',
        # tensor-core kernels (tcgen05 / TMEM / bulk copies)
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tmem.sum.pct_of_peak_sustained_active', 'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__sass_inst_executed_op_tmem_ldt.sum', 'smsp__sass_inst_executed_op_tmem_stt.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('== kernel:', r[hdr.index('Kernel Name')][:70])
    stalls = {}
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f'{h:72s} {units[i]:16s} {r[i]}')
        if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued'):
            stalls[h.replace('smsp__pcsamp_warps_issue_stalled_', '')] = float(r[i] or 0)
    tot = sum(stalls.values()) or 1
    print('stall samples:', ', '.join(f'{k} {100 * v / tot:.1f}%' for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot > 0.01))
