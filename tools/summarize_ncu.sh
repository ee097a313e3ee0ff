#!/bin/bash
# Usage: tools/summarize_ncu.sh gpurun_out/prof.ncu-rep profiles/r01_name
# Writes <out>_metrics.csv (selected raw metrics per launch), <out>_opcodes.txt, <out>_lines.txt
set -e
rep=$1; out=$2
ncu -i "$rep" --page raw --csv > /tmp/_raw.csv 2>/dev/null
python3 - "$out" <<'PY'
import csv, sys
out = sys.argv[1]
rows = list(csv.reader(open('/tmp/_raw.csv')))
hdr, units = rows[0], rows[1]
want = ['ID', 'Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct']
idx = [i for i, h in enumerate(hdr) if h in want]
with open(out + '_metrics.csv', 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
PY
ncu -i "$rep" --page source --csv > /tmp/_src.csv 2>/dev/null
python3 tools/ncu_source_summary.py /tmp/_src.csv 25 > "${out}_opcodes.txt"
ncu -i "$rep" --page source --print-source cuda,sass --csv > /tmp/_src2.csv 2>/dev/null
python3 tools/ncu_line_summary.py /tmp/_src2.csv 40 > "${out}_lines.txt"
echo "wrote ${out}_metrics.csv ${out}_opcodes.txt ${out}_lines.txt"
