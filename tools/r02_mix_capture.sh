#!/bin/bash
# Instruction-mix capture of the pixel kernel: per-class thread instructions and pipe use.
# Usage (through gpurun): bash tools/r02_mix_capture.sh <tag>
TAG=${1:-r02_base}
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
for op in fp32 fp64 integer control memory conversion bit misc inter_thread_communication uniform; do M=$M,smsp__sass_thread_inst_executed_op_${op}_pred_on.sum; done
for p in alu fma fmaheavy fmalite fp64 lsu xu cbu adu uniform tex; do M=$M,smsp__inst_executed_pipe_${p}.sum; done
M=$M,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none -k regex:render_kernel --launch-skip 2 -c 1 --csv --log-file gpurun_out/${TAG}_mix.csv python tools/prof_target.py ${2:-bunny_4k} > gpurun_out/${TAG}_mix.log 2>&1
tail -3 gpurun_out/${TAG}_mix.log
