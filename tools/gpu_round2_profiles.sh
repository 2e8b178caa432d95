#!/bin/bash
# Round-2 evidence run (one GPU): ncu captures of the shipped kernels, each after the same command has run clean.
mkdir -p gpurun_out
export PROF_B=16 PROF_REPS=2
# 1. sampled-line SENSE kernels: launch list + full capture (DRAM bytes, occupancy, stalls)
timeout 300 python tools/prof_ops.py > gpurun_out/r2k_prof_ops_clean.log 2>&1 || { echo "clean run failed"; tail -5 gpurun_out/r2k_prof_ops_clean.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2k_lines_launches.csv python tools/prof_ops.py > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lines --launch-skip 9 -c 9 -f -o gpurun_out/r2k_lines_full python tools/prof_ops.py > gpurun_out/r2k_ncu_lines.log 2>&1
tail -1 gpurun_out/r2k_ncu_lines.log
# 2. Swin block kernels: tensor-pipe counters
export SWIN_B=8
timeout 300 python tools/bench_swin.py > gpurun_out/r2k_bench_swin.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__inst_executed_pipe_tensor_op_hmma.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"swin_" -s 20 -c 12 --csv --log-file gpurun_out/r2k_swin_tensor_pipe.csv python tools/bench_swin.py > /dev/null 2>&1
tail -3 gpurun_out/r2k_bench_swin.log
