#!/bin/bash
mkdir -p gpurun_out
export DLSG_WORKSPACE_MB=8192 PROF_B=16 PROF_REPS=2
CMD="python tools/prof_ops.py"
timeout 300 $CMD > gpurun_out/plain12.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sense_rows_adj|sense_cols_kernel<6, 2>|sense_rows_fwd" -s 6 -c 6 -o gpurun_out/prof_r1f $CMD > gpurun_out/ncu_full12.log 2>&1
tail -2 gpurun_out/ncu_full12.log
