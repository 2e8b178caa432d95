#!/bin/bash
mkdir -p gpurun_out
export DLSG_WORKSPACE_MB=8192 PROF_B=16 PROF_REPS=2
CMD="python tools/prof_ops.py"
timeout 300 $CMD > gpurun_out/plain5.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sense_ -s 10 -c 9 -o gpurun_out/prof_r1c $CMD > gpurun_out/ncu_full5.log 2>&1
tail -2 gpurun_out/plain5.log; tail -3 gpurun_out/ncu_full5.log
