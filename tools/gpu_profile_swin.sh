#!/bin/bash
mkdir -p gpurun_out
export SWIN_B=4
CMD="python tools/bench_swin.py"
timeout 300 $CMD > gpurun_out/plain14.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"swin_attn_core|swin_linear_tc" -s 20 -c 8 -o gpurun_out/prof_r1h_swin $CMD > gpurun_out/ncu_full14.log 2>&1
tail -2 gpurun_out/ncu_full14.log
