#!/bin/bash
mkdir -p gpurun_out
SWEEP_MB=8192 timeout 600 python tools/sweep_workspace.py 2>&1 | grep workspace > gpurun_out/quick.log
SWEEP_MB=8192 SWEEP_E=2 SWEEP_B=32 timeout 600 python tools/sweep_workspace.py 2>&1 | grep workspace >> gpurun_out/quick.log
cat gpurun_out/quick.log
