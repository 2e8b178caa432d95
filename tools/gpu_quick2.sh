#!/bin/bash
mkdir -p gpurun_out
SWEEP_MB=400,640,1280,1700,2600,8192 timeout 600 python tools/sweep_workspace.py 2>&1 | grep workspace > gpurun_out/quick2.log
cat gpurun_out/quick2.log
