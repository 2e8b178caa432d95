#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
SWEEP_MB=64,8192 timeout 600 python tools/sweep_workspace.py > gpurun_out/sweep4.log 2>&1
SWEEP_MB=8192 SWEEP_E=2 SWEEP_B=32 timeout 600 python tools/sweep_workspace.py > gpurun_out/sweep4_e2.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep4.log gpurun_out/sweep4_e2.log
