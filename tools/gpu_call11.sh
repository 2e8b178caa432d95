#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sense_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu11.log
SWEEP_MB=20,40,80,120,160,240,8192 timeout 600 python tools/sweep_workspace.py > gpurun_out/sweep11.log 2>&1
tail -3 gpurun_out/pytest_gpu11.log; cat gpurun_out/sweep11.log
