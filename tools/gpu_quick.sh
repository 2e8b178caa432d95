#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sense_gpu.py -m gpu -x -q 2>&1 | tail -5
SWEEP_MB=8192 timeout 600 python tools/sweep_workspace.py 2>&1 | grep workspace
