#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "unrolled_dc or basic_layer" > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu10.log
timeout 1200 python tools/sweep_sizes.py > gpurun_out/sweep_sizes.log 2>&1; echo "exit $?" >> gpurun_out/sweep_sizes.log
tail -3 gpurun_out/pytest_gpu10.log; tail -5 gpurun_out/sweep_sizes.log
