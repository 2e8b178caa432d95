#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu8.log
timeout 600 python tools/bench_swin.py > gpurun_out/swin8.log 2>&1
tail -25 gpurun_out/pytest_gpu8.log; cat gpurun_out/swin8.log | tail -4
