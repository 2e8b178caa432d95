#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_swin_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_swin.py 2>&1 | tail -4
