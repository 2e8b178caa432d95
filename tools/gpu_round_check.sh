#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_full.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_full.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_full.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_full.log
timeout 300 python tools/bench_swin.py > gpurun_out/swin_full.log 2>&1
tail -3 gpurun_out/pytest_full.log; tail -2 gpurun_out/smoke_full.log; tail -2 gpurun_out/bench_full.log | cut -c1-1500; tail -4 gpurun_out/swin_full.log
