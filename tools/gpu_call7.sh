#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench7.log 2>&1; echo "bench exit $?" >> gpurun_out/bench7.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench7_ref.log 2>&1
CMD="python bench.py --batch 16 --steps 1 --warmup 3 --skip-cpu --skip-e2e"
timeout 300 $CMD > gpurun_out/plain7.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches7.csv $CMD > gpurun_out/ncu_list7.log 2>&1
tail -2 gpurun_out/bench7.log; tail -1 gpurun_out/bench7_ref.log
