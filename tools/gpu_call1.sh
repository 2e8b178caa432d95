#!/bin/bash
# first GPU call of the round: parity, bench, workspace sweep, ncu launch list + full capture of the top kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
timeout 600 python tools/sweep_workspace.py > gpurun_out/sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep.log
CMD="python bench.py --batch 8 --steps 1 --warmup 3 --skip-cpu --skip-e2e"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sense_ -s 60 -c 6 -o gpurun_out/prof_r1a $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log; tail -4 gpurun_out/sweep.log
