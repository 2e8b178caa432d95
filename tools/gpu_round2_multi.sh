#!/bin/bash
# two-GPU checks (gpurun --gpus 2): data-consistency bench (weak + strong scaling, e2e), training slice under DDP
mkdir -p gpurun_out
P=29531
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P++)) bench.py --gpus 2 "$@"; }
run --steps 5 --warmup 3 --skip-cpu --skip-full --skip-swin > gpurun_out/r2l_bench_n2.log 2>&1; tail -c 1800 gpurun_out/r2l_bench_n2.log; echo
run --steps 5 --warmup 3 --skip-cpu --skip-full --skip-swin --skip-e2e --scaling strong > gpurun_out/r2l_bench_n2_strong.log 2>&1; tail -1 gpurun_out/r2l_bench_n2_strong.log | cut -c1-400; echo
run --workload train_swin --steps 4 --warmup 3 > gpurun_out/r2l_train_n2.log 2>&1; tail -1 gpurun_out/r2l_train_n2.log | cut -c1-1200; echo
python -m pytest tests/test_sharding_gpu.py -m gpu -q 2>&1 | tail -2
