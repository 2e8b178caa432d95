#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/sweep_variants.py > gpurun_out/variants.log 2>&1; echo "exit $?" >> gpurun_out/variants.log
SWEEP_E=2 SWEEP_B=32 SWEEP_MB=8192 timeout 600 python tools/sweep_variants.py > gpurun_out/variants_e2.log 2>&1
cat gpurun_out/variants.log gpurun_out/variants_e2.log
