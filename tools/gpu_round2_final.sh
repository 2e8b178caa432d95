#!/bin/bash
# final validation of the round: GPU suite, smoke, bench (N=1), ncu launch list + one full capture of the convolution kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2y_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2y_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r2y_bench_n1.log 2> gpurun_out/r2y_bench_n1.err
timeout 600 python bench.py --workload swin_recon --steps 4 --warmup 3 > gpurun_out/r2y_bench_swin_recon.log 2> gpurun_out/r2y_bench_swin_recon.err
timeout 600 python bench.py --workload recon_resnet --steps 4 --warmup 3 > gpurun_out/r2y_bench_recon_resnet.log 2> gpurun_out/r2y_bench_recon_resnet.err
timeout 300 python tools/conv_kernel_timing.py > gpurun_out/r2y_conv_kernel_timing.log 2>&1
timeout 300 python tools/denoiser_breakdown.py > gpurun_out/r2y_denoiser_breakdown_swin.log 2>&1
CFG=config_recon NB=8 timeout 300 python tools/denoiser_breakdown.py > gpurun_out/r2y_denoiser_breakdown_resnet.log 2>&1
tools/ncu_conv.sh r2y > gpurun_out/r2y_conv_ncu_summary.txt 2>&1
python tools/ncu_hot_regions.py gpurun_out/conv_r2y.ncu-rep conv3d 20 > gpurun_out/r2y_conv_hot_regions.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2y_launches_swin_recon.csv python bench.py --workload swin_recon --steps 1 --warmup 1 --skip-cpu > gpurun_out/r2y_ncu_swin.log 2>&1
tail -3 gpurun_out/r2y_pytest_gpu.log gpurun_out/r2y_smoke.log
