#!/bin/bash
# ncu --set full of the tensor-core convolution kernel; writes the key metrics to gpurun_out/conv_ncu_$1.csv
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d -c 1 -s 1 -o gpurun_out/conv_$tag -f python tools/conv_one.py > gpurun_out/conv_ncu_$tag.log 2>&1
ncu -i gpurun_out/conv_$tag.ncu-rep --page raw --csv > gpurun_out/conv_ncu_raw_$tag.csv 2>/dev/null
python - gpurun_out/conv_ncu_raw_$tag.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
h = rows[0]; v = rows[2] if len(rows) > 2 else rows[1]
want = ["gpu__time_duration.sum", "sm__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__pipe_tensor_subpipe_", "sm__inst_executed_pipe_tensor",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared", "lts__t_bytes.sum", "lts__throughput", "l1tex__throughput",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput", "smsp__average_warp", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "sm__pipe_tensor", "tensor", "smsp__warp_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared"]
for k, val in zip(h, v):
    if any(w in k for w in want) and val not in ("0", "", "n/a"):
        print(k, "=", val)
PY
