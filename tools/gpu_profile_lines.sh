#!/bin/bash
# launch list + one full ncu capture of the sampled-line SENSE kernels (after the same command ran clean)
mkdir -p gpurun_out
export PROF_B=${PROF_B:-16} PROF_REPS=2
timeout 300 python tools/prof_ops.py > gpurun_out/prof_ops_clean.log 2>&1 || { echo "clean run failed"; tail -5 gpurun_out/prof_ops_clean.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/lines_launches.csv python tools/prof_ops.py > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lines --launch-skip 9 -c 9 -f -o gpurun_out/lines_full python tools/prof_ops.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/lines_launches.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    agg.setdefault(r[ki][:70], []).append(float(r[vi].replace(",", "")))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v)/1e3:9.1f} us total  n={len(v):3d}  avg={sum(v)/len(v)/1e3:8.1f} us  {k}")
PY
