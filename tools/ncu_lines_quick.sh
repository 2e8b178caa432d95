#!/bin/bash
# per-kernel average durations (ncu launch list, batch 16) of the sampled-line kernels for each library given as argument
mkdir -p gpurun_out
for so in "$@"; do
  n=$(basename $so .so)
  DLSG_LIB_PATH=$PWD/$so PROF_B=16 PROF_REPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/ql_$n.csv python tools/prof_ops.py > /dev/null 2>&1
  python - "$n" gpurun_out/ql_$n.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[2], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = {}
for r in rows[hdr + 1:]:
    if len(r) > vi and "lines" in r[ki]:
        agg.setdefault(r[ki].split("(")[0][-40:], []).append(float(r[vi].replace(",", "")))
print(sys.argv[1], " | ".join("%s %.1f us" % (k.replace("dlsg::sense_", ""), sum(v) / len(v) / 1e3) for k, v in sorted(agg.items())))
PY
done
