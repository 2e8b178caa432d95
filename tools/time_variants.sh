#!/bin/bash
# time the SENSE ops with each tuning build in build_variants/ (see tools/build_variant.sh)
mkdir -p gpurun_out
for so in build_variants/libdlsg_*.so; do
  n=$(basename $so .so)
  DLSG_LIB_PATH=$PWD/$so timeout 200 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --skip-swin > gpurun_out/variant_$n.log 2>&1
  python - "$n" gpurun_out/variant_$n.log <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    o = j["ops"]
    print(f"{sys.argv[1]:28s} value {j['value']:8.1f}  fwd {o['forward']['ms']:.3f}  adj {o['adjoint']['ms']:.3f}  normal {o['normal']['ms']:.3f}  dc {o['dc_step']['ms']:.3f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
