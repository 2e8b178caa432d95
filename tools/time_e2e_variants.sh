#!/bin/bash
# end-to-end bench value with each tuning build in build_variants/
for so in build_variants/libdlsg_*.so; do
  n=$(basename $so .so)
  for r in 1 2; do
    echo -n "$n: "; DLSG_LIB_PATH=$PWD/$so timeout 200 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-swin 2>&1 | grep -o "\"e2e\": {\"value\": [0-9.]*"
  done
done
