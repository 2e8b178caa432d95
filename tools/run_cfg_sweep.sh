#!/bin/bash
# sweep the buffer configurations of the pruned column kernels: tools/run_cfg_sweep.sh build_variants/libdlsg_X.so
so=${1:-build_variants/libdlsg_p6.so}
export DLSG_LIB_PATH=$PWD/$so
python tools/quick_ops.py 2>&1 | tail -1
for xb in ${XBS:-1 2}; do for axb in ${AXBS:-1 2}; do
  echo "cfg fwd xb=$xb adj xb=$axb"; DLSG_PCOLS_XB=$xb DLSG_PCOLS_ADJ_XB=$axb tools/ncu_lines_quick.sh $so
done; done
