#!/bin/bash
# tools/build_variant.sh NAME [-D...]: tuning build of the library (cine radices only) into build_variants/libdlsg_NAME.so;
# run it with DLSG_LIB_PATH=build_variants/libdlsg_NAME.so.  Not the shipped build (that is __graft_entry__.build()).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
FL="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC"
mkdir -p build_variants
for src in swin_api kt_api; do
  [ build_variants/$src.o -nt dl_swin_gan_b200/csrc/$src.cu ] || nvcc $FL -c -o build_variants/$src.o dl_swin_gan_b200/csrc/$src.cu
done
nvcc $FL -DDLSG_FAST_BUILD "$@" -Xptxas -v -c -o build_variants/sense_$name.o dl_swin_gan_b200/csrc/sense_api.cu 2> build_variants/ptxas_$name.log
nvcc -shared -o build_variants/libdlsg_$name.so build_variants/sense_$name.o build_variants/swin_api.o build_variants/kt_api.o
grep -A2 "lines_kernelILi6\|lines_kernelILi5ELi2" build_variants/ptxas_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo
cuobjdump -sass build_variants/libdlsg_$name.so | python3 -c "
import sys,re
cnt={};name=None
for line in sys.stdin:
    m=re.search(r'Function : (\S+)',line)
    if m: name=m.group(1); cnt[name]=0; continue
    if name and re.match(r'\s+/\*[0-9a-f]{4,6}\*/',line): cnt[name]+=1
for n,c in cnt.items():
    if 'lines_kernel' in n: print(c*16//1024,'KB', n[9:60])
"
