#!/bin/bash
# Build libshadow_b200.so (C ABI declared in include/shadow_b200.h) for sm_100a, in-tree.
#   OUT=<path>        output file (default ../libshadow_b200.so)
#   EXTRA="-DWARP_U=4 ..."   kernel-variant macros (scripts/build_variants.sh)
set -e
cd "$(dirname "$0")"
OUT=${OUT:-../libshadow_b200.so}
NVCC=${NVCC:-nvcc}
SRCS="sampler.cu gather.cu $(ls ppr_push.cu layers.cu 2>/dev/null || true)"
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 \
  ${PTXAS_V:+-Xptxas -v} $EXTRA --shared -o $OUT $SRCS
echo "built $(realpath $OUT)"
