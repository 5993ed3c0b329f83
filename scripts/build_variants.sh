#!/bin/bash
# Build kernel-variant libraries of the PPR fast path into shadow_gnn_b200/variants/ (scratch, git-ignored) for scripts/explore_variants.py
set -e
cd "$(dirname "$0")/../shadow_gnn_b200"
rm -rf variants; mkdir -p variants
build() { name=$1; shift; OUT=../variants/lib_$name.so EXTRA="$*" PTXAS_V=1 bash csrc/build.sh 2>&1 | grep -A2 "ppr_induce_warp" | grep "Used" | head -1 | sed "s/^/$name: /"; }
build k64u2m28 -DWARP_K=64 -DWARP_U=2 -DWARP_MIN_BLOCKS=28 &
build k128u2m28 -DWARP_K=128 -DWARP_U=2 -DWARP_MIN_BLOCKS=28 &
build k128u4m24 -DWARP_K=128 -DWARP_U=4 -DWARP_MIN_BLOCKS=24 &
build k256u4m20 -DWARP_K=256 -DWARP_U=4 -DWARP_MIN_BLOCKS=20 &
build k128u4m16 -DWARP_K=128 -DWARP_U=4 -DWARP_MIN_BLOCKS=16 &
wait
ls -la variants
