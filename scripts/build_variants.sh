#!/bin/bash
# Build kernel-variant libraries of the PPR fast path into shadow_gnn_b200/variants/ (scratch, git-ignored) for scripts/explore_variants.py
set -e
cd "$(dirname "$0")/../shadow_gnn_b200"
rm -rf variants; mkdir -p variants
build() { name=$1; shift; OUT=../variants/lib_$name.so EXTRA="$*" PTXAS_V=1 bash csrc/build.sh 2>&1 | grep -A2 "ppr_induce_warp" | grep "Used" | head -1 | sed "s/^/$name: /"; }
build s2db -DWARP_U_SYM=2 -DWARP_DB=1 &
build s3db -DWARP_U_SYM=3 -DWARP_DB=1 &
build s3 -DWARP_U_SYM=3 &
build s5 -DWARP_U_SYM=5 &
wait
ls -la variants
