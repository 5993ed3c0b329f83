#!/bin/bash
# Build kernel-variant libraries of the PPR fast path into shadow_gnn_b200/variants/ (scratch, git-ignored) for scripts/explore_variants.py
set -e
cd "$(dirname "$0")/../shadow_gnn_b200"
rm -rf variants; mkdir -p variants
build() { name=$1; shift; OUT=../variants/lib_$name.so EXTRA="$*" PTXAS_V=1 bash csrc/build.sh 2>&1 | grep -A2 "ppr_induce_warp" | grep "Used" | head -1 | sed "s/^/$name: /"; }
build u4 -DWARP_U=4 &
build u8 -DWARP_U=8 &
build u4m28 -DWARP_U=4 -DWARP_MIN_BLOCKS=28 &
build u3m28 -DWARP_U=3 -DWARP_MIN_BLOCKS=28 &
build u3db -DWARP_U=3 -DWARP_DB=1 &
build u5 -DWARP_U=5 &
wait
ls -la variants
