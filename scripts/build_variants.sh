#!/bin/bash
# Kernel-variant builds of libshadow_b200.so for exploration runs (scripts/explore_variants.py): selected at run time with
# SHADOW_B200_LIB=<path>.  The variants only differ in compile-time tuning macros of csrc/ppr_warp_kernel.cuh.
set -e
cd "$(dirname "$0")/../shadow_gnn_b200"
mkdir -p variants
build() { name=$1; shift; OUT=../variants/lib_$name.so EXTRA="$*" PTXAS_V=1 bash csrc/build.sh 2>&1 | grep -A2 "ppr_induce_warp" | grep "Used" | sed "s/^/$name: /"; }
build db1u4m20 -DWARP_DB=1 -DWARP_U=4 -DWARP_MIN_BLOCKS=20 &
build db1u3m22 -DWARP_DB=1 -DWARP_U=3 -DWARP_MIN_BLOCKS=22 &
build db1u2m24 -DWARP_DB=1 -DWARP_U=2 -DWARP_MIN_BLOCKS=24 &
build db0u6m24 -DWARP_DB=0 -DWARP_U=6 -DWARP_MIN_BLOCKS=24 &
build db0u4m22 -DWARP_DB=0 -DWARP_U=4 -DWARP_MIN_BLOCKS=22 &
build db1u6m16 -DWARP_DB=1 -DWARP_U=6 -DWARP_MIN_BLOCKS=16 &
wait
ls -la variants
