#!/bin/bash
# Kernel-variant builds of libshadow_b200.so for exploration runs (scripts/explore_variants.py): selected at run time with
# SHADOW_B200_LIB=<path>.  The variants only differ in compile-time tuning macros of csrc/ppr_warp_kernel.cuh.
set -e
cd "$(dirname "$0")/../shadow_gnn_b200"
mkdir -p variants
build() { name=$1; shift; OUT=../variants/lib_$name.so EXTRA="$*" PTXAS_V=1 bash csrc/build.sh 2>&1 | grep -A2 "ppr_induce_warp" | grep "Used" | sed "s/^/$name: /"; }
build u4c1m28k2r3 -DWARP_U=4 -DWARP_CH=1 -DWARP_MIN_BLOCKS=28 -DWARP_BK=2 -DWARP_OVK=3 &
build u4c1m28k2r2 -DWARP_U=4 -DWARP_CH=1 -DWARP_MIN_BLOCKS=28 -DWARP_BK=2 -DWARP_OVK=2 &
build u4c1m24k2r4 -DWARP_U=4 -DWARP_CH=1 -DWARP_MIN_BLOCKS=24 -DWARP_BK=2 -DWARP_OVK=4 &
wait
ls -la variants
