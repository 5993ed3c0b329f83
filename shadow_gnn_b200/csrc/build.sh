#!/bin/bash
# Build libshadow_b200.so (C ABI declared in include/shadow_b200.h) for sm_100a, in-tree.
#   OUT=<path>        output file (default ../libshadow_b200.so)
#   EXTRA="-DWARP_U=4 ..."   kernel-variant macros (scripts/build_variants.sh)
# The tcgen05 Linear (gemm_umma.cu) is assembled from the CUTLASS/CuTe templates vendored in the image; its three variants are compiled
# as separate objects in parallel (about a minute each) and cached in csrc/_obj until gemm_umma.cu changes.
set -e
cd "$(dirname "$0")"
OUT=${OUT:-../libshadow_b200.so}
NVCC=${NVCC:-nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2"
# CUTLASS / CuTe headers vendored in the image's site-packages (no import of the package needed)
if [ -z "$CUTLASS_INC" ]; then
  SITE=$(python -c "import sysconfig; print(sysconfig.get_paths()['purelib'])")
  for d in "$SITE/flashinfer/data/cutlass/include" "$SITE/tilelang/3rdparty/cutlass/include"; do
    if [ -f "$d/cutlass/gemm/collective/builders/sm100_9xBF16_umma_builder.inl" ]; then CUTLASS_INC=$d; break; fi
  done
fi
if [ -z "$CUTLASS_INC" ]; then echo "build.sh: no CUTLASS >= 4.x header tree with the sm100 FastFP32 builder found (set CUTLASS_INC)" >&2; exit 1; fi
mkdir -p _obj
pids=""
for v in 0 1 2; do
  if [ ! -f _obj/gemm_umma_$v.o ] || [ gemm_umma.cu -nt _obj/gemm_umma_$v.o ]; then
    $NVCC $ARCH --expt-relaxed-constexpr -diag-suppress 20012 -I"$CUTLASS_INC" -DUMMA_VARIANT=$v -c gemm_umma.cu -o _obj/gemm_umma_$v.o &
    pids="$pids $!"
  fi
done
SRCS="sampler.cu gather.cu ppr_push.cu layers.cu gemm.cu"
SOBJ=_obj/sampler_$(basename "$OUT" .so).o          # variant builds (EXTRA) differ only in this object
$NVCC $ARCH ${PTXAS_V:+-Xptxas -v} $EXTRA -c sampler.cu -o $SOBJ &
pids="$pids $!"
for f in gather ppr_push layers gemm linear_tc gat; do
  $NVCC $ARCH ${PTXAS_V:+-Xptxas -v} -c $f.cu -o _obj/$f.o &
  pids="$pids $!"
done
for p in $pids; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a --shared -o $OUT $SOBJ _obj/gather.o _obj/ppr_push.o _obj/layers.o _obj/gemm.o _obj/linear_tc.o _obj/gat.o _obj/gemm_umma_0.o _obj/gemm_umma_1.o _obj/gemm_umma_2.o
echo "built $(realpath $OUT)"
