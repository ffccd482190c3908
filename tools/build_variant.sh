#!/bin/sh
# Build a variant of librt_b200.so with extra nvcc flags into build/ab/librt_<name>.so (git-ignored, ships with gpurun);
# select it at run time with RT_B200_LIB=build/ab/librt_<name>.so.   usage: tools/build_variant.sh name "-DRT_DEQ_MODE=2"
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$HERE/build/ab"
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
  -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -Xcompiler -Wno-unused-function -cudart static $2 \
  -shared -o "$HERE/build/ab/librt_$1.so" "$HERE"/raytracing2-fork_b200/csrc/*.cu -ldl
echo "built build/ab/librt_$1.so ($2)"
