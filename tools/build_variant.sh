#!/bin/bash
# Build a compile-time variant of the library + CLI next to the default build, for A/B runs on the GPU box:
#   tools/build_variant.sh t64 -DMV_TILE_V=64 -DMV_ECAP=1024      ->  variants/t64/lib/libmvgpu.so, variants/t64/bin/miniVite_b200
# (variants/ is git-ignored but travels with gpurun).  Macros: MV_TILE_V (vertices per CTA), MV_ECAP (edges staged per
# pass), MV_STAGE_U (loads in flight per thread in phase A).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=variants/$name
mkdir -p $out/lib $out/bin
C=minivite_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-pthread \
     "$@" -shared -o $out/lib/libmvgpu.so $C/mvgpu.cu $C/narrow.cpp -ldl -lpthread
g++ -std=c++17 -O3 -fopenmp -ffp-contract=off -fPIC -I include -o $out/bin/miniVite_b200 $C/host/main.cpp \
    -L $out/lib -lmvgpu '-Wl,-rpath,$ORIGIN/../lib'
echo "built $out"
