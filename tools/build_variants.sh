#!/bin/bash
# Build experimental variants of libggs_b200.so into genetic-gaussian-splats_b200/lib/variants/
# usage: tools/build_variants.sh name "extra nvcc flags (all TUs; geometry macros live in the common header)" ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=$ROOT/genetic-gaussian-splats_b200
OUT=$PKG/lib/variants; mkdir -p $OUT $PKG/build/variants
COMMON="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I $ROOT/include -I $PKG/csrc"
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  B=$PKG/build/variants/$name; mkdir -p $B
  nvcc $COMMON $flags -fmad=false -c $PKG/csrc/ggs_decode.cu -o $B/decode.o &
  nvcc $COMMON $flags -c $PKG/csrc/ggs_probe.cu -o $B/probe.o &
  nvcc $COMMON $flags -c $PKG/csrc/ggs_api.cu -o $B/api.o &
  nvcc $COMMON $flags -fmad=false -c $PKG/csrc/ggs_breed.cu -o $B/breed.o &
  nvcc $COMMON -fmad=false -c $PKG/csrc/ggs_mask.cu -o $B/mask.o &
  nvcc $COMMON -c $PKG/csrc/ggs_engine.cu -o $B/engine.o &
  nvcc $COMMON -c $PKG/csrc/ggs_peers.cu -o $B/peers.o &
  nvcc $COMMON $flags ${RASTER_FMAD--fmad=false} -Xptxas -v -c $PKG/csrc/ggs_raster.cu -o $B/raster.o 2>&1 | grep -E "Used|spill" | grep -A1 -B0 "spill" | head -4 | sed "s/^/$name: /"
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libggs_$name.so $B/decode.o $B/raster.o $B/probe.o $B/api.o $B/breed.o $B/mask.o $B/engine.o $B/peers.o
done
ls $OUT
