#!/bin/bash
# Build experimental variants of libggs_b200.so into genetic-gaussian-splats_b200/lib/variants/
# usage: tools/build_variants.sh name "extra nvcc flags for ggs_raster.cu" [name flags ...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=$ROOT/genetic-gaussian-splats_b200
OUT=$PKG/lib/variants; mkdir -p $OUT $PKG/build/variants
COMMON="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I $ROOT/include -I $PKG/csrc"
python $PKG/build.py > /dev/null
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc $COMMON $flags -Xptxas -v -c $PKG/csrc/ggs_raster.cu -o $PKG/build/variants/raster_$name.o 2>&1 | grep -E "Used" | sed "s/^/$name: /"
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libggs_$name.so $PKG/build/ggs_decode.o $PKG/build/variants/raster_$name.o $PKG/build/ggs_probe.o $PKG/build/ggs_api.o
done
ls $OUT
