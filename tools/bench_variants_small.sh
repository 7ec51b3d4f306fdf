#!/bin/bash
# Small-batch latency (tools/time_small_batch.py, default entry only) for every library in lib/variants/.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
for lib in $ROOT/genetic-gaussian-splats_b200/lib/variants/*.so; do
  echo "== $(basename $lib)"
  GGS_B200_LIB=$lib timeout 300 python $ROOT/tools/time_small_batch.py 2>&1 | grep -E "default entry|^config|^256|^512" | paste - - | awk -F'[:|]' '{print "   " $1 " -> " $NF}'
done
