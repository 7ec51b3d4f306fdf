#!/bin/bash
# Run bench.py against every library in lib/variants/ (on the GPU box): one summary line each.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
for lib in $ROOT/genetic-gaussian-splats_b200/lib/variants/*.so; do
  GGS_B200_LIB=$lib timeout 300 python $ROOT/bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-28s value %9.0f  raster_ms %.4f  frac %.4f  e2e %9.0f' % ('$(basename $lib)', d['value'], r['raster_ms_per_launch'], r['frac'], d['e2e']['value']))"
done
