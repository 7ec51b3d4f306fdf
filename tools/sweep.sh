#!/bin/bash
# BASELINE config 5: render+fitness roofline sweep over splats x image side x population.
# Prints one line per point: candidates/s and the fraction of the nominal fp32 peak.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
printf "%-6s %-7s %-6s %12s %10s %8s %10s\n" side splats pop cand_per_s raster_ms frac pairs/cand
while read side splats pop; do
  timeout 600 python $ROOT/bench.py --side $side --splats $splats --population $pop --pool 2 --steps ${STEPS:-5} --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-6d %-7d %-6d %12.0f %10.3f %8.4f %10.3e' % ($side, $splats, $pop, d['value'], r['raster_ms_per_launch'], r['frac'], r['pairs_per_candidate']))"
done <<'PTS'
128 100 32
128 1000 1024
256 500 8
256 500 1024
256 1000 64
256 1000 1024
256 1000 4096
256 4000 512
512 1000 512
512 4000 256
512 16000 64
1024 1000 256
1024 4000 64
1024 16000 16
PTS
