#!/bin/bash
# BASELINE config 5: render+fitness roofline sweep over splats x image side x population.
# One line per point: candidates/s, raster time, `frac` (the reference's 23 flops for every in-AABB
# pair / time / measured FFMA peak -- exceeds 1 where the saturation stop skips hidden splats),
# the pairs really evaluated, `frac_evaluated` (flops really executed on them / time / peak), and,
# with NCU=1, the FMA-pipe and issue utilisation of the raster launch from ncu.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
printf "%-6s %-7s %-6s %12s %10s %8s %10s %10s %9s %9s %9s\n" side splats pop cand_per_s raster_ms frac pairs/cand eval/cand frac_eval fma_pipe% issue%
while read side splats pop; do
  args="--side $side --splats $splats --population $pop --pool 2 --warmup 3 --no-cpu --no-reference-gpu --no-config4"
  line=$(timeout 600 python $ROOT/bench.py $args --steps ${STEPS:-5} 2>/dev/null)
  pipe="-"; issue="-"
  if [ "${NCU:-0}" = "1" ]; then
    m=$(timeout 600 ncu --metrics sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
        --clock-control none -k regex:raster --launch-skip 4 --launch-count 1 --csv python $ROOT/bench.py $args --steps 2 2>/dev/null | grep -E "pipe_fma_cycles_active|issue_active")
    pipe=$(echo "$m" | grep pipe_fma | awk -F'","' '{print $NF}' | tr -d '"')
    issue=$(echo "$m" | grep issue_active | awk -F'","' '{print $NF}' | tr -d '"')
  fi
  echo "$line" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-6d %-7d %-6d %12.0f %10.3f %8.4f %10.3e %10.3e %9.4f %9s %9s' % ($side, $splats, $pop, d['value'], r['raster_ms_per_launch'], r['frac'], r['pairs_per_candidate'], r['evaluated_pairs_per_candidate'], r['frac_evaluated'], '$pipe', '$issue'))"
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
