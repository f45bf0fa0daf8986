# A/B of the keyframe map update: full re-sort (FLOAM_MAP_MERGE=0) vs classify + sort of the out-of-place points + merge (1)
for m in 0 1; do
  echo "== FLOAM_MAP_MERGE=$m configs[1] hdl64"
  FLOAM_MAP_MERGE=$m python tools/kernel_breakdown.py hdl64 0.4 60 2>&1 | head -16
  echo "== FLOAM_MAP_MERGE=$m configs[3] os1-128 res 0.08"
  FLOAM_MAP_MERGE=$m python tools/kernel_breakdown.py os1-128 0.08 170 90 0.5 2>&1 | head -16
done
for m in 0 1 0 1; do
  FLOAM_MAP_MERGE=$m python bench.py --no-cpu-baseline --steps 400 --warmup 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('merge $m', 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'multi %.0f' % d['multi_sequence']['value'], 'launches/frame', d['launches_per_frame'])
"
done
