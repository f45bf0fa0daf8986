for m in 4 1 2 4 1 2; do
  FLOAM_GRID_MULT=$m python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['chain']
print('grid mult $m value %.0f e2e %.0f multi %.0f half %.1f solve %.1f map %.1f' % (d['value'], d['e2e']['value'], d['multi_sequence']['value'], c['pose_dependent_half_us'], c['solve_part_us'], c['map_update_part_us']))"
done
