# A/B of programmatic dependent launch on the solve chain (FLOAM_PDL_SOLVE) and everywhere (FLOAM_PDL)
for cfg in "0 0" "1 0" "0 0" "1 0" "0 1"; do
  set -- $cfg
  FLOAM_PDL_SOLVE=$1 FLOAM_PDL=$2 python bench.py --no-cpu-baseline --steps 400 --warmup 20 --sequences-per-gpu 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['chain']
print('pdl_solve $1 pdl_all $2', 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'half %.1f solve %.1f map %.1f rest %.1f' % (c['pose_dependent_half_us'], c['solve_part_us'], c['map_update_part_us'], c['rest_of_frame_us']))
"
done
