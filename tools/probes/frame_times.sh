# configs[0], [1], [3] frame times and the bench line (single and multi-sequence) of the current build
python tools/kernel_breakdown.py vlp16 0.4 100 2>&1 | head -1
python tools/kernel_breakdown.py hdl64 0.4 100 2>&1 | head -1
python tools/kernel_breakdown.py os1-128 0.08 170 90 0.5 2>&1 | head -1
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['chain']
print('value %.0f e2e %.0f multi %.0f half %.1f solve %.1f map %.1f' % (d['value'], d['e2e']['value'], d['multi_sequence']['value'], c['pose_dependent_half_us'], c['solve_part_us'], c['map_update_part_us']))"
