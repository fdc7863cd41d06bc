# usage (under gpurun): bash tools/ab_hoist.sh  -> device-resident step time for generator residencies beside the localisation stages
for C in 8 4 3 2 1; do
  MOTES_B200_HOIST_CTAS=$C python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('hoist ctas=$C', round(d['ms_per_step'],2), 'ms', round(d['value']), 'col/s  e2e', round(d['e2e']['value']), ' single', round(d['single_frame']['ms'],2))"
done
