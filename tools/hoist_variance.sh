# usage (under gpurun): bash tools/hoist_variance.sh [runs]  -> the timed step of fresh processes, with the main stream's stage times under the hoisted generator
for i in $(seq 1 ${1:-6}); do
  MOTES_B200_PROFILE_HOIST=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-frame --no-parity --profile-steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']
print('run $i', round(d['ms_per_step'],2), 'ms/step | hoisted stages:', {k: round(v,1) for k,v in s.items() if v>0.5}, 'sum', round(sum(s.values()),1))"
done
