# usage (under gpurun): bash tools/ab_env.sh "VAR=val ..." "VAR=val ..." ...  -> device-resident step time under each environment
for E in "$@"; do
  env $E python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-frame --no-parity --profile-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('== $E', round(d['ms_per_step'],2), 'ms/step', round(d['value']), 'col/s')"
done
