# usage (under gpurun): bash tools/ab_env_stage.sh "VAR=val" ...  -> step time and stage times under each environment
for E in "$@"; do
  env $E python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-frame --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']
print('== $E', round(d['ms_per_step'],2), 'ms/step', {k: s[k] for k in ('median_fit','bin_fits','boot_stream','boot_column')})"
done
