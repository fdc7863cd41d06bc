# usage (under gpurun): bash tools/ab_hoist.sh  -> device-resident step time with / without the hoisted generator
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -2
for H in 0 1; do
  MOTES_B200_HOIST=$H python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('hoist=$H', round(d['ms_per_step'],2), 'ms', round(d['value']), 'col/s  e2e', round(d['e2e']['value']), ' single', round(d['single_frame']['ms'],2))"
done
