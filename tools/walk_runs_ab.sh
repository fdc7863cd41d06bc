# usage (under gpurun): bash tools/walk_runs_ab.sh [modes]  -> table walk by runs ("table") against one column at a time ("serial"):
# GPU tests with the table walk forced for every batch size, then one-frame latency and the long-frame configurations
MODES=${1:-"serial table"}
MOTES_B200_WALK=table python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -2
for W in $MODES; do
MOTES_B200_WALK=$W python bench.py --steps 3 --warmup 2 --frames-per-gpu 1 --no-cpu-baseline --no-e2e --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W C2 F=1', round(d['value']), round(d['ms_per_step'],2), 'single', round(d['single_frame']['ms'],2), 'walk', d['stage_ms_per_step']['boot_walk'])"
MOTES_B200_WALK=$W python bench.py --steps 3 --warmup 2 --frames-per-gpu 32 --nspat 80 --ndisp 24000 --cosmic-fraction 0.01 --no-cpu-baseline --no-e2e --no-single-frame 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W C3', round(d['value']), round(d['ms_per_step'],2), 'walk', d['stage_ms_per_step']['boot_walk'], 'parity', (d['parity'] or {}).get('ok'))"
MOTES_B200_WALK=$W python bench.py --steps 3 --warmup 2 --frames-per-gpu 16 --nspat 512 --ndisp 16384 --no-cpu-baseline --no-e2e --no-single-frame 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W C5b', round(d['value']), round(d['ms_per_step'],2), 'walk', d['stage_ms_per_step']['boot_walk'], 'parity', (d['parity'] or {}).get('ok'))"
done
