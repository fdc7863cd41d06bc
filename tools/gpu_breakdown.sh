python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -4
MOTES_B200_STREAM_GB=0.05 python -m pytest tests/test_gpu_stages.py -m gpu -x -q 2>&1 | tail -2
for F in 1 8 128; do python bench.py --steps 2 --warmup 1 --frames-per-gpu $F --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['frames_per_gpu'], round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['stage_ms_per_step'])"; done
