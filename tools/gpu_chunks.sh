python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for C in 1 2 4 8; do MOTES_B200_CHUNKS=$C python bench.py --steps 3 --warmup 2 --frames-per-gpu 128 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $C', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {k:v for k,v in d['stage_ms_per_step'].items() if v>10})"; done
