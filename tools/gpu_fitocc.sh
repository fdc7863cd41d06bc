for O in 4 5 6; do MOTES_B200_FIT_OCC=$O python bench.py --steps 2 --warmup 1 --frames-per-gpu 128 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('occ $O:', round(d['value']), round(d['ms_per_step'],2), 'bin_fits', d['stage_ms_per_step']['bin_fits'])"; done
