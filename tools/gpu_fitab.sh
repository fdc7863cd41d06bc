for cfg in "4 0 0" "4 1 0" "5 0 28" "5 0 16" "6 0 16" "6 0 8"; do set -- $cfg
MOTES_B200_FIT_OCC=$1 MOTES_B200_FIT_TEAM=$2 MOTES_B200_FIT_GROUP=$3 python bench.py --steps 2 --warmup 1 --frames-per-gpu 128 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('occ $1 team $2 group $3:', round(d['value']), round(d['ms_per_step'],2), 'bin_fits', d['stage_ms_per_step']['bin_fits'])"; done
