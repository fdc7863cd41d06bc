# usage (under gpurun): bash tools/config_sweep.sh -> one bench line per BASELINE.json frame shape (device-resident + end to end)
run() {  # name nspat ndisp frames
  python bench.py --steps 3 --warmup 3 --nspat $2 --ndisp $3 --frames-per-gpu $4 --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_$1.json
  python -c "
import json; d=json.load(open('gpurun_out/sweep_$1.json')); s=d['stage_ms_per_step']
print('$1 ${2}x$3 F=$4:', round(d['value']), 'col/s', round(d['ms_per_step'],1), 'ms/step  e2e', round(d['e2e']['value']), ' single', round(d['single_frame']['ms'],2), 'ms  fits', s['bin_fits'], 'stream', s['boot_stream'], 'walk', s['boot_walk'], 'column', s['boot_column'])"
}
run C1 200 3000 128
run C3 80 24000 32
run C5a 256 2048 128
run C5b 512 16384 16
