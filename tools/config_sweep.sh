# usage (under gpurun): bash tools/config_sweep.sh [tag] -> one bench line per BASELINE.json configuration other than the
# headline one (parity cases, not bench lines: device-resident + end to end + parity of frame 0 against the reference)
TAG=${1:-r02}
run() {  # name frames then bench flags
  name=$1; frames=$2; shift 2
  python bench.py --steps 3 --warmup 3 --frames-per-gpu $frames --no-cpu-baseline "$@" 2>gpurun_out/${TAG}_sweep_$name.err > gpurun_out/${TAG}_sweep_$name.json
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_sweep_$name.json')); s=d['stage_ms_per_step']; p=d['parity'] or {}
print('$name F=$frames:', round(d['value']), 'col/s', round(d['ms_per_step'],1), 'ms/step  e2e', round(d['e2e']['value']), ' single', round(d['single_frame']['ms'],2), 'ms  fits', s['bin_fits'], 'stream', s['boot_stream'], 'walk', s['boot_walk'], 'column', s['boot_column'], ' parity ok', p.get('ok'), 'flips', p.get('limit_flips'), 'flux', p.get('flux_max_rel'))"
}
run C1_gmos_3000x200 128 --nspat 200 --ndisp 3000
run C3_xshooter_24000x80_cr 32 --nspat 80 --ndisp 24000 --cosmic-fraction 0.01
run C5a_2048x256_binned 128 --nspat 256 --ndisp 2048
run C5a_2048x256_unbinned 32 --nspat 256 --ndisp 2048 --binning off
run C5b_16384x512_binned 16 --nspat 512 --ndisp 16384
run C2_4096x400_unbinned 16 --binning off
