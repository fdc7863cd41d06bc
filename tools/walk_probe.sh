# usage (under gpurun): bash tools/walk_probe.sh -> segmented == literal check, then walk / single-frame time against the spatial width
python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
for NS in 100 400; do
  python bench.py --steps 2 --warmup 3 --frames-per-gpu 16 --nspat $NS --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('nspat=$NS', 'walk', s['boot_walk'], 'column', s['boot_column'], 'stream', s['boot_stream'], 'fits', s['bin_fits'], 'single', round(d['single_frame']['ms'],2))"
done
