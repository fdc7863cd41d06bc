# usage (under gpurun): bash tools/ab_libs.sh name1 name2 ...  -> A/B of builds under motes_b200/variants/lib_<name>.so
# ("base" = motes_b200/libmotes_b200_base.so, "tree" = the in-tree library): GPU test subset, then the device-resident
# step and the per-stage CUDA-event times of one profiling step.  TESTS=all runs the whole GPU suite per build.
for N in "$@"; do
  case $N in
    base) LIB=$PWD/motes_b200/libmotes_b200_base.so;;
    tree) LIB=$PWD/motes_b200/libmotes_b200.so;;
    *) LIB=$PWD/motes_b200/variants/lib_$N.so;;
  esac
  export MOTES_B200_LIB=$LIB
  if [ "$TESTS" = all ]; then T="tests"; else T="tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py"; fi
  if [ "$TESTS" != none ]; then echo "== $N: $(python -m pytest $T -m gpu -x -q 2>&1 | tail -1)"; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-frame --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']
print('== $N', round(d['ms_per_step'],2), 'ms/step', {k: s[k] for k in ('boot_stream','boot_walk','boot_column','bin_fits','sky_prepare')})"
done
