# usage (under gpurun): bash tools/round_artifacts.sh r02   -> gpurun_out/<tag>_*
# 1. pytest -m gpu   2. both bench arms   3. ncu launch list of ONE step (this repo's kernels only: torch's frame
# generator is filtered out by name) with time, DRAM bytes and FP64 op counts per launch   4. ncu --set full of the
# dominant kernels of that step
TAG=${1:-r02}
MINE='regex:moffat|fit_|sky_|mt_|seg_plan|row_median|frame_p|column_stats|qual_|bins_kernel|bin_p|limits_|optimal_|widen_|poly_|fp64_peak|walk_'
ONE_STEP="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-e2e --no-single-frame --profile-steps 0"
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
# (generator un-hoisted under ncu so that the launch order is the same on every run)
MOTES_B200_HOIST=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
    --clock-control none -k "$MINE" --csv --log-file gpurun_out/${TAG}_launches.csv $ONE_STEP > gpurun_out/${TAG}_launches.log 2>&1
MOTES_B200_HOIST=0 ncu --set full --clock-control none --import-source on -k regex:"fit_sweep_kernel|fit_step_kernel" -s 8 -c 2 \
    -o gpurun_out/${TAG}_full_fit -f $ONE_STEP > gpurun_out/${TAG}_full_fit.log 2>&1
MOTES_B200_HOIST=0 ncu --set full --clock-control none --import-source on -k regex:"sky_column_hist|mt_stream|sky_walk_kernel|sky_prepare" -c 4 \
    -o gpurun_out/${TAG}_full_boot -f $ONE_STEP > gpurun_out/${TAG}_full_boot.log 2>&1
ls -la gpurun_out | grep ${TAG}_
