TAG=${1:-col}
bash tools/gpu_breakdown.sh 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:"sky_column_hist" -c 1 \
    -o gpurun_out/${TAG}_col -f python bench.py --steps 1 --warmup 0 --frames-per-gpu 32 --no-cpu-baseline --profile-steps 0 > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log | cut -c1-200
