TAG=${1:-boot}
ncu --set full --clock-control none --import-source on -k regex:"mt_stream|sky_column|sky_walk" -c 3 \
    -o gpurun_out/${TAG}_boot -f python bench.py --steps 1 --warmup 0 --frames-per-gpu 32 --no-cpu-baseline --profile-steps 0 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-300
