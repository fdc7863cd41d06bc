TAG=${1:-fit}
ncu --set full --clock-control none --import-source on -k regex:"moffat_fit_group" -s 1 -c 1 \
    -o gpurun_out/${TAG}_fit -f python bench.py --steps 1 --warmup 0 --frames-per-gpu 32 --no-cpu-baseline --profile-steps 0 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
