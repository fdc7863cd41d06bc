# usage: bash tools/ncu_round.sh <tag>   (run under gpurun; writes gpurun_out/<tag>_*)
TAG=${1:-r01}
python bench.py --steps 2 --warmup 1 --frames-per-gpu 8 --no-cpu-baseline --profile-steps 0 > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 0 --frames-per-gpu 8 --no-cpu-baseline --profile-steps 0 > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"moffat_fit_batched|sky_column|mt_stream|sky_walk|sky_prepare|mt_jump" -c 8 \
    -o gpurun_out/${TAG}_full -f python bench.py --steps 1 --warmup 0 --frames-per-gpu 8 --no-cpu-baseline --profile-steps 0 > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out | tail -8
