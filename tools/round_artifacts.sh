# usage (under gpurun): bash tools/round_artifacts.sh r01   -> gpurun_out/<tag>_*
TAG=${1:-r01}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --profile-steps 0 > gpurun_out/${TAG}_launches.log 2>&1
# (generator un-hoisted so that the launch order, and with it -s / -c, is the same on every run)
MOTES_B200_HOIST=0 ncu --set full --clock-control none --import-source on -k regex:"moffat_fit_group|sky_column_hist|mt_stream|sky_walk|sky_prepare" -s 1 -c 6 \
    -o gpurun_out/${TAG}_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --profile-steps 0 > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out | grep ${TAG}_
