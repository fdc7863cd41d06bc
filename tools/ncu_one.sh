# usage: bash tools/ncu_one.sh <tag> <kernel-regex> [frames] [skip]
TAG=$1; K=$2; F=${3:-32}; S=${4:-0}
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 1 \
    -o gpurun_out/${TAG} -f python bench.py --steps 1 --warmup 0 --frames-per-gpu $F --no-cpu-baseline --profile-steps 0 > gpurun_out/${TAG}.log 2>&1
tail -1 gpurun_out/${TAG}.log | cut -c1-160
