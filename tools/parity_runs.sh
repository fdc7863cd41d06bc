# usage (under gpurun): bash tools/parity_runs.sh -> gpurun_out/parity_<case>.json (shipped build) and
# gpurun_out/parity_realpow_<case>.json (build with -DMOTES_FIT_REAL_POW: motes_b200/variants/realpow.so, made by
# `make -C motes_b200/csrc OUT=../variants/realpow.so EXTRA=-DMOTES_FIT_REAL_POW` in the build container)
python -m pytest tests/test_gpu_configs.py -q -m gpu 2>&1 | tail -2
if [ -f motes_b200/variants/realpow.so ]; then
  MOTES_B200_LIB=$PWD/motes_b200/variants/realpow.so MOTES_PARITY_TAG=realpow_ python -m pytest tests/test_gpu_configs.py -q -m gpu -k full_size 2>&1 | tail -2
fi
