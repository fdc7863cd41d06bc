python -m pytest tests -m gpu -x -q 2>&1 | tail -2
MOTES_B200_COLUMN=24 python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
MOTES_B200_COLUMN=200 python -m pytest tests/test_gpu_stages.py -m gpu -x -q 2>&1 | tail -2
bash tools/gpu_breakdown.sh 2>&1 | tail -1 | cut -c1-480
