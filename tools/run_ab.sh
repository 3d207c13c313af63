set -x
python -m pytest tests/test_gpu_cloud.py tests/test_golden.py tests/test_gpu_compat.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -12
python bench.py --workload backproject 2>&1 | cut -c1-900
python bench.py --workload trajectory 2>&1 | cut -c1-1200
