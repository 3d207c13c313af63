set -x
python -m pytest tests/test_gpu_cloud.py tests/test_golden.py tests/test_gpu_compat.py tests/test_gpu_map.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload backproject 2>&1 | tail -1 | cut -c1-700
python bench.py --workload map1cm 2>&1 | tail -1 | cut -c1-200
