set -x
python -m pytest tests/test_gpu_map.py tests/test_gpu_pipeline.py tests/test_golden.py tests/test_gpu_compat.py -m gpu -x -q 2>&1 | tail -4
python bench.py --workload map1cm 2>&1 | cut -c1-1300
python bench.py --workload trajectory 2>&1 | cut -c1-1300
python tools/profile_case.py --iters 0 --map --cm 1 | tail -3
python tools/profile_case.py --iters 0 --map --cm 2 | tail -3
