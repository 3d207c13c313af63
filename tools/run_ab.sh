set -x
python -m pytest tests/test_gpu_cloud.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload normals 2>&1 | tail -1 | cut -c1-1200
python tools/profile_case.py --iters 2 --repeat 2 | tail -1
