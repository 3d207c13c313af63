set -x
python -m pytest tests/test_gpu_cloud.py -m gpu -x -q 2>&1 | tail -3
ICPB_BP_TWOPASS=1 python -m pytest tests/test_gpu_cloud.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload backproject 2>&1 | cut -c1-420
ICPB_BP_TWOPASS=1 python bench.py --workload backproject 2>&1 | cut -c1-420
