set -x
python -m pytest tests -m gpu -x -q tests/test_gpu_cloud.py tests/test_golden.py tests/test_gpu_compat.py tests/test_gpu_pipeline.py 2>&1 | tail -4
python bench.py --workload backproject 2>&1 | cut -c1-900
echo "=== fullres q16 (252 regs, 2 CTAs/SM)"
ICPB_QPT=16 python tools/profile_case.py --iters 2 --repeat 2 | tail -1
ICPB_QPT=16 ICPB_SPLITS=8 python tools/profile_case.py --iters 2 --repeat 2 | tail -1
