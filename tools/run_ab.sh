set -x
python -m pytest tests -m gpu -x -q tests/test_gpu_keypoints.py tests/test_gpu_compat.py 2>&1 | tail -30
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
