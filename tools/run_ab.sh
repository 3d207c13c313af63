python -m pytest tests/test_gpu_icp.py tests/test_gpu_nn.py -m gpu -x -q 2>&1 | tail -4
