python -m pytest tests/test_gpu_icp.py -m gpu -x -q 2>&1 | tail -2
python tools/profile_case.py --iters 20 --grid 0 --repeat 3 | tail -2
