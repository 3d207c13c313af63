set -x
ICPB_NN_FILTER=2 python -m pytest tests/test_gpu_nn.py tests/test_gpu_icp.py tests/test_gpu_keypoints.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -5
echo "=== fullres"
python tools/profile_case.py --iters 2 --repeat 2 | tail -1
for q in 8 12 16; do ICPB_NN_FILTER=2 ICPB_QPT=$q python tools/profile_case.py --iters 2 --repeat 2 | tail -1; done
ICPB_NN_FILTER=2 python tools/profile_case.py --iters 20 --repeat 1 | tail -1
echo "=== 10k"
python tools/profile_case.py --points 10000 --iters 20 --repeat 3 | tail -1
for q in 2 4 8; do ICPB_NN_FILTER=2 ICPB_QPT=$q python tools/profile_case.py --points 10000 --iters 20 --repeat 3 | tail -1; done
