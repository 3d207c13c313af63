set -x
python -m pytest tests -m gpu -x -q tests/test_gpu_nn.py tests/test_gpu_icp.py 2>&1 | tail -4
echo "=== fullres"
for q in 8 12 16 24; do
  ICPB_NN_FILTER=1 ICPB_QPT=$q python tools/profile_case.py --iters 2 --repeat 2 | tail -1
done
echo "=== 10k"
for q in 2 4 8; do ICPB_NN_FILTER=1 ICPB_QPT=$q python tools/profile_case.py --points 10000 --iters 20 --repeat 3 | tail -1; done
