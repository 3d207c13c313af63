set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== fullres"
python tools/profile_case.py --iters 2 --repeat 2 | tail -1
echo "=== 10k"
for q in 2 4 8; do ICPB_QPT=$q python tools/profile_case.py --points 10000 --iters 20 --repeat 3 | tail -1; done
ICPB_QPT=4 ICPB_SPLITS=15 python tools/profile_case.py --points 10000 --iters 20 --repeat 3 | tail -1
ICPB_QPT=4 ICPB_SPLITS=20 python tools/profile_case.py --points 10000 --iters 20 --repeat 3 | tail -1
ICPB_QPT=8 ICPB_SPLITS=30 python tools/profile_case.py --points 10000 --iters 20 --repeat 3 | tail -1
ICPB_NN_FILTER=1 python tools/profile_case.py --points 10000 --iters 20 --repeat 3 | tail -1
./tools/ubench4
