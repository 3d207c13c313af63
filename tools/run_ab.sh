set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ICPB_NN_FILTER=2 python -m pytest tests/test_gpu_icp.py tests/test_gpu_nn.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload batch10k 2>/dev/null | cut -c1-500
ICPB_NN_FILTER=3 python bench.py --workload batch10k 2>/dev/null | cut -c1-200
for cfg in "8 2" "8 4" "12 3" "4 2"; do set -- $cfg; ICPB_QPT=$1 ICPB_SPLITS=$2 python bench.py --workload batch10k 2>/dev/null | sed "s/.*\"value\": \([0-9.]*\).*\"ms_per_step\": \([0-9.]*\).*/q$1 s$2 reg\/s \1 ms \2/"; done
