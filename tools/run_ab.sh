for pdl in 0 1; do export ICPB_PDL=$pdl; echo "== PDL $pdl"
python tools/profile_case.py --points 10000 --iters 20 --repeat 4 | tail -1
python tools/profile_case.py --iters 2 --repeat 2 | tail -1
python bench.py --workload batch10k 2>/dev/null | cut -c1-100
python bench.py --workload trajectory 2>/dev/null | cut -c1-100
python bench.py --workload 10k 2>/dev/null | cut -c1-160
done
