cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests21.log
tail -6 gpurun_out/r2_tests21.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1_v3.json 2> gpurun_out/r2_bench_n1_v3.err; tail -3 gpurun_out/r2_bench_n1_v3.err; cut -c1-200 gpurun_out/r2_bench_n1_v3.json
python bench.py --workload 10k --steps 5 --warmup 3 > gpurun_out/r2_bench_10k.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_10k.json
python bench.py --workload live > gpurun_out/r2_bench_live.json 2> gpurun_out/r2_bench_live.err; tail -2 gpurun_out/r2_bench_live.err; cat gpurun_out/r2_bench_live.json | cut -c1-1500
