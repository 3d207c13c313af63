cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests13.log
tail -5 gpurun_out/r2_tests13.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1_v2.json 2> gpurun_out/r2_bench_n1_v2.err; tail -3 gpurun_out/r2_bench_n1_v2.err; cut -c1-300 gpurun_out/r2_bench_n1_v2.json
python -c "import __graft_entry__ as g; g.smoke()"
