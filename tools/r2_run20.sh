cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_fullshape.py -m gpu -x -q > gpurun_out/r2_tests20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests20.log
tail -12 gpurun_out/r2_tests20.log
python tools/profile_case.py --grid 0 --iters 20 --repeat 3 --noprof | tail -1
for m in brute grid; do ICPB_BATCH_NN=$m python bench.py --workload batch10k --steps 3 --warmup 3 > gpurun_out/r2_batch10k_$m.json 2> gpurun_out/r2_batch10k_$m.err; tail -2 gpurun_out/r2_batch10k_$m.err; cut -c1-200 gpurun_out/r2_batch10k_$m.json; done
