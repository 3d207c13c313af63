set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests1.log
ICPB_RAY_BRICKS=0 python tools/profile_case.py --iters 0 --map --cm 1 > gpurun_out/r2_map_old.log 2>&1
ICPB_RAY_BRICKS=1 python tools/profile_case.py --iters 0 --map --cm 1 > gpurun_out/r2_map_brick.log 2>&1
timeout 600 python bench.py --workload map1cm --steps 3 --warmup 3 > gpurun_out/r2_map1cm_n1.json 2> gpurun_out/r2_map1cm_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_rays -s 2 -c 1 -o gpurun_out/r2_map_rays_brick -f python tools/profile_case.py --iters 0 --map --cm 1 > gpurun_out/r2_ncu_rays.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -3 gpurun_out/r2_tests1.log; cat gpurun_out/r2_map_old.log gpurun_out/r2_map_brick.log; head -c 600 gpurun_out/r2_map1cm_n1.json
