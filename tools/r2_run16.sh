cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_fullshape.py tests/test_gpu_nn.py tests/test_gpu_map.py -m gpu -x -q > gpurun_out/r2_tests16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests16.log
tail -4 gpurun_out/r2_tests16.log
python tools/profile_case.py --grid 0 --iters 20 --repeat 3 | tail -2
python tools/profile_case.py --grid 0 --iters 20 --repeat 3 --noprof | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_grid16.csv python tools/profile_case.py --grid 0 --iters 20 --noprof > /dev/null 2>&1
