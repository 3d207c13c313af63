set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_fullshape.py -m gpu -x -q -k "grid or fullres" > gpurun_out/r2_tests6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests6.log
tail -4 gpurun_out/r2_tests6.log
(
for cell in 0.05 0.06 0.075 0.1 0.125; do echo "# cell=$cell"; python tools/profile_case.py --grid $cell --iters 20 --repeat 3 --noprof | tail -1; done
) > gpurun_out/r2_coop_sweep6.log 2>&1
grep -v "^+" gpurun_out/r2_coop_sweep6.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_grid6.csv python tools/profile_case.py --grid 0.1 --iters 20 --noprof > /dev/null 2>&1
