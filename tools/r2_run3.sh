set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_icp.py tests/test_gpu_fullshape.py tests/test_gpu_compat.py tests/test_gpu_keypoints.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/r2_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests3.log
tail -15 gpurun_out/r2_tests3.log
(
for cell in 0 0.1 0.125 0.15; do echo "# cell=$cell"; python tools/profile_case.py --grid $cell --iters 20 --repeat 3 --noprof | tail -1; done
echo "# coop_cm=40"; ICPB_GRID_COOP_CM=40 python tools/profile_case.py --grid 0 --iters 20 --repeat 3 --noprof | tail -1
) > gpurun_out/r2_coop_sweep3.log 2>&1
cat gpurun_out/r2_coop_sweep3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_grid_coop -s 3 -c 1 -o gpurun_out/r2_nn_coop3 -f python tools/profile_case.py --grid 0 --iters 6 --noprof > gpurun_out/r2_ncu_coop3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_finalize -s 3 -c 1 -o gpurun_out/r2_nn_fin3 -f python tools/profile_case.py --grid 0 --iters 6 --noprof > gpurun_out/r2_ncu_fin3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_grid3.csv python tools/profile_case.py --grid 0 --iters 20 --noprof > /dev/null 2>&1
