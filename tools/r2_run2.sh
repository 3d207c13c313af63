set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_icp.py tests/test_gpu_fullshape.py tests/test_gpu_nn.py -m gpu -x -q > gpurun_out/r2_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests2.log
tail -15 gpurun_out/r2_tests2.log
python tools/profile_case.py --iters 0 --map --cm 1 > gpurun_out/r2_map_brick2.log 2>&1; cat gpurun_out/r2_map_brick2.log
python tools/profile_case.py --iters 0 --map --cm 2 >> gpurun_out/r2_map_brick2.log 2>&1; tail -3 gpurun_out/r2_map_brick2.log
(
for cm in 0 20 40 75; do echo "# ICPB_GRID_COOP_CM=$cm"; ICPB_GRID_COOP_CM=$cm python tools/profile_case.py --grid 0 --iters 20 --repeat 3 --noprof | tail -1; done
for cell in 0.03 0.04 0.05 0.06 0.1; do echo "# cell=$cell"; python tools/profile_case.py --grid $cell --iters 20 --repeat 3 --noprof | tail -1; done
) > gpurun_out/r2_coop_sweep.log 2>&1
cat gpurun_out/r2_coop_sweep.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_rays -s 2 -c 1 -o gpurun_out/r2_map_rays_brick2 -f python tools/profile_case.py --iters 0 --map --cm 1 > gpurun_out/r2_ncu_rays2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_grid_coop -s 3 -c 1 -o gpurun_out/r2_nn_coop -f python tools/profile_case.py --grid 0 --iters 6 --noprof > gpurun_out/r2_ncu_coop.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_grid.csv python tools/profile_case.py --grid 0 --iters 20 --noprof > /dev/null 2>&1
