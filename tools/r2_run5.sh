set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_nn.py tests/test_gpu_icp.py tests/test_gpu_fullshape.py tests/test_gpu_compat.py -m gpu -x -q > gpurun_out/r2_tests5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests5.log
tail -6 gpurun_out/r2_tests5.log
(
for cell in 0 0.05 0.1; do echo "# cell=$cell"; python tools/profile_case.py --grid $cell --iters 20 --repeat 3 --noprof | tail -1; done
) > gpurun_out/r2_coop_sweep5.log 2>&1
grep -v "^+" gpurun_out/r2_coop_sweep5.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_grid_coop -s 3 -c 1 -o gpurun_out/r2_nn_coop5 -f python tools/profile_case.py --grid 0 --iters 6 --noprof > gpurun_out/r2_ncu_coop5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_finalize_coop -s 3 -c 1 -o gpurun_out/r2_nn_fin5 -f python tools/profile_case.py --grid 0 --iters 6 --noprof > gpurun_out/r2_ncu_fin5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_grid5.csv python tools/profile_case.py --grid 0 --iters 20 --noprof > /dev/null 2>&1
python bench.py --workload normals --steps 5 --warmup 3 > gpurun_out/r2_normals_base.json 2>/dev/null
ICPB_LIB=$GRAFT_REPO_ROOT/icp-slam-prototype_b200/variants/libicpb200_nf.so python bench.py --workload normals --steps 5 --warmup 3 > gpurun_out/r2_normals_fast.json 2>/dev/null
ICPB_LIB=$GRAFT_REPO_ROOT/icp-slam-prototype_b200/variants/libicpb200_nf.so timeout 300 python -m pytest tests/test_gpu_cloud.py -m gpu -q -k "normal" > gpurun_out/r2_normals_fast_tests.log 2>&1
tail -2 gpurun_out/r2_normals_fast_tests.log; cut -c1-200 gpurun_out/r2_normals_base.json; cut -c1-200 gpurun_out/r2_normals_fast.json
