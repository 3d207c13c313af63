cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_fullshape.py -m gpu -x -q -k "grid or fullres" > gpurun_out/r2_tests17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests17.log
tail -4 gpurun_out/r2_tests17.log
echo "# sub8 g4"; python tools/profile_case.py --grid 0 --iters 20 --repeat 3 | tail -2
echo "# sub1 g4"; ICPB_GRID_SUB=1 python tools/profile_case.py --grid 0 --iters 20 --repeat 3 | tail -2
echo "# sub8 g8"; ICPB_LIB=$GRAFT_REPO_ROOT/icp-slam-prototype_b200/variants/lib_g8.so python tools/profile_case.py --grid 0 --iters 20 --repeat 3 | tail -2
for cell in 0.08 0.125 0.15 0.2; do echo "# sub8 g4 cell=$cell"; python tools/profile_case.py --grid $cell --iters 20 --repeat 3 | tail -2; done
for cell in 0.125 0.15; do echo "# sub8 g8 cell=$cell"; ICPB_LIB=$GRAFT_REPO_ROOT/icp-slam-prototype_b200/variants/lib_g8.so python tools/profile_case.py --grid $cell --iters 20 --repeat 3 | tail -2; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_grid_coop -s 5 -c 1 -o gpurun_out/r2_nn_coop17 -f python tools/profile_case.py --grid 0 --iters 8 --noprof > gpurun_out/r2_ncu_coop17.log 2>&1
