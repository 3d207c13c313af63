cd $GRAFT_REPO_ROOT
python tools/profile_case.py --grid 0 --iters 20 --repeat 3
python tools/profile_case.py --grid 0 --iters 20 --repeat 2 --noprof
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_finalize_coop -s 3 -c 1 -o gpurun_out/r2_nn_fin11 -f python tools/profile_case.py --grid 0 --iters 6 --noprof > gpurun_out/r2_ncu_fin11.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_grid_coop -s 5 -c 1 -o gpurun_out/r2_nn_coop11 -f python tools/profile_case.py --grid 0 --iters 8 --noprof > gpurun_out/r2_ncu_coop11.log 2>&1
