set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests4.log
tail -15 gpurun_out/r2_tests4.log
(
for cell in 0 0.05 0.06 0.1; do echo "# cell=$cell"; python tools/profile_case.py --grid $cell --iters 20 --repeat 3 --noprof | tail -1; done
) > gpurun_out/r2_coop_sweep4.log 2>&1
grep -v "^+" gpurun_out/r2_coop_sweep4.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_grid_coop -s 3 -c 1 -o gpurun_out/r2_nn_coop4 -f python tools/profile_case.py --grid 0 --iters 6 --noprof > gpurun_out/r2_ncu_coop4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_grid4.csv python tools/profile_case.py --grid 0 --iters 20 --noprof > /dev/null 2>&1
timeout 600 python bench.py --workload map1cm --steps 3 --warmup 3 > gpurun_out/r2_map1cm_n1_v2.json 2> gpurun_out/r2_map1cm_n1_v2.err; tail -3 gpurun_out/r2_map1cm_n1_v2.err; head -c 400 gpurun_out/r2_map1cm_n1_v2.json
