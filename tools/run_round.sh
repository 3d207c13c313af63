# Round-end style check on one B200 (tools/README.md): GPU tests, smoke, both bench arms, the ncu launch list of the
# bench command and one full capture (with source) of each kernel DESIGN.md quotes a roofline for.
# usage: gpurun --timeout 2400 -- 'bash tools/run_round.sh [tag]'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r02}
O=gpurun_out/$TAG
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; echo "pytest rc=$?" >> ${O}_tests.log; tail -4 ${O}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > ${O}_bench_n1.json 2> ${O}_bench_n1.err; tail -2 ${O}_bench_n1.err; cut -c1-300 ${O}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${O}_bench_ref_n1.json 2> ${O}_bench_ref_n1.err; cut -c1-300 ${O}_bench_ref_n1.json
for wl in 10k live map1cm batch10k; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > ${O}_bench_$wl.json 2> ${O}_bench_$wl.err; cut -c1-300 ${O}_bench_$wl.json
done
# launch list of the bench command (after it exited 0 above without ncu)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file ${O}_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-sharded > ${O}_ncu_bench.log 2>&1
# full captures: cooperative search (pass 5 of a registration), its finalize + solve, the brick ray walk (1 cm map)
python tools/profile_case.py --grid 0 --iters 8 --noprof > ${O}_plain_grid.log 2>&1 && {
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_grid_coop -s 5 -c 1 -f -o ${O}_nn_grid_coop python tools/profile_case.py --grid 0 --iters 8 --noprof > ${O}_ncu_coop.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_finalize_coop -s 5 -c 1 -f -o ${O}_nn_finalize_coop python tools/profile_case.py --grid 0 --iters 8 --noprof > ${O}_ncu_fin.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_solve -s 5 -c 1 -f -o ${O}_nn_solve python tools/profile_case.py --grid 0 --iters 8 --noprof > ${O}_ncu_solve.log 2>&1
}
python tools/profile_case.py --points 10000 --iters 0 --map --cm 1 > ${O}_plain_map.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_rays_brick -s 2 -c 1 -f -o ${O}_map_rays_brick python tools/profile_case.py --points 10000 --iters 0 --map --cm 1 > ${O}_ncu_rays.log 2>&1
python tools/profile_case.py --iters 0 > ${O}_plain_nn.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_partial -c 1 -f -o ${O}_nn_partial_warp python tools/profile_case.py --iters 0 > ${O}_ncu_nn.log 2>&1
ls -la gpurun_out/${TAG}_*
