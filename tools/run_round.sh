# Round-end style check on one B200: GPU tests, smoke, bench, ncu launch list and one full capture of the top kernel.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; cat gpurun_out/bench_r1c.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1c.json 2> gpurun_out/bench_ref_r1c.err; cat gpurun_out/bench_ref_r1c.json
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r1c.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python tools/profile_case.py --iters 0 > gpurun_out/ncu_nn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:nn_partial -c 1 -f -o gpurun_out/prof_nn_warp_q12 python tools/profile_case.py --iters 0 > gpurun_out/ncu_nn_warp_q12.log 2>&1
tail -1 gpurun_out/ncu_nn_plain.log
