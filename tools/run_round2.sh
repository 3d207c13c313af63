set -x
python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; cat gpurun_out/bench_r1d.json
for w in 10k batch10k trajectory map1cm backproject; do
  python bench.py --workload $w > gpurun_out/bench_r1d_$w.json 2> gpurun_out/bench_r1d_$w.err; cat gpurun_out/bench_r1d_$w.json | cut -c1-1800
done
ICPB_NN_FILTER=1 python bench.py --workload batch10k 2>/dev/null | cut -c1-400
python tools/profile_case.py --iters 0 > gpurun_out/ncu_nn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:nn_partial -c 1 -f -o gpurun_out/prof_nn_centred_q12 python tools/profile_case.py --iters 0 > gpurun_out/ncu_nn_q12.log 2>&1
tail -1 gpurun_out/ncu_nn_plain.log
