set -x
python bench.py --workload map1cm > gpurun_out/bench_n1_map1cm.json 2> gpurun_out/bench_n1_map1cm.err; cat gpurun_out/bench_n1_map1cm.json | cut -c1-1300; tail -n 3 gpurun_out/bench_n1_map1cm.err
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$T 29513 bench.py --gpus 2 --workload map1cm > gpurun_out/bench_n2_map1cm.json 2> gpurun_out/bench_n2_map1cm.err; cat gpurun_out/bench_n2_map1cm.json | cut -c1-1300
tail -n 4 gpurun_out/bench_n2_map1cm.err
