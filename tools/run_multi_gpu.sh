set -x
N=${1:-4}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
$T 29511 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json | cut -c1-400
$T 29512 bench.py --gpus $N --workload batch10k > gpurun_out/bench_n${N}_batch10k.json 2> gpurun_out/bench_n${N}_batch10k.err; cat gpurun_out/bench_n${N}_batch10k.json | cut -c1-400
$T 29513 bench.py --gpus $N --workload map1cm > gpurun_out/bench_n${N}_map1cm.json 2> gpurun_out/bench_n${N}_map1cm.err; cat gpurun_out/bench_n${N}_map1cm.json | cut -c1-1500
