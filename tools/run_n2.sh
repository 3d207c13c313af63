T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$T 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; wc -l gpurun_out/bench_n2.json; cut -c1-200 gpurun_out/bench_n2.json
$T 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/bench_n2_ref.json 2>/dev/null; wc -l gpurun_out/bench_n2_ref.json; cut -c1-200 gpurun_out/bench_n2_ref.json
