set -x
python -m pytest tests -m gpu -x -q tests/test_gpu_icp.py tests/test_gpu_pipeline.py 2>&1 | tail -4
python tools/profile_case.py --iters 20 --grid 0 --repeat 2 | tail -1
python tools/profile_case.py --points 10000 --iters 20 --grid 0 --repeat 2 | tail -1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$T 29513 bench.py --gpus 2 --workload map1cm > gpurun_out/bench_n2_map1cm.json 2> gpurun_out/bench_n2_map1cm.err; cat gpurun_out/bench_n2_map1cm.json | cut -c1-1200
tail -n 3 gpurun_out/bench_n2_map1cm.err
