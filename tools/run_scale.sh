# bench.py (both arms) on N GPUs of one box, as the driver launches it: gpurun --gpus N -- 'bash tools/run_scale.sh N [tag]'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-r02}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 1200 $T 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; tail -3 gpurun_out/${TAG}_bench_n$N.err; cut -c1-200 gpurun_out/${TAG}_bench_n$N.json
timeout 600 $T 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_n$N.json 2> gpurun_out/${TAG}_bench_ref_n$N.err; cut -c1-200 gpurun_out/${TAG}_bench_ref_n$N.json
