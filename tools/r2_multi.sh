set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-a}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
python tools/profile_case.py --grid 0 --iters 20 --repeat 3 --noprof | tail -1
timeout 900 $T 29513 bench.py --gpus $N --workload map1cm --steps 3 --warmup 3 > gpurun_out/r2_n${N}_map1cm_$TAG.json 2> gpurun_out/r2_n${N}_map1cm_$TAG.err; tail -3 gpurun_out/r2_n${N}_map1cm_$TAG.err; cut -c1-300 gpurun_out/r2_n${N}_map1cm_$TAG.json
timeout 900 $T 29512 bench.py --gpus $N --workload batch10k --steps 3 --warmup 3 > gpurun_out/r2_n${N}_batch10k_$TAG.json 2> gpurun_out/r2_n${N}_batch10k_$TAG.err; tail -3 gpurun_out/r2_n${N}_batch10k_$TAG.err; cut -c1-300 gpurun_out/r2_n${N}_batch10k_$TAG.json
