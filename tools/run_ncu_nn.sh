# ncu capture of the brute-force scan kernel at full resolution: plain run first, then the capture
set -e
Q=${1:-8}
export ICPB_QPT=$Q
python tools/profile_case.py --iters 0 > gpurun_out/ncu_nn_plain_q$Q.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:nn_partial -c 1 -f -o gpurun_out/prof_nn_centred_q$Q python tools/profile_case.py --iters 0 > gpurun_out/ncu_nn_q$Q.log 2>&1
tail -2 gpurun_out/ncu_nn_plain_q$Q.log
