#!/bin/bash
# One full ncu capture (with source) of the batched back-projection kernel: 64 resident frames in one launch.
mkdir -p gpurun_out
python bench.py --workload backproject --frames 64 --steps 3 --warmup 2 > gpurun_out/bp_plain.log 2>&1 || { tail -5 gpurun_out/bp_plain.log; exit 1; }
tail -1 gpurun_out/bp_plain.log
ncu --set full --clock-control none --import-source on -k regex:backproject -s 3 -c 1 -f -o gpurun_out/bp_full \
    python bench.py --workload backproject --frames 64 --steps 3 --warmup 2 > gpurun_out/bp_ncu.log 2>&1
tail -2 gpurun_out/bp_ncu.log; ls -la gpurun_out/bp_full.ncu-rep
