#!/bin/bash
# One full ncu capture (with source) of the batched normals kernel: 64 resident frames in one launch.
mkdir -p gpurun_out
python bench.py --workload normals --frames 64 --steps 3 --warmup 2 > gpurun_out/nm_plain.log 2>&1 || { tail -5 gpurun_out/nm_plain.log; exit 1; }
tail -1 gpurun_out/nm_plain.log | cut -c1-160
ncu --set full --clock-control none --import-source on -k regex:normals -s 3 -c 1 -f -o gpurun_out/nm_full \
    python bench.py --workload normals --frames 64 --steps 3 --warmup 2 > gpurun_out/nm_ncu.log 2>&1
tail -1 gpurun_out/nm_ncu.log
