#!/bin/bash
# Sweep nn_partial build variants and work decompositions (full resolution unless POINTS is set).
P=${POINTS:-0}
for lib in ${LIBS:-default u1 u4 u8 sg32 sg8}; do
  if [ $lib = default ]; then unset ICPB_LIB; else export ICPB_LIB=$PWD/icp-slam-prototype_b200/variants/lib_$lib.so; fi
  CFGS=${CFGS:-8:3 8:5 8:7 4:3 4:5 4:2}
  for cfg in $CFGS; do
    q=${cfg%%:*}; s=${cfg#*:}
    echo -n "lib=$lib qpt=$q splits=$s  "
    ICPB_QPT=$q ICPB_SPLITS=$s python tools/profile_case.py --points $P --iters 2 | sed 's/.*gpu_ms/gpu_ms/'
  done
done
