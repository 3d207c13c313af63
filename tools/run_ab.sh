for v in default u1 u4 sg8 st3; do
  if [ $v = default ]; then unset ICPB_LIB; else export ICPB_LIB=$PWD/icp-slam-prototype_b200/variants/lib_$v.so; fi
  echo -n "$v  "; python tools/profile_case.py --iters 2 --repeat 2 | tail -1 | sed 's/.*gpu_ms/gpu_ms/'
done
