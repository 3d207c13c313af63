for v in default g8b16 g8b32 g4b32 g8b32r4 g8b32r16 g16b32; do
  if [ $v = default ]; then unset ICPB_LIB; else export ICPB_LIB=$PWD/icp-slam-prototype_b200/variants/lib_$v.so; fi
  echo "== $v"; python tools/profile_case.py --iters 0 --map --cm 1 | tail -2 | sed 's/.*visited//'
  python bench.py --workload map1cm 2>/dev/null | sed 's/.*"ms_per_step": \([0-9.]*\).*/map1cm ms_per_step \1/'
done
