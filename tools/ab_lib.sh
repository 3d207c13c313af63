cd $GRAFT_REPO_ROOT
L=$GRAFT_REPO_ROOT/icp-slam-prototype_b200/lib
for v in libicpb200.so libicpb200_w2.so libicpb200_w1.so; do
  echo "## $v"
  ICPB_LIB=$L/$v python tools/profile_case.py --grid 0 --iters 20 --repeat 3 --noprof | tail -2
  ICPB_LIB=$L/$v python tools/profile_case.py --grid 0 --iters 20 --repeat 2 | tail -1
done
