# A/B of library variants (make BUILD=build_x OUT=lib/libicpb200_x.so EXTRA=-D... lib/libicpb200_x.so) on the
# full-resolution cell-grid registration: bash tools/ab_lib.sh libicpb200.so libicpb200_x.so ...
cd $GRAFT_REPO_ROOT
L=$GRAFT_REPO_ROOT/icp-slam-prototype_b200/lib
for v in "$@"; do
  echo "## $v"
  ICPB_LIB=$L/$v python tools/profile_case.py --grid 0 --iters 20 --repeat 3 --noprof | tail -2
  ICPB_LIB=$L/$v python tools/profile_case.py --grid 0 --iters 20 --repeat 2 | tail -1
done
