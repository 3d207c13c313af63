# A/B of an environment knob on the three registration shapes: bash tools/ab_env.sh "ICPB_PDL=0" "ICPB_PDL=1"
cd $GRAFT_REPO_ROOT
for setting in "$@"; do
  echo "## $setting"
  env $setting python tools/profile_case.py --points 10000 --iters 20 --repeat 6 --noprof | tail -2
  env $setting python tools/profile_case.py --grid 0 --iters 20 --repeat 4 --noprof | tail -2
  env $setting python tools/profile_case.py --points 10000 --grid 0 --iters 20 --repeat 4 --noprof | tail -1
done
