# A/B of environment knobs on the full-resolution cell-grid registration: bash tools/ab_grid.sh "A=1" "A=2 B=3" ...
cd $GRAFT_REPO_ROOT
for setting in "$@"; do
  echo "## $setting"
  env $setting python tools/profile_case.py --grid 0 --iters 20 --repeat 3 | tail -2
done
