cd $GRAFT_REPO_ROOT
for s in "$@"; do
  echo "## $s"; env $s python tools/profile_case.py --points 10000 --iters 20 --repeat 4 --noprof | tail -1
done
