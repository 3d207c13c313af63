cd $GRAFT_REPO_ROOT
echo "# base"; python tools/profile_case.py --grid 0 --iters 20 --repeat 3 | tail -2
for v in c256b8 c384b7 c256b10 g4 g16 g4c256b8; do echo "# $v"; ICPB_LIB=$GRAFT_REPO_ROOT/icp-slam-prototype_b200/variants/lib_$v.so python tools/profile_case.py --grid 0 --iters 20 --repeat 3 | tail -2; done
python -m pytest tests/test_gpu_icp.py -m gpu -q -x -k grid 2>&1 | tail -2
