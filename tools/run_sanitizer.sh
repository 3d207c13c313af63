# one compute-sanitizer tool per call (B200_PROFILING.md); small cases only
TOOL=${1:-memcheck}
timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 99 python -m pytest -m gpu -x -q \
  "tests/test_gpu_nn.py::test_nn_ragged_sizes" "tests/test_gpu_nn.py::test_nn_lattice_ties" \
  "tests/test_gpu_keypoints.py::test_keypoint_loop_fewer_than_three_associations" \
  "tests/test_gpu_keypoints.py::test_keypoint_loop_matches_oracle" \
  "tests/test_gpu_cloud.py::test_backproject_ragged_and_empty" "tests/test_gpu_cloud.py::test_backproject_rules_v1" \
  "tests/test_gpu_map.py::test_sync_free_frame_path_equals_the_synchronising_calls" \
  "tests/test_gpu_icp.py" > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer_$TOOL.log
tail -15 gpurun_out/sanitizer_$TOOL.log
