"""The host-only half of the C++ drop-in layer: the scalar helpers SURVEY.md 8b keeps on the host (distance,
meanSquareError, calculateOffset, makeRotationMatrix, getVoxelCoordinates, processVoxel, the empty-map answer of
getNearestMappedPoint, Quaternion) in a C++ program written against the reference's names.  No GPU is touched."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "icp-slam-prototype_b200", "lib", "test_compat_host")


def test_host_helpers_program():
    assert os.path.exists(BIN), "lib/test_compat_host missing: run __graft_entry__.build()"
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1] == "ok", r.stdout + r.stderr


def test_drop_in_headers_declare_every_reference_entry_point():
    """the names of icp.hpp:23-43, pointcloud.hpp:27-49 and map.hpp:20-37 all exist in the drop-in headers."""
    inc = os.path.join(ROOT, "include", "icpb200")
    text = "".join(open(os.path.join(inc, f)).read() for f in ("icp.hpp", "pointcloud.hpp", "map.hpp"))
    names = ["getTransformation", "makeRotationMatrix", "meanSquareError", "showAssocations", "calculateOffset", "distance",
             "findGlobalNearestNeighborAssociations", "findGlobalKeyPointAssociations",
             "findMappedNearestNeighborAssociations", "processVoxel", "getNearestMappedPoint", "getNearestPoint",
             "getNearestKeyPoint", "rotate", "translate", "matrix", "centered_matrix", "centered_keypoint_matrix",
             "center_points", "displayColorPoints", "displayKeyPoints", "displayAll", "update", "rayTrace",
             "drawCertaintyMap", "getVoxelCoordinates", "isOccupied", "bound", "mapCloud", "world", "empty", "center",
             "points", "keypoints"]
    missing = [n for n in names if not re.search(r"\b%s\b" % n, text)]
    assert not missing, missing
