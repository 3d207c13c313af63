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


def test_host_helpers_equal_the_reference_build():
    """The same helpers against the reference's OWN functions compiled by path (Route B, oracle/_ref): makeRotationMatrix
    (icp.cpp:640-653), getVoxelCoordinates (map.cpp:55-85), distance (icp.cpp:606-620), meanSquareError (:622-638), calculateOffset (:314-344) --
    bit for bit on a seeded sweep."""
    import numpy as np
    import pytest
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (no reference checkout)")
    r = subprocess.run([BIN, "dump"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0
    seen = {"R": 0, "P": 0, "E": 0, "O": 0}
    for line in r.stdout.splitlines():
        kind, *vals = line.split()
        seen[kind] += 1
        if kind == "R":
            f = np.array(vals, dtype=np.float32)
            want = ref.make_rotation(float(f[0]), float(f[1]), float(f[2]))
            assert np.array_equal(f[3:].reshape(3, 3).view(np.uint32), want.view(np.uint32)), line
        elif kind == "P":
            a = np.array(vals[0:3], dtype=np.float32); b = np.array(vals[3:6], dtype=np.float32)
            assert tuple(int(x) for x in vals[6:9]) == ref.voxel(a), line
            pa = np.zeros(1, dtype=_point_dtype()); pb = np.zeros(1, dtype=_point_dtype())
            for k, n in enumerate("xyz"):
                pa[n] = a[k]; pb[n] = b[k]
            assert np.float32(vals[9]) == np.float32(ref.distance(pa, pb)), line
        elif kind == "O":
            n = int(vals[0]); f = np.array(vals[1:], dtype=np.float32)
            pairs = f[:6 * n].reshape(n, 6)
            pa = np.zeros(n, dtype=_point_dtype()); pb = np.zeros(n, dtype=_point_dtype())
            for k, name in enumerate("xyz"):
                pa[name] = pairs[:, k]; pb[name] = pairs[:, 3 + k]
            assert np.array_equal(f[6 * n:].view(np.uint32), ref.calculate_offset(pa, pb).view(np.uint32)), line[:80]
        else:
            n = int(vals[0]); e = np.array(vals[2:], dtype=np.float32)
            assert len(e) == n
            assert np.float32(vals[1]) == np.float32(ref.mse(e)), line
    assert seen == {"R": 64, "P": 256, "E": 6, "O": 4}


def _point_dtype():
    from oracle import oracle as orc
    return orc.POINT_DTYPE
