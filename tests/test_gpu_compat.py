"""The C++ drop-in layer (include/icpb200/{icp,pointcloud,map}.hpp over the C-ABI): a C++ program written
against the reference's own names runs on the GPU; every output is checked against the CPU oracle."""
import ctypes
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "icp-slam-prototype_b200", "lib", "test_compat")


def _rand_stream(libc, n):
    return np.array([(libc.rand() % 40) == 0 for _ in range(n)], dtype=np.uint8)


def _pts(path, orc):
    return np.fromfile(path, dtype=orc.POINT_DTYPE)


def _same(a, b):
    assert len(a) == len(b), (len(a), len(b))
    assert np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def test_compat_program_matches_oracle(tmp_path, orc):
    from icpb200 import synth
    assert os.path.exists(BIN), "lib/test_compat missing: run __graft_entry__.build()"
    sensor = dict(synth.KINECT_V1)
    prev, cur, col, _ = synth.frame_pair(angle_deg=2.0, shift_m=0.02)
    h, w = prev.shape
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("ii", w, h)); f.write(prev.tobytes()); f.write(cur.tobytes()); f.write(col.tobytes())
    out = tmp_path / "out"; out.mkdir()
    r = subprocess.run([BIN, str(inp), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "compat ok" in r.stdout
    o = str(out)

    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    dec_cur = _rand_stream(libc, int((cur != 0).sum()))
    dec_prev = _rand_stream(libc, int((prev != 0).sum()))
    data, _, c_data = orc.backproject(cur, col, orc.kinect_v1(), orc.SUB_STREAM, 40, 0, dec_cur)
    target, _, c_tgt = orc.backproject(prev, col, orc.kinect_v1(), orc.SUB_STREAM, 40, 0, dec_prev)
    _same(_pts(o + "/data_pts.bin", orc), data)
    _same(_pts(o + "/target_pts.bin", orc), target)
    assert np.array_equal(np.fromfile(o + "/centers.bin", np.float32), np.concatenate([c_data, c_tgt]))

    # makeRotationMatrix (icp.cpp:640-653)
    f32 = np.float32
    # `x * PI / 180` is evaluated in float (PI is a float macro, icp.hpp:4), then widened to double (icp.cpp:641)
    rx, ry, rz = [float(f32(f32(a) * f32(3.14159265358979)) / f32(180)) for a in (3.0, -2.0, 1.0)]
    d = np.array([[1, 0, 0], [0, np.cos(rx), np.sin(rx)], [0, -np.sin(rx), np.cos(rx)]], f32)
    fm = np.array([[np.cos(ry), 0, -np.sin(ry)], [0, 1, 0], [np.sin(ry), 0, np.cos(ry)]], f32)
    g = np.array([[np.cos(rz), np.sin(rz), 0], [-np.sin(rz), np.cos(rz), 0], [0, 0, 1]], f32)
    R = orc.gemm33f(orc.gemm33f(d, fm), g)
    assert np.array_equal(np.fromfile(o + "/rotation.bin", f32).reshape(3, 3), R)
    cam = np.array([5, 5, 5], f32)
    moved = orc.translate(orc.rotate(data, R), cam)
    tgt5 = orc.translate(target, cam)
    _same(_pts(o + "/data_moved.bin", orc), moved)

    idx, dist = orc.nn(moved, tgt5, 8)
    keep = dist < 0.75
    assert np.array_equal(np.fromfile(o + "/errors.bin", f32), dist[keep])
    _same(_pts(o + "/assoc_first.bin", orc), moved[keep])
    _same(_pts(o + "/assoc_second.bin", orc), tgt5[idx][keep])
    s = f32(0)
    for e in dist[keep]:
        s = f32(s + e)
    s = f32(s / f32(keep.sum()))
    mse = f32(np.float64(s) * np.float64(s))
    off = np.zeros(3, f32)
    a_xyz, b_xyz = orc.xyz_of(moved[keep]), orc.xyz_of(tgt5[idx][keep])
    for k in range(len(a_xyz)):
        off = (off + (a_xyz[k] - b_xyz[k])).astype(f32)
    off = (off / f32(keep.sum())).astype(f32)
    scal = np.fromfile(o + "/scalars.bin", f32)
    assert scal[0] == mse and np.array_equal(scal[1:4], off)
    assert scal[4] == dist[0] and scal[5] == dist[0]

    dims = (300, 300, 300)
    cell = float(f32(10.0) / f32(300.0))
    grid = np.zeros(dims, np.uint8)
    firsts = np.ascontiguousarray(moved[keep])
    orc.map_update_endpoints(grid, dims, cell, firsts, orc.RULE_A, 25, 180)
    orc.map_update_endpoints(grid, dims, cell, firsts, orc.RULE_A, 25, 180)
    orc.map_integrate_rays(grid, dims, cell, moved, (5.0, 5.0, 5.0), 25, 25)
    # the three update overloads incl. mapCloud bookkeeping (second map of the C++ program)
    g2 = np.zeros(dims, np.uint8)
    tbl = np.full(dims, -1, np.int32)
    kp_list = []
    app = orc.map_update_tracked(g2, tbl, dims, cell, np.ascontiguousarray(moved[:1500]), 0, 180, 180, 0)
    kp_list.extend(moved[:1500][app])
    non = np.ascontiguousarray(tgt5[:1200])
    for _ in range(7):
        app = orc.map_update_tracked(g2, tbl, dims, cell, non, 2, 25, 180, len(kp_list))
        kp_list.extend(non[app])
    _same(_pts(o + "/mapcloud_kp.bin", orc), np.array(kp_list, dtype=orc.POINT_DTYPE))
    assert len(kp_list) > 1500 - 200 and np.array_equal(np.fromfile(o + "/world2.bin", np.uint8).reshape(dims), g2)

    look = np.fromfile(o + "/lookup.bin", np.int32)
    assert look[0] == 0 and look[1] == 1, "pointLookupTable view (map.hpp:24) does not lead back to the stored points"
    world = np.fromfile(o + "/world.bin", np.uint8).reshape(dims)
    assert np.array_equal(world, grid)
    vox = np.fromfile(o + "/voxel.bin", np.int32)
    v = orc.voxel_coords(orc.xyz_of(moved[:1])[0], cell, dims)
    assert tuple(vox[:3]) == v and vox[3] == int(grid[v] >= 180)

    # getTransformation x2 (icp.cpp:28-285, all-point association), replayed with the oracle
    libc.srand(7)
    d1 = _rand_stream(libc, int((cur != 0).sum())); p1 = _rand_stream(libc, int((prev != 0).sum()))
    dc1 = orc.backproject(cur, col, orc.kinect_v1(), orc.SUB_STREAM, 40, 0, d1)[0]
    pc1 = orc.backproject(prev, col, orc.kinect_v1(), orc.SUB_STREAM, 40, 0, p1)[0]
    I = np.eye(3, dtype=f32)
    tg = orc.translate(orc.rotate(pc1, I), cam)
    da = orc.translate(orc.rotate(dc1, I), cam)
    # the 50 key-points of the C++ program ride along with the points (pointcloud.cpp:321-359) and feed the map (:62, :270)
    from test_live_loop_vs_ref import lift_keypoints
    kps50 = [(20 + (i * 37) % (w - 40), 20 + (i * 53) % (h - 40)) for i in range(50)]
    ap_grid = np.zeros(dims, np.uint8)
    ap_tbl = np.full(dims, -1, np.int32)
    ap_kp = []
    pk1 = orc.translate(orc.rotate(lift_keypoints(orc, prev, col, kps50), I), cam)
    app = orc.map_update_tracked(ap_grid, ap_tbl, dims, cell, pk1, 0, 180, 180, 0)          # icp.cpp:62, map.cpp:220-269
    ap_kp.extend(pk1[app])
    dk1 = orc.translate(orc.rotate(lift_keypoints(orc, cur, col, kps50), I), cam)
    r1, _, dk1m = orc.icp_carry(da, dk1, tg, 3, 0.0, 0.75, orc.SOLVE_REFERENCE, n_threads=8)
    app = orc.map_update_tracked(ap_grid, ap_tbl, dims, cell, dk1m, 0, 25, 180, len(ap_kp))
    ap_kp.extend(dk1m[app])
    assert np.array_equal(np.fromfile(o + "/T1.bin", f32).reshape(4, 4), r1["rigid"])
    camR = orc.gemm33f(I, r1["cam_rotation"])
    camP = (cam + r1["cam_position"]).astype(f32)
    d2 = _rand_stream(libc, int((prev != 0).sum())); p2 = _rand_stream(libc, int((cur != 0).sum()))
    dc2 = orc.backproject(prev, col, orc.kinect_v1(), orc.SUB_STREAM, 40, 0, d2)[0]
    pc2 = orc.backproject(cur, col, orc.kinect_v1(), orc.SUB_STREAM, 40, 0, p2)[0]
    tg2 = orc.translate(orc.rotate(pc2, camR), camP)
    da2 = orc.translate(orc.rotate(dc2, camR), camP)
    dk2 = orc.translate(orc.rotate(lift_keypoints(orc, prev, col, kps50), camR), camP)
    r2, _, dk2m = orc.icp_carry(da2, dk2, tg2, 3, 0.0, 0.75, orc.SOLVE_REFERENCE, last_translation=tuple(-r1["offset"]), n_threads=8)
    app = orc.map_update_tracked(ap_grid, ap_tbl, dims, cell, dk2m, 0, 25, 180, len(ap_kp))
    ap_kp.extend(dk2m[app])
    assert np.array_equal(np.fromfile(o + "/T2.bin", f32).reshape(4, 4), r2["rigid"])
    _same(_pts(o + "/ap_mapkp.bin", orc), np.array(ap_kp, dtype=orc.POINT_DTYPE))
    assert np.array_equal(np.fromfile(o + "/ap_world.bin", np.uint8).reshape(dims), ap_grid)
    assert (ap_grid > 0).sum() > 40
    pose = np.fromfile(o + "/pose.bin", f32)
    assert np.array_equal(pose[:9].reshape(3, 3), orc.gemm33f(camR, r2["cam_rotation"]))
    assert np.array_equal(pose[9:], (camP + r2["cam_position"]).astype(f32))
    # the loop as the reference runs it (icp::setAssociationMode(ASSOCIATE_KEYPOINTS)), two frames from a fresh state
    from test_live_loop_vs_ref import OracleSlam
    kxy = np.fromfile(o + "/live_kxy.bin", f32).reshape(-1, 2)
    kps = [(int(x), int(y)) for x, y in kxy]
    assert len(kps) > 300
    libc.srand(11)
    mine = OracleSlam(orc)
    d_cur = _rand_stream(libc, int((cur != 0).sum())); d_prev = _rand_stream(libc, int((prev != 0).sum()))
    l1 = mine.frame(cur, prev, col, kps, d_cur, d_prev)
    d_cur = _rand_stream(libc, int((prev != 0).sum())); d_prev = _rand_stream(libc, int((cur != 0).sum()))
    l2 = mine.frame(prev, cur, col, kps, d_cur, d_prev)
    assert l1["iterations"] >= 1
    assert np.array_equal(np.fromfile(o + "/L1.bin", f32).reshape(4, 4), l1["rigid"])
    assert np.array_equal(np.fromfile(o + "/L2.bin", f32).reshape(4, 4), l2["rigid"])
    lpose = np.fromfile(o + "/live_pose.bin", f32)
    assert np.array_equal(lpose[:9].reshape(3, 3), mine.camR) and np.array_equal(lpose[9:], mine.camP)
    _same(_pts(o + "/live_mapkp.bin", orc), mine.map_kp)
    assert np.array_equal(np.fromfile(o + "/live_world.bin", np.uint8).reshape(300, 300, 300), mine.grid)
    # pose reporting through the drop-in Quaternion / toEulerianAngle (SLAM.cpp:284-293)
    eu = np.fromfile(o + "/euler.bin", f32)
    q = orc.quat_from_rot(pose[:9])
    assert np.array_equal(eu[0:3], orc.quat_to_euler_deg(q))
    assert np.array_equal(eu[3:6], orc.mat_to_euler_deg(pose[:9]))
    assert np.array_equal(eu[6:10], orc.quat_mul(q, orc.quat_inverse(q)))
