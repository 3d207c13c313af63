"""SURVEY.md 8f-2 pinned against the reference itself: the reference's OWN icp::getTransformation (icp.cpp:28-285,
live key-point variant, compiled unmodified by path -- Route B) is run over a short synthetic sequence, and the
oracle's composition of the same frame (back-projection with the replayed rand() draws, key-point lifting, the
key-point loop, rule-A / rule-C map updates) must land on the same camera pose, map cloud and certainty grid.
The GPU path is bit-exact against this oracle (tests/test_gpu_keypoints.py).  CPU only."""
import ctypes

import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")

FX, CX = np.float32(468.60), np.float32(318.27)   # pointcloud.hpp:7-10 (y uses CX / FX too, pointcloud.cpp:86-87)
DIMS, CELL = (300, 300, 300), np.float32(np.float32(10.0) / np.float32(300.0))


def lift_keypoints(orc, depth, bgr, kps):
    """pointcloud.cpp:64-97: the key-point's pixel, skipped when its depth is zero; same float formula as the points."""
    out = []
    for (x, y) in kps:
        d = depth[y, x]
        if d == 0:
            continue
        pz = np.float32(d) / np.float32(5000.0)
        px = (np.float32(x) - CX) * pz / FX
        py = (np.float32(y) - CX) * pz / FX
        out.append((px, py, pz, bgr[y, x, 0], bgr[y, x, 1], bgr[y, x, 2], 0))
    return np.array(out, dtype=orc.POINT_DTYPE)


class OracleSlam:
    """icp.cpp:22-26 globals + getTransformation, composed from oracle functions."""

    def __init__(self, orc):
        self.orc = orc
        self.grid = np.zeros(DIMS, np.uint8)
        self.table = np.full(DIMS, -1, np.int32)
        self.map_kp = np.zeros(0, orc.POINT_DTYPE)
        self.map_pts = np.zeros(0, orc.POINT_DTYPE)
        self.camR = np.eye(3, dtype=np.float32)
        self.camP = np.zeros(3, np.float32)
        self.last_t = np.zeros(3, np.float32)

    def frame(self, cur, prev, bgr, kps, dec_cur, dec_prev, max_it=16, thr=1e-4):
        orc = self.orc
        K = orc.kinect_v1()
        d_pts = orc.backproject(cur, bgr, K, orc.SUB_STREAM, 40, 0, dec_cur)[0]
        p_pts = orc.backproject(prev, bgr, K, orc.SUB_STREAM, 40, 0, dec_prev)[0]
        d_kp, p_kp = lift_keypoints(orc, cur, bgr, kps), lift_keypoints(orc, prev, bgr, kps)
        if len(self.map_pts) == 0:                                             # icp.cpp:47-68
            self.camR = np.eye(3, dtype=np.float32); self.camP = np.array([5, 5, 5], np.float32)
            self.last_t = np.zeros(3, np.float32)
            p_pts = orc.translate(orc.rotate(p_pts, self.camR), self.camP)
            p_kp = orc.translate(orc.rotate(p_kp, self.camR), self.camP)
            app = orc.map_update_tracked(self.grid, self.table, DIMS, CELL, p_kp, 0, 180, 180, len(self.map_kp))
            self.map_kp = np.concatenate([self.map_kp, p_kp[app]])
            self.map_pts = p_pts
        d_pts = orc.translate(orc.rotate(d_pts, self.camR), self.camP)         # :70-71
        d_kp = orc.translate(orc.rotate(d_kp, self.camR), self.camP)
        r, _, _, non = orc.icp_keypoints(d_kp, d_pts, self.map_kp, max_it, thr, 0.1,
                                         last_translation=tuple(self.last_t), n_threads=4)
        self.camR = orc.gemm33f(self.camR, r["cam_rotation"])                  # :237 accumulated
        self.camP = (self.camP + r["cam_position"]).astype(np.float32)         # :246 accumulated
        self.last_t = (-r["offset"]).astype(np.float32)                        # :260
        if r["n_assoc"] > 0:                                                   # map.cpp:124-126
            app = orc.map_update_tracked(self.grid, self.table, DIMS, CELL, non, 2, 25, 180, len(self.map_kp))
            self.map_kp = np.concatenate([self.map_kp, non[app]])
        return r


def _rand_stream(libc, n):
    return np.array([(libc.rand() % 40) == 0 for _ in range(n)], dtype=np.uint8)


def _keypoint_pixels(depth_frames, n, seed):
    """Stand-in for cv::FAST (out of scope): seeded pixels that have depth in every frame, integer coordinates."""
    rng = np.random.default_rng(seed)
    ok = np.ones_like(depth_frames[0], bool)
    for d in depth_frames:
        ok &= d != 0
    ys, xs = np.nonzero(ok[8:-8, 8:-8])
    sel = rng.choice(len(xs), n, replace=False)
    return [(int(xs[i]) + 8, int(ys[i]) + 8) for i in sel]


def test_live_keypoint_loop_matches_the_references_getTransformation(orc):
    from icpb200 import synth
    poses = synth.trajectory(4, step_deg=0.6, step_m=0.012)
    frames = [synth.render_depth(R, t, synth.KINECT_V1, seed=f) for f, (R, t) in enumerate(poses)]
    bgr = np.random.default_rng(3).integers(0, 255, frames[0].shape + (3,), dtype=np.uint8)
    kps = _keypoint_pixels(frames, 600, 7)
    kxy = np.array(kps, np.float32)

    ref.map_reset()
    libc = ctypes.CDLL("libc.so.6")
    mine = OracleSlam(orc)
    for f in range(1, len(frames)):
        cur, prev = frames[f], frames[f - 1]
        seed = 100 + f
        T_ref, camR_ref, camP_ref = ref.get_transformation(cur, prev, bgr, kxy, 16, 1e-4, seed)
        libc.srand(seed)
        dec_cur = _rand_stream(libc, int((cur != 0).sum()))
        dec_prev = _rand_stream(libc, int((prev != 0).sum()))
        r = mine.frame(cur, prev, bgr, kps, dec_cur, dec_prev)
        # pose: the reference sums M and the offset in float, sequentially; the oracle in the canonical FP64 order
        assert np.abs(mine.camR - camR_ref).max() < 1e-5, (f, np.abs(mine.camR - camR_ref).max())
        assert np.abs(mine.camP - camP_ref).max() < 1e-5, (f, np.abs(mine.camP - camP_ref).max())
        assert np.abs(r["rigid"][:3] - T_ref[:3]).max() < 1e-5
        assert r["iterations"] >= 1
    ref_kp = ref.map_cloud(0)
    ref_pts = ref.map_cloud(1)
    assert len(ref_pts) == len(mine.map_pts)
    assert np.array_equal(ref_pts.view(np.uint8), mine.map_pts.view(np.uint8))
    # map cloud key-points: same count; same points up to the float noise of the pose
    assert len(ref_kp) == len(mine.map_kp), (len(ref_kp), len(mine.map_kp))
    dx = np.abs(orc.xyz_of(ref_kp) - orc.xyz_of(mine.map_kp)).max()
    assert dx < 1e-4, dx
    world = ref.map_world()
    diff = int((world != mine.grid).sum())
    assert diff <= 4, diff   # a reject within float noise of a voxel wall may land next door
    assert int((mine.grid > 0).sum()) > 500


def test_live_loop_fewer_than_three_associations_matches_the_reference(orc):
    """icp.cpp:163-182 inside the real loop: a frame that brings only two key-points replays the last translation on
    the data cloud, leaves the camera pose alone and reports -lastTranslation as the offset -- checked against the
    reference's own getTransformation, frame after a normal frame."""
    from icpb200 import synth
    poses = synth.trajectory(4, step_deg=0.6, step_m=0.012)
    frames = [synth.render_depth(R, t, synth.KINECT_V1, seed=f) for f, (R, t) in enumerate(poses)]
    bgr = np.random.default_rng(5).integers(0, 255, frames[0].shape + (3,), dtype=np.uint8)
    kps_all = _keypoint_pixels(frames, 400, 9)
    per_frame = {1: kps_all, 2: kps_all[:2], 3: kps_all}

    ref.map_reset()
    libc = ctypes.CDLL("libc.so.6")
    mine = OracleSlam(orc)
    seen_small = False
    for f in range(1, len(frames)):
        cur, prev, kps = frames[f], frames[f - 1], per_frame[f]
        seed = 300 + f
        T_ref, camR_ref, camP_ref = ref.get_transformation(cur, prev, bgr, np.array(kps, np.float32), 16, 1e-4, seed)
        libc.srand(seed)
        dec_cur = _rand_stream(libc, int((cur != 0).sum()))
        dec_prev = _rand_stream(libc, int((prev != 0).sum()))
        r = mine.frame(cur, prev, bgr, kps, dec_cur, dec_prev)
        seen_small |= bool(r["small_assoc_exit"])
        assert np.abs(mine.camR - camR_ref).max() < 1e-5 and np.abs(mine.camP - camP_ref).max() < 1e-5, f
        assert np.abs(r["rigid"][:3, 3] - T_ref[:3, 3]).max() < 1e-5, f          # offset column (icp.cpp:266-268)
    assert seen_small
    assert len(ref.map_cloud(0)) == len(mine.map_kp)
    assert int((ref.map_world() != mine.grid).sum()) <= 4
