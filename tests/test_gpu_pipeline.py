"""BASELINE config 2 shape on a short prefix: per frame back-projection, ICP on a strided subsample against the
previous frame, full-resolution ray integration into the README grid (300x300x250 at 2 cm) - the whole GPU
pipeline against the same pipeline assembled from oracle calls.  Bar: final grid SHA-256 and poses identical."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_trajectory_prefix_matches_oracle(ctx, orc):
    import icpb200
    from icpb200 import synth
    frames = 4
    dims, cell = (300, 300, 250), 0.02
    poses = synth.trajectory(frames)
    depths = [synth.render_depth(R, t, synth.KINECT_V2, seed=f)[::2, ::2].copy() for f, (R, t) in enumerate(poses)]
    h, w = depths[0].shape
    Kg, Ko = icpb200.reference_intrinsics_v2(), orc.kinect_v2()

    # ---- GPU
    full, sub, prev = ctx.cloud(w * h), ctx.cloud(w * h), ctx.cloud(w * h)
    m = ctx.map(dims, cell)
    R, t = poses[0][0].astype(np.float64), poses[0][1].astype(np.float64)
    gpu_poses = []
    for f in range(frames):
        n = full.from_depth(depths[f], None, Kg)
        stride = max(1, -(-n // 3000))
        sub.from_depth(depths[f], None, Kg, icpb200.SUB_STRIDE, stride)
        sub.transform(R.astype(np.float32), t.astype(np.float32))
        if f > 0:
            res, _, _ = ctx.icp_register(sub, prev, 8, 0.0, 0.75, icpb200.SOLVE_KABSCH)
            R, t = res["pose_R"] @ R, res["pose_R"] @ t + res["pose_t"]
        prev.copy_from(sub)
        full.transform(R.astype(np.float32), t.astype(np.float32))
        m.integrate_rays(full, tuple(float(x) for x in t), 25, 25)
        gpu_poses.append((R.copy(), t.copy()))
    got = m.download()

    # ---- oracle
    grid = np.zeros(dims, np.uint8)
    R, t = poses[0][0].astype(np.float64), poses[0][1].astype(np.float64)
    prev_pts = None
    for f in range(frames):
        full_pts, _, _ = orc.backproject(depths[f], None, Ko)
        stride = max(1, -(-len(full_pts) // 3000))
        sub_pts, _, _ = orc.backproject(depths[f], None, Ko, orc.SUB_STRIDE, stride)
        sub_pts = orc.translate(orc.rotate(sub_pts, R.astype(np.float32)), t.astype(np.float32))
        if f > 0:
            res, sub_pts, _, _ = orc.icp(sub_pts, prev_pts, 8, 0.0, 0.75, orc.SOLVE_KABSCH, n_threads=8)
            R, t = res["pose_R"] @ R, res["pose_R"] @ t + res["pose_t"]
        prev_pts = sub_pts
        world = orc.translate(orc.rotate(full_pts, R.astype(np.float32)), t.astype(np.float32))
        orc.map_integrate_rays(grid, dims, cell, world, tuple(float(x) for x in t), 25, 25)
        assert np.array_equal(gpu_poses[f][0], R) and np.array_equal(gpu_poses[f][1], t), f"pose differs at frame {f}"
    assert hashlib.sha256(got.tobytes()).hexdigest() == hashlib.sha256(grid.tobytes()).hexdigest()
    assert (got > 0).sum() > 1000
    m.close(); full.close(); sub.close(); prev.close()
