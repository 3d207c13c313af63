"""SURVEY.md 8f-2 -- the loop as the reference runs it (icp.cpp:98,155-258): key-points associated with the map
cloud's key-points, the cloud's points carried along, rejects accumulated over the passes, rule-C map update.
GPU (icpb_icp_register_keypoints + icpb_map_update_tracked) vs the CPU oracle.  Bar: bit-exact."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

pytestmark = pytest.mark.gpu


def _same(a, b):
    assert len(a) == len(b), (len(a), len(b))
    assert np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def _scene(orc, seed, n_map=800, n_kp=300, n_out=50, n_pts=2000, deg=(1, 2, -1), shift=0.01):
    rng = np.random.default_rng(seed)
    mk = orc.make_points(rng.uniform(4, 6, (n_map, 3)))
    R = Rotation.from_euler("xyz", deg, degrees=True).as_matrix()
    sel = rng.choice(n_map, n_kp, replace=False)
    xyz = (np.stack([mk["x"], mk["y"], mk["z"]], 1)[sel] - 5) @ R.T + 5 + shift
    kp = orc.make_points(np.concatenate([xyz, rng.uniform(4, 6, (n_out, 3))]))
    for a in (mk, kp):
        a["c0"] = rng.integers(0, 255, len(a)); a["c1"] = rng.integers(0, 255, len(a)); a["c2"] = rng.integers(0, 255, len(a))
    pts = orc.make_points(rng.uniform(4, 6, (n_pts, 3)))
    return kp, pts, mk


def _run(ctx, orc, kp, pts, mk, **kw):
    import icpb200
    kc, pc, mc = ctx.cloud_from_points(kp), ctx.cloud_from_points(pts), ctx.cloud_from_points(mk)
    it = kw.get("max_iterations", 16)
    non = ctx.cloud((it + 1) * max(len(kp), 1))
    res = ctx.icp_register_keypoints(kc, pc, mc, non_associations=non, **kw)
    out = (res, kc.download(), pc.download(), non.download())
    for c in (kc, pc, mc, non):
        c.close()
    ref = orc.icp_keypoints(kp, pts, mk, n_threads=4, **kw)
    return out, ref


def _compare(out, ref):
    (res, kp2, pts2, non), (r, rkp, rpts, rnon) = out, ref
    for k in ("iterations", "nn_passes", "n_assoc", "small_assoc_exit"):
        assert res[k] == r[k], k
    assert np.float32(res["mse"]).view(np.uint32) == np.float32(r["mse"]).view(np.uint32)
    for k in ("rigid", "cam_rotation", "cam_position", "offset", "pose_R", "pose_t"):
        assert np.array_equal(res[k], r[k]), k
    _same(kp2, rkp)
    _same(pts2, rpts)
    _same(non, rnon)
    assert res["n_nonassoc"] == len(rnon)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_keypoint_loop_matches_oracle(ctx, orc, seed):
    kp, pts, mk = _scene(orc, seed)
    out, ref = _run(ctx, orc, kp, pts, mk)
    _compare(out, ref)
    assert out[0]["iterations"] > 1 and len(out[3]) > 0


def test_keypoint_loop_converges_early(ctx, orc):
    """threshold 1e-4 = mean distance 1 cm (SLAM.cpp:277): a 2 mm offset exits after the first association."""
    kp, pts, mk = _scene(orc, 3, deg=(0, 0, 0), shift=0.002, n_out=5)
    out, ref = _run(ctx, orc, kp, pts, mk)
    _compare(out, ref)
    assert out[0]["iterations"] == 0 and out[0]["nn_passes"] == 1


def test_keypoint_loop_fewer_than_three_associations(ctx, orc):
    """icp.cpp:163-182: the last motion is replayed -- on the carried points as well."""
    kp, pts, mk = _scene(orc, 4, n_kp=2, n_out=0)
    out, ref = _run(ctx, orc, kp, pts, mk, last_translation=(0.01, -0.02, 0.03))
    _compare(out, ref)
    assert out[0]["small_assoc_exit"] == 1
    assert not np.array_equal(out[2]["x"], pts["x"])


def test_keypoint_loop_empty_map_touches_nothing(ctx, orc):
    import icpb200
    kp, pts, _ = _scene(orc, 5)
    kc, pc = ctx.cloud_from_points(kp), ctx.cloud_from_points(pts)
    mc = ctx.cloud(8)                                # icp.cpp:490-491
    non = ctx.cloud(17 * len(kp))
    res = ctx.icp_register_keypoints(kc, pc, mc, non_associations=non)
    assert res["nn_passes"] == 0 and res["iterations"] == 0 and non.n == 0
    _same(kc.download(), kp)
    _same(pc.download(), pts)
    assert np.array_equal(res["pose_R"], np.eye(3))


def test_frame_flow_with_rule_c_map_update(ctx, orc):
    """One frame of the live loop: associate + move on the device, then Map::update(assoc, errors, nonAssoc, 25)
    (map.cpp:122-151) on the accumulated rejects: grid, lookup-table insertions and mapCloud growth equal the oracle."""
    import icpb200
    kp, pts, mk = _scene(orc, 6, n_map=1200, n_kp=400, n_out=120)
    dims, cell = (300, 300, 300), np.float32(10.0 / 300.0)
    # init: Map::update(PointCloud, 180) on the map key-points (map.cpp:220-269)
    m = ctx.map(dims, float(cell))
    grid = np.zeros(dims, np.uint8); table = np.full(dims, -1, np.int32)
    mapc = ctx.cloud(1 << 16)
    app0 = m.update_tracked(ctx.cloud_from_points(mk), icpb200.TRACK_INIT, 180, 180, mapc)
    r_app0 = orc.map_update_tracked(grid, table, dims, cell, mk, 0, 180, 180, 0)
    assert app0 == len(r_app0)
    map_kp = mk[r_app0]
    _same(mapc.download(), map_kp)
    # the frame
    kc, pc = ctx.cloud_from_points(kp), ctx.cloud_from_points(pts)
    non = ctx.cloud(17 * len(kp))
    res = ctx.icp_register_keypoints(kc, pc, mapc, non_associations=non)
    r, rkp, rpts, rnon = orc.icp_keypoints(kp, pts, map_kp, n_threads=4)
    _same(non.download(), rnon)
    assert res["n_assoc"] == r["n_assoc"] > 0
    app1 = m.update_tracked(non, icpb200.TRACK_NONASSOC, 25, 180, mapc)
    r_app1 = orc.map_update_tracked(grid, table, dims, cell, rnon, 2, 25, 180, len(map_kp))
    assert app1 == len(r_app1)
    _same(mapc.download(), np.concatenate([map_kp, rnon[r_app1]]))
    assert np.array_equal(m.download(), grid)
