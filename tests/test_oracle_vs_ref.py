"""Route B pin: the hand-written oracle against the reference's OWN sources (icp.cpp, pointcloud.cpp,
map.cpp compiled unmodified by path into oracle/_ref).  Bit-exact for points, distances, nearest
points and the certainty grid; pose within tolerance (the N-pair cross-covariance sum differs in order).
Skipped where oracle/_ref has not been built (it needs /root/reference at build time)."""
import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no reference checkout)")


def _same(a, b):
    assert len(a) == len(b)
    assert np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def test_distance_bit_exact(orc):
    rng = np.random.default_rng(0)
    a = orc.make_points(rng.uniform(-9, 9, (3000, 3)))
    b = orc.make_points(rng.uniform(-9, 9, (3000, 3)))
    for i in range(3000):
        assert ref.distance(a[i:i + 1], b[i:i + 1]) == orc.nn(a[i:i + 1], b[i:i + 1])[1][0]


def test_backproject_with_the_references_rand_stream(orc):
    """pointcloud.cpp:109-165 incl. `rand() % 40` (:125): replay its draws through ICPB_SUB_STREAM semantics."""
    from icpb200 import synth
    d0, _, col, _ = synth.frame_pair()
    pts, dec, center = ref.backproject(d0, col, seed=1)
    assert 0.015 * d0.size < len(pts) < 0.035 * d0.size
    mine, cc, cr = orc.backproject(d0, col, orc.kinect_v1(), orc.SUB_STREAM, 40, 0, dec)
    _same(mine, pts)
    assert np.array_equal(cr, center)           # the float running mean, same order
    assert np.allclose(cc, center, atol=1e-5)   # canonical FP64 mean agrees to float accuracy


def test_rotate_translate_bit_exact(orc, pair10k):
    from icpb200 import synth
    data, _ = pair10k
    R = synth.rot_axis_angle([1, -2, 0.5], 0.07).astype(np.float32)
    t = np.array([5, 5, 5], np.float32)
    _same(ref.rotate(data, R), orc.rotate(data, R))
    _same(ref.translate(data, t), orc.translate(data, t))


def test_nearest_and_associations_bit_exact(orc, pair10k):
    data, target = pair10k
    data, target = data[:1200], target[:5000]
    b, d = ref.nearest(data, target)
    idx, dist = orc.nn(data, target, 8)
    assert np.array_equal(d, dist)
    _same(b, target[idx])
    # compacted associations (icp.cpp:553): order preserved, d < 0.75
    a2, b2, e2 = ref.nn_assoc(data, target)
    keep = dist < 0.75
    _same(a2, data[keep]); _same(b2, target[idx][keep]); assert np.array_equal(e2, dist[keep])


def test_nearest_ties_lowest_index(orc):
    g = np.arange(0, 6, dtype=np.float32) * 0.5 + 4
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    target = orc.make_points(np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1))
    target["c0"] = np.arange(len(target)) % 251    # make equal-position points distinguishable
    q = g[:-1] + 0.25
    X, Y, Z = np.meshgrid(q, q, q, indexing="ij")
    data = orc.make_points(np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1))
    b, d = ref.nearest(data, target)
    idx, dist = orc.nn(data, target)
    assert np.array_equal(d, dist)
    _same(b, target[idx])


def test_mse_and_rotation_matrix(orc):
    rng = np.random.default_rng(1)
    e = rng.uniform(0, 0.7, 1000).astype(np.float32)
    # reference: sequential float sum; oracle: canonical FP64 sum -> equal to float accuracy
    s = np.float32(0)
    for v in e:
        s = np.float32(s + v)
    m = np.float32(s / np.float32(len(e)))
    assert ref.mse(e) == np.float32(np.float64(m) * np.float64(m))
    assert np.array_equal(ref.make_rotation(0, 0, 0), np.eye(3, dtype=np.float32))


def test_voxel_coordinates(orc):
    rng = np.random.default_rng(2)
    cell = float(np.float32(10.0) / np.float32(300.0))
    pts = np.concatenate([rng.uniform(-1, 11, (3000, 3)), [[0, 0, 0], [10, 10, 10], [-0.01, 9.9999, 3.3333333]]])
    for p in pts.astype(np.float32):
        assert ref.voxel(p) == orc.voxel_coords(p, cell, (300, 300, 300))


@pytest.mark.parametrize("kind,rule,delta", [("cloud", 0, 180), ("cloud", 0, 25), ("nonassoc", 1, 25), ("assoc", 0, 25)])
def test_map_updates_bit_exact(orc, pair10k, kind, rule, delta):
    """Map::update overloads (map.cpp:88-119, 122-151, 220-269) on the reference's 300^3 grid."""
    data, target = pair10k
    cell = float(np.float32(10.0) / np.float32(300.0))
    ref.map_reset()
    grid = np.zeros((300, 300, 300), np.uint8)
    for rep in range(4):
        pts = data if rep % 2 == 0 else target
        ref.map_update(pts, delta, kind)
        orc.map_update_endpoints(grid, (300, 300, 300), cell, pts, rule, delta, 180)
    w = ref.map_world()
    assert np.array_equal(w, grid)
    assert w.max() == 255
    ref.map_reset()


def test_registration_loop_pose_within_tolerance(orc, pair10k):
    """icp.cpp:155-258 driven through the reference's own functions vs the oracle's canonical loop."""
    data, target = pair10k
    data, target = data[:2500], target[:3000]
    r, rout = ref.icp_allpoints(data, target, 6, 0.0)
    o, oout, it, dt = orc.icp(data, target, 6, 0.0, 0.75, orc.SOLVE_REFERENCE, n_threads=8, trace=True)
    assert r["iterations"] == o["iterations"] == 6
    assert r["n_assoc"] == o["n_assoc"]
    assert np.abs(r["rigid"] - o["rigid"]).max() < 1e-5
    assert np.abs(r["cam_rotation"] - o["cam_rotation"]).max() < 1e-5
    assert np.abs(r["cam_position"] - o["cam_position"]).max() < 1e-5
    assert abs(r["mse"] - o["mse"]) < 1e-7
    assert np.abs(orc.xyz_of(rout) - orc.xyz_of(oout)).max() < 1e-5


@pytest.mark.parametrize("kind,variant,delta,which", [("cloud", 0, 180, 0), ("cloud", 0, 25, 0), ("nonassoc", 2, 25, 0),
                                                      ("assoc", 1, 25, 1)])
def test_map_cloud_bookkeeping_bit_exact(orc, pair10k, kind, variant, delta, which):
    """pointLookupTable / mapCloud insertion of the Map::update overloads (map.cpp:104-110, 142-145, 256-259):
    which points are appended, and in which order."""
    data, target = pair10k
    cell = float(np.float32(10.0) / np.float32(300.0))
    dims = (300, 300, 300)
    ref.map_reset()
    grid = np.zeros(dims, np.uint8)
    table = np.full(dims, -1, np.int32)
    mine = []
    for rep in range(12):
        pts = np.ascontiguousarray((data if rep % 2 == 0 else target)[rep * 37: rep * 37 + 2500])
        ref.map_update(pts, delta, kind)
        app = orc.map_update_tracked(grid, table, dims, cell, pts, variant, delta, 180, len(mine))
        mine.extend(pts[app])
    got = ref.map_cloud(which)
    assert len(got) == len(mine) and len(mine) > 0
    assert np.array_equal(got.view(np.uint8), np.array(mine, dtype=orc.POINT_DTYPE).view(np.uint8))
    assert np.array_equal(ref.map_world(), grid)
    ref.map_reset()
