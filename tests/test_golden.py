"""Golden vectors frozen from the reference's own code (tests/golden/make_golden.py, `ref_*` arrays) and the
oracle's canonical results (`orc_*`).  CPU tests check the oracle against them; `-m gpu` tests check the CUDA
path through the C-ABI.  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_DIMS = (300, 300, 300)
REF_CELL = float(np.float32(10.0) / np.float32(300.0))   # map.hpp:9-10,17


def _load(name):
    return np.load(os.path.join(G, name))


def _same(a, b):
    assert len(a) == len(b)
    assert np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def _dense(z, tag):
    w = np.zeros(REF_DIMS, np.uint8)
    w.ravel()[z[f"ref_{tag}_idx"]] = z[f"ref_{tag}_val"]
    return w


MAP_CASES = [("A180", 0, 180), ("A25", 0, 25), ("C25", 1, 25)]

# ---------------------------------------------------------------- CPU: oracle vs the reference's vectors


def test_oracle_nn_golden(orc):
    z = _load("nn_small.npz")
    idx, dist = orc.nn(z["data"], z["target"])
    assert np.array_equal(dist, z["ref_dist"])
    _same(z["target"][idx], z["ref_nearest"])
    assert np.array_equal(idx, z["idx"])


def test_oracle_backproject_golden(orc):
    z = _load("backproject_small.npz")
    pts, cc, cr = orc.backproject(z["depth"], z["bgr"], orc.kinect_v1(), orc.SUB_STREAM, 40, 0, z["ref_decisions"])
    _same(pts, z["ref_points"])
    assert np.array_equal(cr, z["ref_center"])
    _same(orc.backproject(z["depth"], z["bgr"])[0], z["orc_all_points"])


@pytest.mark.parametrize("tag,rule,delta", MAP_CASES)
def test_oracle_map_golden(orc, tag, rule, delta):
    z = _load("map_small.npz")
    grid = np.zeros(REF_DIMS, np.uint8)
    for _ in range(3):
        orc.map_update_endpoints(grid, REF_DIMS, REF_CELL, z["points"], rule, delta, 180)
    assert np.array_equal(grid, _dense(z, tag))


def test_oracle_icp_golden(orc):
    z = _load("icp_small.npz")
    o, out, it, dt = orc.icp(z["data"], z["target"], 5, 0.0, 0.75, orc.SOLVE_REFERENCE, trace=True)
    # against the reference's own loop: tolerance (the cross-covariance summation order differs)
    assert np.abs(o["rigid"] - z["ref_rigid"]).max() < 1e-5
    assert np.abs(o["cam_rotation"] - z["ref_cam_rotation"]).max() < 1e-5
    assert np.abs(o["cam_position"] - z["ref_cam_position"]).max() < 1e-5
    assert np.abs(orc.xyz_of(out) - orc.xyz_of(z["ref_out"])).max() < 1e-5
    assert o["n_assoc"] == int(z["ref_n_assoc"])
    # against its own frozen canonical results: bit-exact
    assert np.array_equal(it, z["orc_refmode_idx"]) and np.array_equal(dt, z["orc_refmode_dist"])
    _same(out, z["orc_refmode_out"])
    k, kout, kit, kdt = orc.icp(z["data"], z["target"], 5, 0.0, 0.75, orc.SOLVE_KABSCH, trace=True)
    assert np.array_equal(kit, z["orc_kabsch_idx"])
    _same(kout, z["orc_kabsch_out"])


def test_oracle_rays_golden(orc):
    z = _load("rays_small.npz")
    grid = z["start"].copy()
    v = orc.map_integrate_rays(grid, tuple(int(d) for d in z["dims"]), float(z["cell"]), z["points"],
                               tuple(float(x) for x in z["origin"]), 25, 25)
    assert v == int(z["orc_visited"]) and np.array_equal(grid, z["orc_grid"])


# ---------------------------------------------------------------- GPU: CUDA path vs the same vectors


@pytest.mark.gpu
def test_gpu_nn_golden(ctx):
    z = _load("nn_small.npz")
    dc, tc = ctx.cloud_from_points(z["data"]), ctx.cloud_from_points(z["target"])
    idx, dist, _ = ctx.nn_search(dc, tc)
    assert np.array_equal(idx, z["idx"]) and np.array_equal(dist, z["ref_dist"])
    _same(z["target"][idx], z["ref_nearest"])
    dc.close(); tc.close()


@pytest.mark.gpu
def test_gpu_backproject_golden(ctx):
    import icpb200
    z = _load("backproject_small.npz")
    c = ctx.cloud(z["depth"].size)
    c.from_depth(z["depth"], z["bgr"], None, icpb200.SUB_STREAM, 40, 0, z["ref_decisions"])
    _same(c.download(), z["ref_points"])
    c.from_depth(z["depth"], z["bgr"], None)
    _same(c.download(), z["orc_all_points"])
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,rule,delta", MAP_CASES)
def test_gpu_map_golden(ctx, tag, rule, delta):
    z = _load("map_small.npz")
    m = ctx.map(REF_DIMS, REF_CELL)
    c = ctx.cloud_from_points(z["points"])
    for _ in range(3):
        m.update_endpoints(c, rule, delta, 180)
    assert np.array_equal(m.download(), _dense(z, tag))
    m.close(); c.close()


@pytest.mark.gpu
def test_gpu_icp_golden(ctx):
    z = _load("icp_small.npz")
    for mode, tag in [(0, "refmode"), (1, "kabsch")]:
        dc, tc = ctx.cloud_from_points(z["data"]), ctx.cloud_from_points(z["target"])
        res, it, dt = ctx.icp_register(dc, tc, 5, 0.0, 0.75, mode, trace=True)
        assert np.array_equal(it, z[f"orc_{tag}_idx"]) and np.array_equal(dt, z[f"orc_{tag}_dist"])
        _same(dc.download(), z[f"orc_{tag}_out"])
        assert np.array_equal(res["pose_R"], z[f"orc_{tag}_pose_R"]) and np.array_equal(res["pose_t"], z[f"orc_{tag}_pose_t"])
        assert np.array_equal(res["rigid"], z[f"orc_{tag}_rigid"])
        if mode == 0:   # and against the reference's own loop, within the north-star tolerance
            assert np.abs(res["rigid"] - z["ref_rigid"]).max() < 1e-5
            assert np.abs(res["cam_position"] - z["ref_cam_position"]).max() < 1e-5
        dc.close(); tc.close()


@pytest.mark.gpu
def test_gpu_rays_golden(ctx):
    z = _load("rays_small.npz")
    m = ctx.map(tuple(int(d) for d in z["dims"]), float(z["cell"]))
    m.upload(z["start"])
    c = ctx.cloud_from_points(z["points"])
    v = m.integrate_rays(c, tuple(float(x) for x in z["origin"]), 25, 25)
    assert v == int(z["orc_visited"]) and np.array_equal(m.download(), z["orc_grid"])
    m.close(); c.close()
