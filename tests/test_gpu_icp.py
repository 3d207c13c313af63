"""GPU parity: the device-resident registration loop (S1-S3 + N1-N3 + P2) vs the CPU oracle.
Bars: nearest indices bit-exact on every pass; pose within 1e-5 rad / 1e-5 m (north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL_RAD = 1e-5
TRANS_TOL_M = 1e-5


def rot_angle(Ra, Rb):
    """Rotation angle between two (nearly) orthonormal matrices.  sin(angle) from the skew part of
    Ra^T Rb: well conditioned at small angles, unlike arccos((trace-1)/2) (the composed pose is a
    product of float-rounded factors, so trace(R^T R) is 3 - O(1e-7), not 3)."""
    D = Ra.T @ Rb
    S = (D - D.T) / 2.0
    s = np.sqrt(S[0, 1] ** 2 + S[0, 2] ** 2 + S[1, 2] ** 2)
    c = (np.trace(D) - 1.0) / 2.0
    return float(np.arctan2(s, c))


def _run(ctx, orc, data, target, mode, iters=20, threshold=0.0, max_d=0.75, last_t=(0, 0, 0)):
    dc = ctx.cloud_from_points(data)
    tc = ctx.cloud_from_points(target)
    res, it, dt = ctx.icp_register(dc, tc, iters, threshold, max_d, mode, last_t, trace=True)
    out = dc.download()
    dc.close(); tc.close()
    ref, rout, rit, rdt = orc.icp(data, target, iters, threshold, max_d, mode, last_t, n_threads=8, trace=True)
    return res, it, dt, out, ref, rit, rdt, rout


@pytest.mark.parametrize("mode", [0, 1])
def test_icp_config1_trace_and_pose(ctx, orc, pair10k, mode):
    data, target = pair10k
    res, it, dt, out, ref, rit, rdt, rout = _run(ctx, orc, data, target, mode)
    assert res["iterations"] == ref["iterations"] == 20
    assert res["nn_passes"] == ref["nn_passes"] == 21
    for k in range(21):
        assert np.array_equal(it[k], rit[k]), f"pass {k}: {(it[k] != rit[k]).sum()} index mismatches"
        assert np.array_equal(dt[k].view(np.uint32), rdt[k].view(np.uint32)), f"pass {k}: distances differ"
    assert res["n_assoc"] == ref["n_assoc"]
    assert rot_angle(res["pose_R"], ref["pose_R"]) <= ROT_TOL_RAD
    assert np.abs(res["pose_t"] - ref["pose_t"]).max() <= TRANS_TOL_M
    # canonical arithmetic: in fact bit-equal
    assert np.array_equal(res["pose_R"], ref["pose_R"]) and np.array_equal(res["pose_t"], ref["pose_t"])
    assert np.array_equal(res["rigid"], ref["rigid"])
    assert np.array_equal(res["cam_rotation"], ref["cam_rotation"])
    assert np.array_equal(res["cam_position"], ref["cam_position"])
    assert res["mse"] == ref["mse"]
    assert np.array_equal(out.view(np.uint8), rout.view(np.uint8)), "transformed cloud differs"


def test_icp_threshold_early_exit(ctx, orc, pair10k):
    data, target = pair10k
    data, target = data[:3000], target[:4000]
    res, it, dt, out, ref, rit, rdt, rout = _run(ctx, orc, data, target, 0, iters=16, threshold=5e-3)
    assert 0 < ref["iterations"] < 16
    assert res["iterations"] == ref["iterations"] and res["nn_passes"] == ref["nn_passes"]
    assert np.array_equal(it[: ref["nn_passes"]], rit[: ref["nn_passes"]])
    assert np.array_equal(out.view(np.uint8), rout.view(np.uint8))


def test_icp_less_than_three_associations(ctx, orc):
    """icp.cpp:163-182: < 3 associations replays the last motion and stops."""
    rng = np.random.default_rng(3)
    target = orc.make_points(rng.uniform(4, 6, (300, 3)))
    data = orc.make_points(np.concatenate([rng.uniform(40, 60, (200, 3)), rng.uniform(4, 6, (2, 3))]))
    res, it, dt, out, ref, rit, rdt, rout = _run(ctx, orc, data, target, 0, iters=8, last_t=(0.01, -0.02, 0.03))
    assert ref["small_assoc_exit"] == 1 and res["small_assoc_exit"] == 1
    assert res["iterations"] == ref["iterations"] == 8 and res["nn_passes"] == ref["nn_passes"] == 1
    assert np.array_equal(out.view(np.uint8), rout.view(np.uint8))
    assert np.array_equal(res["offset"], ref["offset"])


def test_icp_zero_associations(ctx, orc):
    rng = np.random.default_rng(4)
    target = orc.make_points(rng.uniform(4, 6, (100, 3)))
    data = orc.make_points(rng.uniform(40, 60, (100, 3)))
    res, it, dt, out, ref, rit, rdt, rout = _run(ctx, orc, data, target, 0, iters=5)
    assert res["n_assoc"] == ref["n_assoc"] == 0 and res["iterations"] == ref["iterations"] == 0
    assert np.array_equal(out.view(np.uint8), data.view(np.uint8))


def test_icp_recovers_known_motion_kabsch(ctx, orc):
    """Kabsch mode on a well-conditioned cloud recovers a known rigid motion (rigid_transform_3D.py:42-97)."""
    from icpb200 import synth
    rng = np.random.default_rng(11)
    base = rng.uniform(3, 7, (4000, 3))
    R = synth.rot_axis_angle([0.3, -0.5, 0.8], np.deg2rad(2.0))
    t = np.array([0.02, -0.01, 0.015])
    moved = (base - 5.0) @ R.T + 5.0 + t
    target = orc.make_points(moved)
    data = orc.make_points(base)
    dc = ctx.cloud_from_points(data); tc = ctx.cloud_from_points(target)
    res, _, _ = ctx.icp_register(dc, tc, 30, 0.0, 0.75, 1)
    out = orc.xyz_of(dc.download())
    assert np.abs(out - moved).max() < 1e-4
    dc.close(); tc.close()


def test_icp_batch_matches_single(ctx, orc, pair10k):
    """BASELINE config 4 shape: a batch of independent registrations == the same registrations one by one."""
    data, target = pair10k
    sizes = [(2000, 2500), (1500, 1500), (2048, 1000), (777, 3001)]
    singles, datas, targets = [], [], []
    for k, (n, m) in enumerate(sizes):
        d = data[k * 100: k * 100 + n]; t = target[k * 50: k * 50 + m]
        dc = ctx.cloud_from_points(d); tc = ctx.cloud_from_points(t)
        r, _, _ = ctx.icp_register(dc, tc, 6, 0.0, 0.75, 0)
        singles.append((r, dc.download()))
        dc.close()
        datas.append(ctx.cloud_from_points(d)); targets.append(tc)
    res = ctx.icp_register_batch(datas, targets, 6, 0.0, 0.75, 0)
    for k in range(len(sizes)):
        assert np.array_equal(res[k]["pose_R"], singles[k][0]["pose_R"])
        assert np.array_equal(res[k]["pose_t"], singles[k][0]["pose_t"])
        assert res[k]["n_assoc"] == singles[k][0]["n_assoc"]
        assert np.array_equal(datas[k].download().view(np.uint8), singles[k][1].view(np.uint8))
    for c in datas + targets:
        c.close()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("cell", [0.0, 0.11])
def test_icp_grid_mode_is_exact(ctx, orc, pair10k, mode, cell):
    """ICPB_NN_GRID (the reference's intended voxel-indexed scan, icp.cpp:347-486, made exact): every accepted
    association, every sum, the pose and the transformed cloud are bit-identical to the brute-force oracle."""
    import icpb200
    data, target = pair10k
    dc, tc = ctx.cloud_from_points(data), ctx.cloud_from_points(target)
    res, it, dt = ctx.icp_register(dc, tc, 20, 0.0, 0.75, mode, trace=True, nn_mode=icpb200.NN_GRID, grid_cell=cell)
    out = dc.download()
    dc.close(); tc.close()
    ref, rout, rit, rdt = orc.icp(data, target, 20, 0.0, 0.75, mode, n_threads=8, trace=True)
    assert res["nn_passes"] == ref["nn_passes"] == 21
    for k in range(21):
        acc = rdt[k] < 0.75
        assert np.array_equal(it[k][acc], rit[k][acc]), f"pass {k}"
        assert np.array_equal(dt[k][acc], rdt[k][acc]), f"pass {k}"
        assert np.all(it[k][~acc] == -1) and np.all(np.isinf(dt[k][~acc]))
    assert res["n_assoc"] == ref["n_assoc"] and res["mse"] == ref["mse"]
    assert np.array_equal(res["pose_R"], ref["pose_R"]) and np.array_equal(res["pose_t"], ref["pose_t"])
    assert np.array_equal(res["rigid"], ref["rigid"])
    assert np.array_equal(out.view(np.uint8), rout.view(np.uint8))


def test_icp_grid_mode_partial_overlap_and_ties(ctx, orc):
    """Grid mode with rejected queries (no neighbour inside 0.75 m), queries outside the target's bounding box,
    duplicated targets and lattice ties."""
    import icpb200
    rng = np.random.default_rng(21)
    g = np.arange(0, 12, dtype=np.float32) * 0.1 + 4.0
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    lat = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    tgt = np.concatenate([lat, lat[:300], rng.uniform(4, 5.2, (1500, 3)).astype(np.float32)])
    dat = np.concatenate([lat[:500] + np.float32(0.05), rng.uniform(2.5, 7.5, (2000, 3)).astype(np.float32),
                          rng.uniform(20, 30, (50, 3)).astype(np.float32)])
    data, target = orc.make_points(dat), orc.make_points(tgt)
    dc, tc = ctx.cloud_from_points(data), ctx.cloud_from_points(target)
    res, it, dt = ctx.icp_register(dc, tc, 4, 0.0, 0.75, 0, trace=True, nn_mode=icpb200.NN_GRID, grid_cell=0.07)
    out = dc.download()
    ref, rout, rit, rdt = orc.icp(data, target, 4, 0.0, 0.75, 0, n_threads=8, trace=True)
    for k in range(5):
        acc = rdt[k] < 0.75
        assert 0 < acc.sum() < len(acc)
        assert np.array_equal(it[k][acc], rit[k][acc]) and np.array_equal(dt[k][acc], rdt[k][acc])
    assert np.array_equal(out.view(np.uint8), rout.view(np.uint8))
    assert np.array_equal(res["pose_R"], ref["pose_R"])
    dc.close(); tc.close()


def test_icp_ragged_batch_with_the_warp_filter(ctx, orc, pair10k):
    """Batches whose largest registration has >= 4096 queries take the warp-centred filter (batched Morton sort, one
    permutation per registration): ragged sizes, every registration checked against the oracle."""
    data, target = pair10k
    sizes = [(9000, 10000), (4096, 5000), (10000, 3000), (33, 7000), (5001, 4097)]
    datas, targets, want = [], [], []
    for k, (n, m) in enumerate(sizes):
        d = np.ascontiguousarray(data[k * 37: k * 37 + n]); t = np.ascontiguousarray(target[k * 11: k * 11 + m])
        datas.append(ctx.cloud_from_points(d)); targets.append(ctx.cloud_from_points(t))
        want.append(orc.icp(d, t, 4, 0.0, 0.75, orc.SOLVE_REFERENCE, n_threads=8))
    res = ctx.icp_register_batch(datas, targets, 4, 0.0, 0.75, 0)
    for k in range(len(sizes)):
        assert res[k]["nn_filter_used"] == 2          # ICPB_FILTER_WARP
        assert np.array_equal(res[k]["pose_R"], want[k][0]["pose_R"]) and np.array_equal(res[k]["pose_t"], want[k][0]["pose_t"])
        assert res[k]["n_assoc"] == want[k][0]["n_assoc"]
        assert np.array_equal(datas[k].download().view(np.uint8), want[k][1].view(np.uint8))
    for c in datas + targets:
        c.close()


@pytest.mark.parametrize("flt", ["1", "2", "3"])
def test_icp_randomized_differential(ctx, orc, monkeypatch, flt):
    """Random registration problems (sizes, overlap, noise, solve mode) under each filter: iterations, association
    count, composed pose and the transformed cloud are bit-equal to the oracle's."""
    from scipy.spatial.transform import Rotation
    monkeypatch.setenv("ICPB_NN_FILTER", flt)
    rng = np.random.default_rng(500 + int(flt))
    for case in range(8):
        n = int(rng.integers(3, 5000)); m = int(rng.integers(3, 6000))
        uv = rng.uniform(0, 2, (m, 2))
        t = np.stack([4 + uv[:, 0], 4 + uv[:, 1], 5 + 0.2 * np.sin(3 * uv[:, 0]) * np.cos(2 * uv[:, 1])], 1)
        sel = rng.integers(0, m, n)
        R = Rotation.from_euler("xyz", rng.uniform(-3, 3, 3), degrees=True).as_matrix()
        q = (t[sel] - 5) @ R.T + 5 + rng.uniform(-0.03, 0.03, 3) + rng.normal(0, 2e-3, (n, 3))
        far = rng.random(n) < 0.1
        q[far] += rng.uniform(1.0, 2.0, (int(far.sum()), 3))          # some queries beyond the acceptance radius
        data, target = orc.make_points(q), orc.make_points(t)
        mode = int(rng.integers(0, 2)); iters = int(rng.integers(1, 6)); thr = float(rng.choice([0.0, 1e-4]))
        dc, tc = ctx.cloud_from_points(data), ctx.cloud_from_points(target)
        res, _, _ = ctx.icp_register(dc, tc, iters, thr, 0.75, mode)
        want, wout, _, _ = orc.icp(data, target, iters, thr, 0.75, mode, n_threads=8)
        for k in ("iterations", "nn_passes", "n_assoc", "small_assoc_exit"):
            assert res[k] == want[k], (case, k)
        assert np.array_equal(res["pose_R"], want["pose_R"]) and np.array_equal(res["pose_t"], want["pose_t"]), case
        assert np.array_equal(dc.download().view(np.uint8), wout.view(np.uint8)), case
        dc.close(); tc.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_icp_batch_in_grid_mode_matches_the_oracle(ctx, orc, pair10k, mode):
    """ICPB_NN_GRID for a batch (round 2: the cells of all registrations are built together, one sorted array, one scan):
    ragged sizes, one registration with partial overlap, every pose / association count / transformed cloud bit-equal
    to the oracle's brute-force loop."""
    import icpb200
    data, target = pair10k
    rng = np.random.default_rng(7)
    cases = [(data[:6000], target[:7000]), (data[100:4196], target[:4096]), (data[::3], target[::2]),
             (data[:3000].copy(), target[5000:]), (data[:64], target[:33]), (data, target)]
    shifted = cases[3][0].copy()
    shifted["x"] += np.float32(0.4)                    # partial overlap: many queries without a neighbour inside 0.75 m
    cases[3] = (shifted, cases[3][1])
    datas = [ctx.cloud_from_points(d) for d, _ in cases]
    targets = [ctx.cloud_from_points(t) for _, t in cases]
    res = ctx.icp_register_batch(datas, targets, 8, 0.0, 0.75, mode, nn_mode=icpb200.NN_GRID)
    for k, (d, t) in enumerate(cases):
        assert res[k]["nn_mode_used"] == icpb200.NN_GRID
        ref, rout, _, _ = orc.icp(d, t, 8, 0.0, 0.75, mode, n_threads=8)
        assert res[k]["n_assoc"] == ref["n_assoc"], k
        assert np.array_equal(res[k]["pose_R"], ref["pose_R"]) and np.array_equal(res[k]["pose_t"], ref["pose_t"]), k
        assert np.array_equal(datas[k].download().view(np.uint8), rout.view(np.uint8)), k
    for c in datas + targets:
        c.close()


# ---- icpb_icp_register_async: the same loop, the host not blocked ---------------------------------------------------
def _same_result(a, b):
    for k in ("iterations", "nn_passes", "n_assoc", "mse"):
        assert a[k] == b[k], k
    for k in ("pose_R", "pose_t", "rigid", "cam_rotation", "cam_position"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("nn_mode", [0, 1])
def test_icp_async_equals_blocking_call_and_the_oracle(ctx, orc, pair10k, nn_mode):
    data, target = pair10k
    data, target = data[:6000], target[:7000]
    ref, rout, _, _ = orc.icp(data, target, 12, 0.0, 0.75, 0, n_threads=8)
    tc = ctx.cloud_from_points(target)
    d_sync, d_async = ctx.cloud_from_points(data), ctx.cloud_from_points(data)
    r_sync, _, _ = ctx.icp_register(d_sync, tc, 12, 0.0, 0.75, 0, nn_mode=nn_mode)
    pend = ctx.icp_register_async(d_async, tc, 12, 0.0, 0.75, 0, nn_mode=nn_mode)
    # work queued behind the registration on the same stream while it is in flight
    other = ctx.cloud_from_points(target[:100])
    other.transform(None, np.array([1, 2, 3], np.float32))
    r_async = pend.wait()
    assert pend.h is None
    _same_result(r_async, r_sync)
    assert np.array_equal(r_async["pose_R"], ref["pose_R"]) and np.array_equal(r_async["pose_t"], ref["pose_t"])
    assert np.array_equal(d_async.download().view(np.uint8), rout.view(np.uint8))
    assert np.array_equal(d_sync.download().view(np.uint8), rout.view(np.uint8))
    moved = other.download()
    assert np.array_equal(moved["x"], target[:100]["x"] + np.float32(1))
    for c in (tc, d_sync, d_async, other):
        c.close()


def test_icp_async_is_drained_by_the_next_registration(ctx, orc, pair10k):
    """One registration in flight per context: a later call completes the earlier one first; its results wait in the
    handle."""
    data, target = pair10k
    a = (data[:3000], target[:4000])
    b = (data[2000:7000], target[1000:8000])
    refs = [orc.icp(d, t, 8, 0.0, 0.75, 0, n_threads=8)[0] for d, t in (a, b)]
    ca, ta = ctx.cloud_from_points(a[0]), ctx.cloud_from_points(a[1])
    cb, tb = ctx.cloud_from_points(b[0]), ctx.cloud_from_points(b[1])
    pa = ctx.icp_register_async(ca, ta, 8, 0.0, 0.75, 0)
    rb, _, _ = ctx.icp_register(cb, tb, 8, 0.0, 0.75, 0)      # drains pa
    assert pa.ready()
    ra = pa.wait()
    for r, ref in ((ra, refs[0]), (rb, refs[1])):
        assert r["n_assoc"] == ref["n_assoc"]
        assert np.array_equal(r["pose_R"], ref["pose_R"]) and np.array_equal(r["pose_t"], ref["pose_t"])
    with pytest.raises(RuntimeError):
        pa.wait()
    for c in (ca, ta, cb, tb):
        c.close()


def test_icp_batch_async_equals_batch(ctx, pair10k):
    data, target = pair10k
    cases = [(data[:4000], target[:5000]), (data[3000:9000], target), (data[::2], target[::3])]
    def clouds():
        return [ctx.cloud_from_points(d) for d, _ in cases], [ctx.cloud_from_points(t) for _, t in cases]
    d1, t1 = clouds()
    d2, t2 = clouds()
    sync = ctx.icp_register_batch(d1, t1, 10, 0.0, 0.75, 0)
    pend = ctx.icp_register_batch_async(d2, t2, 10, 0.0, 0.75, 0)
    res = pend.wait()
    assert len(res) == len(cases)
    for k in range(len(cases)):
        _same_result(res[k], sync[k])
        assert np.array_equal(d1[k].download().view(np.uint8), d2[k].download().view(np.uint8))
    for c in d1 + t1 + d2 + t2:
        c.close()


@pytest.mark.parametrize("nn_mode", [0, 1])
def test_icp_register_carry_in_both_search_modes(ctx, orc, pair10k, nn_mode):
    """icpb_icp_register_carry (dataCloud.rotate / translate move points AND key-points, pointcloud.cpp:321-359): the
    carried cloud ends where the oracle puts it, with the brute-force scan and with the exact cell-grid search."""
    data, target = pair10k
    data, target = data[:6000], target[:7000]
    carry = data[::7].copy()
    carry["x"] += np.float32(0.01)
    ref, rdata, rcarry = orc.icp_carry(data, carry, target, 10, 0.0, 0.75, 0, n_threads=8)
    dc, tc, cc = ctx.cloud_from_points(data), ctx.cloud_from_points(target), ctx.cloud_from_points(carry)
    res = ctx.icp_register_carry(dc, tc, cc, 10, 0.0, 0.75, 0, nn_mode=nn_mode)
    assert res["nn_mode_used"] == nn_mode
    assert res["n_assoc"] == ref["n_assoc"]
    assert np.array_equal(res["pose_R"], ref["pose_R"]) and np.array_equal(res["pose_t"], ref["pose_t"])
    assert np.array_equal(dc.download().view(np.uint8), rdata.view(np.uint8))
    assert np.array_equal(cc.download().view(np.uint8), rcarry.view(np.uint8))
    for c in (dc, tc, cc):
        c.close()
