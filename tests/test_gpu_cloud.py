"""GPU parity: back-projection (P1), transform (P2), normals (P3), depth filter (8f-1) vs the CPU oracle.
Bar: bit-exact points, order and count."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frame(seed=0, sensor=None):
    from icpb200 import synth
    sensor = sensor or synth.KINECT_V1
    poses = synth.trajectory(3, seed=synth.MASTER_SEED + seed)
    R, t = poses[2]
    return synth.render_depth(R, t, sensor, seed=synth.MASTER_SEED + seed), synth.render_color(sensor, seed)


def _same_points(a, b):
    assert len(a) == len(b), (len(a), len(b))
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.mark.parametrize("rule,arg", [(0, 1), (1, 40), (1, 7), (2, 40), (2, 3)])
def test_backproject_rules_v1(ctx, orc, rule, arg):
    import icpb200
    depth, bgr = _frame(1)
    c = ctx.cloud(depth.size)
    n = c.from_depth(depth, bgr, icpb200.reference_intrinsics_v1(), rule, arg, seed=1234)
    ref, _, _ = orc.backproject(depth, bgr, orc.kinect_v1(), rule, arg, seed=1234)
    assert n == len(ref)
    _same_points(c.download(), ref)
    c.close()


@pytest.mark.parametrize("sub", ["2", "4"])
@pytest.mark.parametrize("rule,arg", [(0, 1), (1, 7), (2, 3)])
def test_backproject_tiles_per_cta(ctx, orc, monkeypatch, sub, rule, arg):
    """The chained kernel walks 2 or 4 tiles per CTA through one staging buffer (4 for launches with thousands of
    tiles): both run lengths, every subsample rule, on a frame and on sizes that end inside a run."""
    import icpb200
    monkeypatch.setenv("ICPB_BP_SUB", sub)
    depth, bgr = _frame(4)
    for rows in (480, 13, 7, 3):      # 150 tiles, and frames whose last run is partial or a single short tile
        d = np.ascontiguousarray(depth[:rows]); b = np.ascontiguousarray(bgr[:rows])
        c = ctx.cloud(d.size)
        n = c.from_depth(d, b, icpb200.reference_intrinsics_v1(), rule, arg, seed=99)
        ref, _, _ = orc.backproject(d, b, orc.kinect_v1(), rule, arg, seed=99)
        assert n == len(ref)
        _same_points(c.download(), ref)
        c.close()


def test_backproject_v2_no_color(ctx, orc):
    import icpb200
    from icpb200 import synth
    depth, _ = _frame(2, synth.KINECT_V2)
    c = ctx.cloud(depth.size)
    c.from_depth(depth, None, icpb200.reference_intrinsics_v2())
    ref, _, _ = orc.backproject(depth, None, orc.kinect_v2())
    _same_points(c.download(), ref)
    c.close()


def test_backproject_rand_stream(ctx, orc):
    """ICPB_SUB_STREAM replays `rand() % 40` (pointcloud.cpp:28): one decision per non-zero pixel."""
    import icpb200
    depth, bgr = _frame(3)
    rng = np.random.default_rng(7)
    stream = (rng.integers(0, 40, int((depth > 0).sum())) == 0).astype(np.uint8)
    c = ctx.cloud(depth.size)
    c.from_depth(depth, bgr, None, icpb200.SUB_STREAM, 40, 0, stream)
    ref, _, _ = orc.backproject(depth, bgr, None, orc.SUB_STREAM, 40, 0, stream)
    _same_points(c.download(), ref)
    assert 0.015 * depth.size < len(ref) < 0.035 * depth.size
    c.close()


def test_backproject_every_depth_value(ctx, orc):
    """T1: every uint16 depth value, on a sweep of pixel coordinates."""
    import icpb200
    w, h = 640, 480
    depth = (np.arange(w * h, dtype=np.uint32) * 7919 % 65536).astype(np.uint16).reshape(h, w)
    depth.ravel()[:65536] = np.arange(65536, dtype=np.uint16)
    c = ctx.cloud(depth.size)
    c.from_depth(depth, None, None)
    ref, _, _ = orc.backproject(depth, None, None)
    _same_points(c.download(), ref)
    c.close()


@pytest.mark.parametrize("w,h", [(8, 1), (16, 3), (33, 5), (2048, 1), (2056, 2), (100, 77)])
def test_backproject_ragged_and_empty(ctx, orc, w, h):
    rng = np.random.default_rng(w * h)
    depth = rng.integers(0, 3, (h, w)).astype(np.uint16) * rng.integers(1000, 20000, (h, w)).astype(np.uint16)
    K = orc.kinect_v1()
    import icpb200
    c = ctx.cloud(max(depth.size, 1))
    if (w * h) % 8 != 0:
        pass  # tail path
    n = c.from_depth(depth, None, None)
    ref, _, _ = orc.backproject(depth, None, K)
    assert n == len(ref)
    _same_points(c.download(), ref)
    zero = np.zeros((h, w), np.uint16)
    assert c.from_depth(zero, None, None) == 0
    c.close()


def test_backproject_capacity_error(ctx, orc):
    import icpb200
    depth, _ = _frame(1)
    c = ctx.cloud(1000)
    with pytest.raises(icpb200.IcpbError) as e:
        c.from_depth(depth, None, None)
    assert e.value.status == icpb200.ERR_CAPACITY
    c.close()


def test_transform_matches_oracle(ctx, orc, pair10k):
    from icpb200 import synth
    data, _ = pair10k
    R = synth.rot_axis_angle([1, 2, 3], 0.1).astype(np.float32)
    t = np.array([5, 5, 5], np.float32)
    c = ctx.cloud_from_points(data)
    c.transform(R, t)
    ref = orc.translate(orc.rotate(data, R), t)
    _same_points(c.download(), ref)
    c.transform(None, -t)
    _same_points(c.download(), orc.translate(ref, -t))
    c.close()


def test_center_canonical(ctx, orc, pair10k):
    data, _ = pair10k
    c = ctx.cloud_from_points(data)
    got = c.center()
    terms = orc.xyz_of(data).astype(np.float64)
    want = orc.canon_reduce(terms) / len(data)
    assert np.array_equal(got, want)
    # the reference's float running mean agrees to float accuracy (pointcloud.cpp:43-45,100-102)
    assert np.allclose(got, terms.mean(0), rtol=0, atol=1e-9)
    c.close()


def test_normals(ctx, orc):
    depth, _ = _frame(4)
    got = ctx.normals(depth)
    want = orc.normals(depth)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("w,h", [(640, 480), (512, 424), (8, 3), (4, 1), (12, 2), (33, 5), (7, 7)])
def test_normals_sizes(ctx, orc, w, h):
    """Widths that are multiples of 4 take the four-pixels-per-thread kernel, the others the scalar one; borders are
    zeros either way (SLAM.cpp:412-430 never writes row / column 0 and reads past the last ones)."""
    rng = np.random.default_rng(w * 100 + h)
    depth = rng.integers(0, 65536, (h, w)).astype(np.uint16)
    got = ctx.normals(depth)
    want = orc.normals(depth)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_normals_batch_device(ctx, orc):
    import torch
    rng = np.random.default_rng(12)
    depth = rng.integers(0, 65536, (5, 424, 512)).astype(np.uint16)
    dev = torch.device("cuda", ctx.device)
    d_in = torch.from_numpy(depth.view(np.int16)).to(dev)
    d_out = torch.empty((5, 424, 512, 3), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    ctx.normals_batch_device(d_in.data_ptr(), 5, 512, 424, d_out.data_ptr())
    ctx.sync()
    got = d_out.cpu().numpy()
    for f in range(5):
        assert np.array_equal(got[f].view(np.uint32), orc.normals(depth[f]).view(np.uint32))


def test_depth_filter(ctx, orc):
    from icpb200 import synth
    R, t = synth.trajectory(1)[0]
    raw = synth.render_depth(R, t, synth.KINECT_V2, dropout=0.1)
    raw[10:20, 30:60] = 40000   # beyond MAX_16_CHANNEL_DISTANCE
    raw[100:110, 5:9] = 500     # closer than MIN_16_CHANNEL_DISTANCE
    got = ctx.depth_filter(raw, 1000, 25000)
    want = orc.depth_filter(raw, 1000, 25000)
    assert np.array_equal(got, want)


def test_backproject_batch_device(ctx, orc):
    """Batched, sync-free back-projection: frame f of the batch == the single-frame constructor == the oracle."""
    import torch
    import icpb200
    from icpb200 import synth
    frames = [_frame(s, synth.KINECT_V2)[0] for s in (1, 2, 3)] * 2
    w, h = synth.KINECT_V2["w"], synth.KINECT_V2["h"]
    dev = torch.device("cuda", 0)
    d_depth = torch.from_numpy(np.stack(frames).astype(np.int16)).to(dev)    # u16 payload, viewed as i16 by torch
    cap = w * h
    d_pts = torch.zeros((len(frames), cap, 4), dtype=torch.float32, device=dev)
    d_cnt = torch.zeros(len(frames), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.backproject_batch_device(d_depth.data_ptr(), None, len(frames), w, h, icpb200.reference_intrinsics_v2(),
                                 d_pts.data_ptr(), cap, d_cnt.data_ptr())
    ctx.sync()
    cnt = d_cnt.cpu().numpy()
    pts = d_pts.cpu().numpy()
    for f, depth in enumerate(frames):
        ref, _, _ = orc.backproject(depth, None, orc.kinect_v2())
        assert cnt[f] == len(ref)
        got = np.ascontiguousarray(pts[f, : cnt[f]]).view(orc.POINT_DTYPE).reshape(-1)
        assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))


def _exhaustive_frames(w, h, along_rows):
    """Frames whose depth is constant along a row (or a column) and runs through 1..65535 over the frames: together
    with the other axis this pairs EVERY depth value with every image coordinate of that axis."""
    n = h if along_rows else w
    frames = []
    d = 1
    while d <= 65535:
        vals = np.clip(np.arange(d, d + n), 0, 65535).astype(np.uint16)
        vals[np.arange(d, d + n) > 65535] = 0
        frames.append(np.repeat(vals[:, None], w, 1) if along_rows else np.repeat(vals[None, :], h, 0))
        d += n
    return frames


@pytest.mark.parametrize("sensor", ["v1", "v2", "distinct_axes"])
def test_backproject_exhaustive_depth_x_coordinate(ctx, orc, sensor):
    """The kernel divides by the (host-known) intrinsics with a reciprocal + FMA-residual sequence instead of `/`.
    Proof by exhaustion that it is the correctly rounded quotient of pointcloud.cpp:37-39: every depth value 1..65535
    paired with every column (x) and with every row (y), bit-equal to the oracle's true divisions."""
    import icpb200
    if sensor == "v1":
        K, Ko, (w, h) = icpb200.reference_intrinsics_v1(), orc.kinect_v1(), (640, 480)
    elif sensor == "v2":
        K, Ko, (w, h) = icpb200.reference_intrinsics_v2(), orc.kinect_v2(), (512, 424)
    else:  # true per-axis intrinsics (pointcloud.hpp:7-10 has FY / CY; the reference never uses them)
        K, Ko, (w, h) = icpb200.Intrinsics(468.60, 318.27, 468.61, 243.99, 5000.0), \
            orc.Intrinsics(468.60, 318.27, 468.61, 243.99, 5000.0), (640, 480)
    c = ctx.cloud(w * h)
    for along_rows in (True, False):
        for depth in _exhaustive_frames(w, h, along_rows):
            c.from_depth(depth, None, K)
            ref, _, _ = orc.backproject(depth, None, Ko)
            _same_points(c.download(), ref)
    c.close()


def test_backproject_unfavourable_divisors_use_true_division(ctx, orc):
    """Divisors outside the proven range of the FMA sequence (all-ones significand, extreme exponents) take the plain
    IEEE division path: still bit-equal."""
    import icpb200
    depth, bgr = _frame(4)
    ones = float(np.uint32(0x43FFFFFF).view(np.float32))   # 511.99997, significand all ones
    for fx, scale in ((ones, 5000.0), (468.6, ones), (1.0e-25, 5000.0)):
        K, Ko = icpb200.Intrinsics(fx, 318.27, fx, 318.27, scale), orc.Intrinsics(fx, 318.27, fx, 318.27, scale)
        c = ctx.cloud(depth.size)
        c.from_depth(depth, bgr, K)
        ref, _, _ = orc.backproject(depth, bgr, Ko)
        _same_points(c.download(), ref)
        c.close()
