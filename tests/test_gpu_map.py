"""GPU parity: certainty grid (M1-M4) vs the CPU oracle.  Bar: bit-exact grid."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REF_DIMS, REF_CELL = (300, 300, 300), np.float32(10.0) / np.float32(300.0)   # map.hpp:9-10,17
README_DIMS, README_CELL = (300, 300, 250), 0.02                               # README.md:8-12


def _world_cloud(orc, seed=0, sensor=None, stride=1):
    from icpb200 import synth
    sensor = sensor or synth.KINECT_V2
    R, t = synth.trajectory(4, seed=synth.MASTER_SEED + seed)[3]
    depth = synth.render_depth(R, t, sensor, seed=seed)
    pts, _, _ = orc.backproject(depth, None, orc.kinect_v2() if sensor is synth.KINECT_V2 else orc.kinect_v1())
    pts = orc.translate(orc.rotate(pts, R.astype(np.float32)), t.astype(np.float32))
    return pts[::stride], t.astype(np.float32)


@pytest.mark.parametrize("rule,delta", [(0, 25), (0, 180), (1, 25), (1, 60)])
@pytest.mark.parametrize("dims,cell", [(REF_DIMS, REF_CELL), (README_DIMS, README_CELL)])
def test_endpoint_rules(ctx, orc, rule, delta, dims, cell):
    pts, _ = _world_cloud(orc, 1)
    m = ctx.map(dims, float(cell))
    c = ctx.cloud_from_points(pts)
    grid = np.zeros(dims, np.uint8)
    for _ in range(3):
        m.update_endpoints(c, rule, delta, 180)
        orc.map_update_endpoints(grid, dims, float(cell), pts, rule, delta, 180)
    got = m.download()
    assert hashlib.sha256(got.tobytes()).hexdigest() == hashlib.sha256(grid.tobytes()).hexdigest()
    assert got.max() == 255 and 0 < (got > 0).sum()
    m.close(); c.close()


def test_endpoint_multiplicities(ctx, orc):
    """T4: 0..20 hits per voxel at both deltas; result must be f^k(c) for every voxel."""
    dims, cell = (32, 32, 32), 0.1
    rng = np.random.default_rng(0)
    for rule, delta in [(0, 25), (0, 180), (1, 25), (1, 180)]:
        vox = rng.integers(0, 32, (400, 3))
        mult = rng.integers(0, 21, 400)
        xyz = np.repeat((vox + 0.5) * cell, mult, axis=0).astype(np.float32)
        rng.shuffle(xyz)
        pts = orc.make_points(xyz)
        start = rng.integers(0, 256, dims).astype(np.uint8)
        m = ctx.map(dims, cell)
        m.upload(start)
        c = ctx.cloud_from_points(pts)
        m.update_endpoints(c, rule, delta, 180)
        want = start.copy()
        orc.map_update_endpoints(want, dims, cell, pts, rule, delta, 180)
        assert np.array_equal(m.download(), want)
        m.close(); c.close()


def test_out_of_range_points_are_clamped_in(ctx, orc):
    dims, cell = (20, 20, 20), 0.1
    xyz = np.array([[-5, 1, 1], [1, -0.05, 1], [1, 1, 99], [2.0, 2.0, 2.0], [1.95, 1.999999, 0.0]], np.float32)
    pts = orc.make_points(xyz)
    m = ctx.map(dims, cell)
    c = ctx.cloud_from_points(pts)
    m.update_endpoints(c, 0, 25, 180)
    want = np.zeros(dims, np.uint8)
    orc.map_update_endpoints(want, dims, cell, pts, 0, 25, 180)
    assert np.array_equal(m.download(), want)
    for p in xyz:
        assert m.voxel_coords(p) == orc.voxel_coords(p, cell, dims)
    m.close(); c.close()


@pytest.mark.parametrize("dims,cell", [(README_DIMS, README_CELL), ((150, 150, 125), 0.04)])
def test_ray_integration_frames(ctx, orc, dims, cell):
    m = ctx.map(dims, cell)
    grid = np.zeros(dims, np.uint8)
    for f in range(3):
        pts, origin = _world_cloud(orc, f, stride=3)
        c = ctx.cloud_from_points(pts)
        v = m.integrate_rays(c, origin, 25, 25)
        rv = orc.map_integrate_rays(grid, dims, cell, pts, origin, 25, 25)
        assert v == rv and v > 0
        c.close()
    got = m.download()
    assert np.array_equal(got, grid), f"{(got != grid).sum()} voxels differ"
    m.close()


def test_ray_decrement_clamps_at_zero(ctx, orc):
    dims, cell = (64, 64, 64), 0.05
    rng = np.random.default_rng(2)
    start = rng.integers(0, 60, dims).astype(np.uint8)
    xyz = rng.uniform(0.1, 3.1, (5000, 3)).astype(np.float32)
    pts = orc.make_points(xyz)
    origin = (1.6, 1.6, 1.6)
    m = ctx.map(dims, cell); m.upload(start)
    c = ctx.cloud_from_points(pts)
    m.integrate_rays(c, origin, 25, 25)
    want = start.copy()
    orc.map_integrate_rays(want, dims, cell, pts, origin, 25, 25)
    assert np.array_equal(m.download(), want)
    m.close(); c.close()


def test_z_slabs_concatenate_to_the_full_grid(ctx, orc):
    """Config 5 shape on one GPU: slab owners see every ray, write only their z-range; the
    concatenation equals the un-sharded grid byte for byte."""
    dims, cell = (120, 120, 100), 0.05
    pts, origin = _world_cloud(orc, 5, stride=2)
    c = ctx.cloud_from_points(pts)
    full = ctx.map(dims, cell)
    full.integrate_rays(c, origin, 25, 25)
    want = full.download()
    parts = []
    for g in range(4):
        s = ctx.map(dims, cell, g * 25, (g + 1) * 25)
        s.integrate_rays(c, origin, 25, 25)
        parts.append(s.download())
        s.close()
    assert np.array_equal(np.concatenate(parts, axis=2), want)
    ref = np.zeros(dims, np.uint8)
    orc.map_integrate_rays(ref, dims, cell, pts, origin, 25, 25)
    assert np.array_equal(want, ref)
    full.close(); c.close()


@pytest.mark.parametrize("origin_z", [0.3, 1.7, 3.1])
def test_slab_clipped_walks_random_rays(ctx, orc, origin_z):
    """Rays in every direction, origin below / inside / above the slabs: each slab handle jumps straight to its
    part of every walk (closed-form entry step) and must still reproduce the un-sharded grid."""
    dims, cell = (64, 48, 72), 0.05
    rng = np.random.default_rng(int(origin_z * 10))
    xyz = rng.uniform([0, 0, 0], [3.2, 2.4, 3.6], (20000, 3)).astype(np.float32)
    pts = orc.make_points(xyz)
    origin = (1.31, 1.07, origin_z)
    start = rng.integers(0, 90, dims).astype(np.uint8)
    want = start.copy()
    orc.map_integrate_rays(want, dims, cell, pts, origin, 25, 25)
    c = ctx.cloud_from_points(pts)
    bounds = [0, 7, 8, 30, 55, 72]
    parts = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        s = ctx.map(dims, cell, lo, hi)
        s.upload(np.ascontiguousarray(start[:, :, lo:hi]))
        s.integrate_rays(c, origin, 25, 25)
        parts.append(s.download())
        s.close()
    got = np.concatenate(parts, axis=2)
    assert np.array_equal(got, want), f"{(got != want).sum()} voxels differ"
    c.close()


@pytest.mark.parametrize("variant,delta", [(0, 180), (0, 25), (1, 25), (2, 25), (2, 60)])
def test_tracked_updates_match_oracle(ctx, orc, pair10k, variant, delta):
    """M3 with the lookup-table / mapCloud bookkeeping on the device: grid, inserted points and their order."""
    data, target = pair10k
    dims, cell = REF_DIMS, float(REF_CELL)
    m = ctx.map(dims, cell)
    mc = ctx.cloud(200000)
    grid = np.zeros(dims, np.uint8)
    table = np.full(dims, -1, np.int32)
    mine = []
    for rep in range(12):
        pts = np.ascontiguousarray((data if rep % 2 == 0 else target)[rep * 37: rep * 37 + 2500])
        c = ctx.cloud_from_points(pts)
        k = m.update_tracked(c, variant, delta, 180, mc)
        app = orc.map_update_tracked(grid, table, dims, cell, pts, variant, delta, 180, len(mine))
        assert k == len(app)
        mine.extend(pts[app])
        c.close()
    assert np.array_equal(m.download(), grid)
    got = mc.download()
    assert len(got) == len(mine) and len(mine) > 0
    assert np.array_equal(got.view(np.uint8), np.array(mine, dtype=orc.POINT_DTYPE).view(np.uint8))
    probe = orc.xyz_of(np.array(mine[:1], dtype=orc.POINT_DTYPE))[0]
    assert m.has_entry(tuple(float(x) for x in probe))
    m.close(); mc.close()


def test_sync_free_frame_path_equals_the_synchronising_calls(ctx, orc):
    """icpb_frame_lift_band_device + icpb_map_integrate_bands_device (point count kept on the device) against
    from_depth / transform / integrate_rays: the same grid, byte for byte -- as one band, and as two row bands feeding
    two z-slabs (the world-size-2 layout emulated on one GPU: bands back to back, as an all-gather leaves them)."""
    import torch
    import icpb200
    from icpb200 import synth
    from icpb200 import dist as D
    K = icpb200.reference_intrinsics_v1()
    dims, cell = (300, 300, 250), 0.02
    poses = synth.trajectory(3, step_deg=1.0, step_m=0.03)
    depths = [synth.render_depth(R, t, synth.KINECT_V1, seed=f) for f, (R, t) in enumerate(poses)]
    h, w = depths[0].shape
    dev = torch.device("cuda", ctx.device)
    d_depths = torch.from_numpy(np.stack(depths).view(np.int16)).to(dev)
    torch.cuda.synchronize()   # the fixture's context runs on its own stream

    ref_map = ctx.map(dims, cell)
    cl = ctx.cloud(w * h)
    for (R, t), d in zip(poses, depths):
        cl.from_depth(d, None, K); cl.transform(R.astype(np.float32), t.astype(np.float32))
        ref_map.integrate_rays(cl, tuple(float(x) for x in t), 25, 25, False)
    want = ref_map.download()
    assert (want > 0).sum() > 1000

    # one band
    m1 = ctx.map(dims, cell)
    cap = w * h
    band = torch.zeros((cap + 1, 4), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    for f, (R, t) in enumerate(poses):
        ctx.frame_lift_band_device(d_depths.data_ptr() + f * w * h * 2, w, h, 0, h, K, R, t, band.data_ptr(), cap)
        m1.integrate_bands_device(band.data_ptr(), 1, cap, tuple(float(x) for x in t))
    assert np.array_equal(m1.download(), want)

    # two row bands, two z-slabs
    world = 2
    bcap = (-(-h // world)) * w
    bands = torch.zeros((world * (bcap + 1), 4), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    slabs = [ctx.map(dims, cell, *D.slab_bounds(dims[2], g, world)) for g in range(world)]
    for f, (R, t) in enumerate(poses):
        for g in range(world):
            r0, r1 = D.row_band(h, g, world)
            ctx.frame_lift_band_device(d_depths.data_ptr() + f * w * h * 2, w, h, r0, r1, K, R, t,
                                       bands.data_ptr() + g * (bcap + 1) * 16, bcap)
        for g in range(world):
            slabs[g].integrate_bands_device(bands.data_ptr(), world, bcap, tuple(float(x) for x in t))
    got = np.concatenate([s_.download() for s_ in slabs], axis=2)
    assert np.array_equal(got, want)
    for x in (ref_map, m1, cl, *slabs):
        x.close()


def test_frame_band_argument_checks(ctx):
    """The sync-free calls cannot report a capacity problem after the fact (the count never reaches the host), so
    everything that can be checked is checked before anything is enqueued."""
    import torch
    import icpb200
    K = icpb200.reference_intrinsics_v1()
    dev = torch.device("cuda", ctx.device)
    depth = torch.zeros((480, 640), dtype=torch.int16, device=dev)
    band = torch.zeros((640 * 480 + 1, 4), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    with pytest.raises(icpb200.IcpbError) as e:      # band smaller than its pixel count
        ctx.frame_lift_band_device(depth.data_ptr(), 640, 480, 0, 480, K, None, None, band.data_ptr(), 640 * 240)
    assert e.value.status == icpb200.ERR_CAPACITY
    for r0, r1 in ((-1, 10), (10, 10), (0, 481)):   # rows outside the image / empty band
        with pytest.raises(icpb200.IcpbError) as e:
            ctx.frame_lift_band_device(depth.data_ptr(), 640, 480, r0, r1, K, None, None, band.data_ptr(), 640 * 480)
        assert e.value.status == icpb200.ERR_INVALID
    # an all-zero band: count 0 in the header, integrating it changes nothing
    ctx.frame_lift_band_device(depth.data_ptr(), 640, 480, 0, 480, K, None, None, band.data_ptr(), 640 * 480)
    m = ctx.map((300, 300, 250), 0.02)
    m.integrate_bands_device(band.data_ptr(), 1, 640 * 480, (3.0, 3.0, 1.0))
    ctx.sync()
    assert int(band[0, 0].view(torch.int32).item()) == 0
    assert int(m.download().sum()) == 0
    m.close()


# ---- brick occupancy (round 2): the walk crosses bricks whose bit is clear in one jump -----------------------------

def _sparse_start(rng, dims, p_brick, brick=8):
    """A grid that is zero except inside a random subset of 8^3 bricks (there: random bytes, many of them zero)."""
    nb = [-(-d // brick) for d in dims]
    occ = rng.random(nb) < p_brick
    mask = np.repeat(np.repeat(np.repeat(occ, brick, 0), brick, 1), brick, 2)[: dims[0], : dims[1], : dims[2]]
    vals = rng.integers(0, 120, dims).astype(np.uint8) * (rng.random(dims) < 0.4)
    return (vals * mask).astype(np.uint8)


@pytest.mark.parametrize("p_brick", [0.0, 0.05, 0.3, 1.0])
@pytest.mark.parametrize("dims,bounds", [((64, 48, 72), [0, 72]), ((61, 43, 70), [0, 7, 8, 30, 55, 70]),
                                         ((9, 200, 17), [0, 3, 17])])
def test_brick_skipping_walk_sparse_grids(ctx, orc, p_brick, dims, bounds):
    """Random rays through grids whose occupied voxels sit in a random subset of bricks, whole map and ragged z-slabs
    (slab heights that are not multiples of the brick, dims that are not either): every slab byte-equal to the oracle,
    over two frames so that the second walk crosses the bricks the first frame's endpoints occupied."""
    cell = 0.05
    rng = np.random.default_rng(int(p_brick * 100) + dims[0])
    start = _sparse_start(rng, dims, p_brick)
    want = start.copy()
    hi = np.array(dims) * cell
    frames = []
    for f in range(2):
        xyz = rng.uniform([0, 0, 0], hi, (12000, 3)).astype(np.float32)
        origin = tuple(float(v) for v in rng.uniform([0, 0, 0], hi))
        frames.append((orc.make_points(xyz), origin))
        orc.map_integrate_rays(want, dims, cell, frames[-1][0], origin, 25, 25)
    parts = []
    for lo, hi_z in zip(bounds[:-1], bounds[1:]):
        s = ctx.map(dims, cell, lo, hi_z)
        s.upload(np.ascontiguousarray(start[:, :, lo:hi_z]))
        for pts, origin in frames:
            c = ctx.cloud_from_points(pts)
            s.integrate_rays(c, origin, 25, 25)
            c.close()
        parts.append(s.download())
        s.close()
    got = np.concatenate(parts, axis=2)
    assert np.array_equal(got, want), f"{(got != want).sum()} voxels differ"


def test_brick_bits_follow_clear_upload_and_endpoint_updates(ctx, orc):
    """The occupancy bits must never claim "empty" for a brick that holds a non-zero voxel, whichever call wrote it:
    endpoint updates (rule A and C), tracked updates, upload; clear resets them.  Checked through the walk's result."""
    import icpb200
    dims, cell = (48, 48, 48), 0.05
    rng = np.random.default_rng(5)
    hi = np.array(dims) * cell
    m = ctx.map(dims, cell)
    want = np.zeros(dims, np.uint8)

    def rays(seed):
        r = np.random.default_rng(seed)
        pts = orc.make_points(r.uniform([0, 0, 0], hi, (6000, 3)).astype(np.float32))
        origin = tuple(float(v) for v in r.uniform([0, 0, 0], hi))
        c = ctx.cloud_from_points(pts)
        m.integrate_rays(c, origin, 40, 25)
        orc.map_integrate_rays(want, dims, cell, pts, origin, 40, 25)
        c.close()
        assert np.array_equal(m.download(), want)

    for rule in (0, 1):
        pts = orc.make_points(rng.uniform([0, 0, 0], hi, (3000, 3)).astype(np.float32))
        c = ctx.cloud_from_points(pts)
        m.update_endpoints(c, rule, 60, 180)
        orc.map_update_endpoints(want, dims, cell, pts, rule, 60, 180)
        c.close()
        rays(10 + rule)
    kp = orc.make_points(rng.uniform([0, 0, 0], hi, (2000, 3)).astype(np.float32))
    c = ctx.cloud_from_points(kp)
    mc = ctx.cloud(4096)
    m.update_tracked(c, icpb200.TRACK_INIT, 180, 180, mc)
    orc.map_update_endpoints(want, dims, cell, kp, 0, 180, 180)
    c.close(); mc.close()
    rays(20)
    m.clear(); want[:] = 0
    rays(30)
    start = _sparse_start(rng, dims, 0.2)
    m.upload(start); want[:] = start
    rays(40)
    m.close()


@pytest.mark.parametrize("fpe", [1, 3, 8])
def test_slabmap_sequence_single_rank_matches_oracle(ctx, orc, fpe):
    """icpb_slabmap (the C-ABI z-slab map) with one rank: a sequence of device-resident frames, k frames per lift /
    walk group on three streams, must give the oracle's grid -- the multi-rank path differs only by the all-gather."""
    import ctypes
    import icpb200
    from icpb200 import synth
    import torch
    frames, dims, cell = 7, (300, 300, 250), 0.02
    poses = synth.trajectory(frames, step_deg=0.8, step_m=0.02)
    depths = [synth.render_depth(R, t, synth.KINECT_V2, seed=f) for f, (R, t) in enumerate(poses)]
    K = icpb200.reference_intrinsics_v2()
    want = np.zeros(dims, np.uint8)
    for (R, t), dpt in zip(poses, depths):
        pts, _, _ = orc.backproject(dpt, None, orc.kinect_v2())
        pts = orc.translate(orc.rotate(pts, np.asarray(R, np.float32)), np.asarray(t, np.float32))
        orc.map_integrate_rays(want, dims, cell, pts, tuple(float(x) for x in t), 25, 25)
    h, w = depths[0].shape
    sm = icpb200.SlabMapC(ctx, None, dims, cell, w, h)
    d_depths = torch.from_numpy(np.stack(depths).astype(np.uint16).view(np.int16)).cuda()
    torch.cuda.synchronize()
    Rs = np.stack([np.asarray(R, np.float32) for R, _ in poses])
    ts = np.stack([np.asarray(t, np.float32) for _, t in poses])
    sm.integrate_sequence_device(d_depths.data_ptr(), frames, K, Rs, ts, 25, 25, fpe)
    ctx.sync()
    got = sm.download()
    assert np.array_equal(got, want), f"{(got != want).sum()} voxels differ"
    # a second pass over the same frames (occupied bricks everywhere the first pass put endpoints)
    sm.integrate_sequence_device(d_depths.data_ptr(), frames, K, Rs, ts, 25, 25, fpe)
    ctx.sync()
    for (R, t), dpt in zip(poses, depths):
        pts, _, _ = orc.backproject(dpt, None, orc.kinect_v2())
        pts = orc.translate(orc.rotate(pts, np.asarray(R, np.float32)), np.asarray(t, np.float32))
        orc.map_integrate_rays(want, dims, cell, pts, tuple(float(x) for x in t), 25, 25)
    assert np.array_equal(sm.download(), want)
    sm.close()


def test_profiled_integration_same_grid_and_a_work_histogram(ctx, orc):
    """icpb_map_integrate_rays_profiled: the calibration call behind the z-slab boundaries changes nothing in the grid."""
    import icpb200
    dims, cell = (150, 150, 125), 0.04
    m = ctx.map(dims, cell)
    grid = np.zeros(dims, np.uint8)
    work = np.zeros(dims[2], np.uint64)
    for f in range(2):
        pts, origin = _world_cloud(orc, f, stride=3)
        c = ctx.cloud_from_points(pts)
        m.integrate_rays_profiled(c, origin, work, 25, 25)
        v = orc.map_integrate_rays(grid, dims, cell, pts, origin, 25, 25)
        c.close()
        assert 0 < int(work.sum()) <= 6 * len(pts) * (f + 1) + 4 * 2 * v * (f + 1)
    assert np.array_equal(m.download(), grid)
    b = icpb200.slab_bounds_from_work(work, 4)
    shares = [int(work[b[g]:b[g + 1]].sum()) for g in range(4)]
    assert max(shares) <= work.sum() / 4 + work.max() + 1
    m.close()
