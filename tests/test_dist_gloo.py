"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group, the oracle standing in for the kernels.
Checks that row-band lifting + all-gather reproduces the single-process point list and that z-slab
ownership reproduces the single-process grid byte for byte (T7), and the batch sharding bookkeeping."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "icp-slam-prototype_b200", "python"))
    import torch
    import torch.distributed as dist
    from icpb200 import dist as D
    from icpb200 import synth
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dims, cell = (60, 60, 50), 0.1
        z_lo, z_hi = D.slab_bounds(dims[2], rank, world)
        grid = np.zeros(dims, np.uint8)
        poses = synth.trajectory(3)
        for f, (R, t) in enumerate(poses):
            depth = synth.render_depth(R, t, synth.KINECT_V2, seed=f)[::4, ::4].copy()
            h, w = depth.shape
            r0, r1 = D.row_band(h, rank, world)
            local, _, _ = orc.backproject(D.mask_rows(depth, r0, r1), None, orc.kinect_v2())
            local = orc.translate(orc.rotate(local, R.astype(np.float32)), t.astype(np.float32))
            lt = torch.from_numpy(local.view(np.float32).reshape(-1, 4).copy())
            allp, counts = D.all_gather_points(lt)
            pts = allp.numpy().copy().view(orc.POINT_DTYPE).reshape(-1)
            orc.map_integrate_rays(grid, dims, cell, pts, tuple(float(x) for x in t), 25, 25, z_lo, z_hi)
            if f == 0:
                np.save(os.path.join(out_dir, f"pts_rank{rank}.npy"), pts)
        np.save(os.path.join(out_dir, f"slab_rank{rank}.npy"), grid[:, :, z_lo:z_hi])
        # batch sharding: each rank "registers" its block, poses are gathered in batch order
        lo, hi = D.shard_range(11, rank, world)
        rows = torch.tensor([[float(i), float(i) * 2] for i in range(lo, hi)], dtype=torch.float64).reshape(-1, 2)
        allrows = D.gather_results(rows)
        np.save(os.path.join(out_dir, f"rows_rank{rank}.npy"), allrows.numpy())
    finally:
        dist.destroy_process_group()


def test_shard_helpers():
    sys.path.insert(0, os.path.join(ROOT, "icp-slam-prototype_b200", "python"))
    from icpb200 import dist as D
    for n in [0, 1, 7, 8, 1024, 1025]:
        for world in [1, 2, 3, 4, 8]:
            blocks = [D.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert D.slab_bounds(500, 3, 8) == (189, 252) and D.slab_bounds(500, 7, 8) == (438, 500)


@pytest.mark.timeout(300)
def test_two_rank_slabs_and_gather(tmp_path, orc):
    import torch.multiprocessing as mp
    from icpb200 import dist as D
    from icpb200 import synth
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    # single-process truth
    dims, cell = (60, 60, 50), 0.1
    grid = np.zeros(dims, np.uint8)
    first_pts = None
    for f, (R, t) in enumerate(synth.trajectory(3)):
        depth = synth.render_depth(R, t, synth.KINECT_V2, seed=f)[::4, ::4].copy()
        pts, _, _ = orc.backproject(depth, None, orc.kinect_v2())
        pts = orc.translate(orc.rotate(pts, R.astype(np.float32)), t.astype(np.float32))
        if f == 0:
            first_pts = pts
        orc.map_integrate_rays(grid, dims, cell, pts, tuple(float(x) for x in t), 25, 25)
    for r in range(2):
        got = np.load(tmp_path / f"pts_rank{r}.npy")
        assert np.array_equal(got.view(np.uint8), first_pts.view(np.uint8)), "gathered points differ from raster order"
    slabs = [np.load(tmp_path / f"slab_rank{r}.npy") for r in range(2)]
    assert np.array_equal(np.concatenate(slabs, axis=2), grid)
    assert grid.max() > 0
    want = np.array([[float(i), float(i) * 2] for i in range(11)])
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"rows_rank{r}.npy"), want)


def test_balanced_slab_bounds_equalise_ray_work():
    """Boundaries are a partition of [0, Z), deterministic, and split the per-layer ray coverage evenly."""
    import numpy as np
    from icpb200 import dist as D
    rng = np.random.default_rng(1)
    ez = [rng.uniform(2.5, 4.9, 50000), rng.uniform(3.0, 4.5, 50000)]
    oz = [1.0, 1.2]
    for world in (2, 4, 8):
        b = D.balanced_slab_bounds(500, world, 0.01, oz, ez)
        assert b == D.balanced_slab_bounds(500, world, 0.01, oz, ez)
        assert b[0] == 0 and b[-1] == 500 and all(b[i] < b[i + 1] for i in range(world))
        cover = np.zeros(500)
        for o, e in zip(oz, ez):
            for z in e[::50]:
                cover[int(o / 0.01): int(z / 0.01) + 1] += 1
        work = [cover[b[i]: b[i + 1]].sum() for i in range(world)]
        assert max(work) < 1.15 * (sum(work) / world), (b, work)
    # degenerate input: every slab still owns at least one layer
    b = D.balanced_slab_bounds(16, 8, 0.01, [0.05], [np.array([0.05])])
    assert b[0] == 0 and b[-1] == 16 and all(b[i] < b[i + 1] for i in range(8))
