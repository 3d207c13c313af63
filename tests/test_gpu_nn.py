"""GPU parity: brute-force NN (N1-N3) through the C-ABI vs the CPU oracle.  Bar: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(ctx, orc, data, target, expect_rescans=None):
    import icpb200
    dc = ctx.cloud_from_points(data)
    tc = ctx.cloud_from_points(target)
    idx, dist, resc = ctx.nn_search(dc, tc)
    ridx, rdist = orc.nn(data, target, n_threads=8)
    dc.close(); tc.close()
    assert np.array_equal(idx, ridx), f"{(idx != ridx).sum()} index mismatches"
    assert np.array_equal(dist.view(np.uint32), rdist.view(np.uint32)), "distances not bit-equal"
    if expect_rescans is not None:
        assert expect_rescans(resc), resc
    return resc


def test_nn_config1(ctx, orc, pair10k):
    data, target = pair10k
    _check(ctx, orc, data, target)


@pytest.mark.parametrize("n,m", [(1, 1), (1, 33), (7, 5), (31, 32), (257, 1000), (1025, 4097), (3000, 31)])
def test_nn_ragged_sizes(ctx, orc, n, m):
    rng = np.random.default_rng(n * 1000 + m)
    data = orc.make_points(rng.uniform(3, 8, (n, 3)))
    target = orc.make_points(rng.uniform(3, 8, (m, 3)))
    _check(ctx, orc, data, target)


def test_nn_lattice_ties(ctx, orc):
    """Targets on a lattice, queries at cell centres: many exactly equal distances -> lowest index wins."""
    g = np.arange(0, 8, dtype=np.float32) * 0.25 + 4.0
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    target = orc.make_points(np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1))
    q = g[:-1] + 0.125
    X, Y, Z = np.meshgrid(q, q, q, indexing="ij")
    data = orc.make_points(np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1))
    resc = _check(ctx, orc, data, target, expect_rescans=lambda r: r > 0)
    assert resc > 0


def test_nn_duplicate_targets(ctx, orc):
    rng = np.random.default_rng(5)
    base = rng.uniform(3, 8, (500, 3)).astype(np.float32)
    target = orc.make_points(np.concatenate([base, base, base[::-1]], 0))   # every point three times
    data = orc.make_points(base[:200] + rng.normal(0, 1e-3, (200, 3)).astype(np.float32))
    _check(ctx, orc, data, target, expect_rescans=lambda r: r > 0)


def test_nn_identical_clouds(ctx, orc):
    rng = np.random.default_rng(6)
    pts = orc.make_points(rng.uniform(3, 8, (2048, 3)))
    dc = ctx.cloud_from_points(pts)
    tc = ctx.cloud_from_points(pts)
    idx, dist, _ = ctx.nn_search(dc, tc)
    assert np.array_equal(idx, np.arange(2048)) and np.all(dist == 0)


def test_nn_sqrt_collapsed_ties(ctx, orc):
    """Distinct squared distances that round to the same float sqrt: ties are judged on the sqrt (icp.cpp:578)."""
    rng = np.random.default_rng(8)
    n = 512
    data = orc.make_points(np.tile(np.array([[5.0, 5.0, 5.0]], np.float32), (n, 1)))
    # targets at nearly equal radius from the query: radii differ by a few float ulps
    dirs = rng.standard_normal((4096, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    radius = 1.0 + rng.integers(0, 4, 4096) * 2.0 ** -24
    target = orc.make_points(np.array([5.0, 5.0, 5.0]) + dirs * radius[:, None])
    _check(ctx, orc, data, target, expect_rescans=lambda r: r > 0)


def test_nn_far_and_near_scales(ctx, orc):
    rng = np.random.default_rng(9)
    data = orc.make_points(rng.uniform(-50, 50, (1500, 3)))
    target = orc.make_points(np.concatenate([rng.uniform(-50, 50, (2500, 3)), rng.uniform(-1e-3, 1e-3, (500, 3))]))
    _check(ctx, orc, data, target)


def test_nn_empty_cloud_is_an_error(ctx, orc):
    import icpb200
    dc = ctx.cloud(4)
    tc = ctx.cloud_from_points(orc.make_points(np.ones((3, 3), np.float32)))
    with pytest.raises(icpb200.IcpbError) as e:
        ctx.nn_search(dc, tc)
    assert e.value.status == icpb200.ERR_EMPTY


def test_nn_spread_queries_with_near_ties(ctx, orc):
    """Worst case for the centred filter: the queries one thread owns are metres apart (large |a - c|^2, hence a
    wide error band) and every query sees many targets at radii that differ by a few float ulps.  The band must
    flag those queries for exact evaluation; results stay bit-exact."""
    rng = np.random.default_rng(10)
    centres = np.array([[3.0, 3.5, 4.0], [8.0, 7.5, 3.0], [5.5, 2.5, 7.75]], np.float32)
    n = 768
    data = orc.make_points(centres[np.arange(n) % 3] + rng.normal(0, 1e-4, (n, 3)).astype(np.float32))
    tg = []
    for c in centres:
        dirs = rng.standard_normal((1500, 3))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        radius = 0.05 * (1.0 + rng.integers(0, 6, 1500) * 2.0 ** -23)
        tg.append(c + dirs * radius[:, None])
    target = orc.make_points(np.concatenate(tg).astype(np.float32)[rng.permutation(4500)])
    _check(ctx, orc, data, target, expect_rescans=lambda r: r > 0)


def test_nn_surface_like_clouds_rarely_rescan(ctx, orc, pair10k):
    """On Kinect-shaped clouds the three-best-group records settle almost every query without the full exact scan."""
    data, target = pair10k
    resc = _check(ctx, orc, data, target)
    assert resc < len(data) // 100, resc


@pytest.mark.parametrize("flt", ["1", "2", "3"])
def test_nn_every_filter_is_exact(ctx, orc, monkeypatch, flt):
    """ICPB_FILTER_DIRECT, _WARP and _CENTRED feed the same exact resolution."""
    monkeypatch.setenv("ICPB_NN_FILTER", flt)
    rng = np.random.default_rng(11)
    data = orc.make_points(rng.uniform(3, 8, (3001, 3)))
    target = orc.make_points(rng.uniform(3, 8, (5003, 3)))
    _check(ctx, orc, data, target)


@pytest.mark.parametrize("flt", ["2", "3"])
def test_nn_centred_filters_on_tie_cases(ctx, orc, monkeypatch, flt):
    """The lattice / duplicate / near-tie cases again with each centred filter forced (AUTO picks by size)."""
    monkeypatch.setenv("ICPB_NN_FILTER", flt)
    test_nn_lattice_ties(ctx, orc)
    test_nn_duplicate_targets(ctx, orc)
    test_nn_sqrt_collapsed_ties(ctx, orc)
    test_nn_spread_queries_with_near_ties(ctx, orc)
    test_nn_far_and_near_scales(ctx, orc)
    for n, m in [(1, 1), (31, 32), (257, 1000), (1025, 4097)]:
        test_nn_ragged_sizes(ctx, orc, n, m)


def test_nn_large_cloud_takes_the_warp_filter(ctx, orc):
    """>= 50,000 queries: ICPB_FILTER_AUTO orders the queries along a Morton curve and centres per warp.  Surface-like
    clouds (two noisy sheets) and a shuffled query order -- the ordering is rebuilt from the coordinates."""
    rng = np.random.default_rng(21)
    n, m = 60000, 24000
    uv = rng.uniform(0, 3, (n, 2))
    q = np.stack([3 + uv[:, 0], 4 + uv[:, 1], 5 + 0.3 * np.sin(uv[:, 0] * 3) + rng.normal(0, 2e-3, n)], 1)
    uv = rng.uniform(0, 3, (m, 2))
    t = np.stack([3.01 + uv[:, 0], 4.02 + uv[:, 1], 5.01 + 0.3 * np.sin(uv[:, 0] * 3) + rng.normal(0, 2e-3, m)], 1)
    data, target = orc.make_points(q[rng.permutation(n)]), orc.make_points(t)
    resc = _check(ctx, orc, data, target)
    assert resc < n // 200, resc


def _random_case(rng, orc):
    """One random association problem: sizes, coordinate scale, geometry and ordering all drawn from the seed."""
    n = int(rng.integers(1, 3000)); m = int(rng.integers(1, 4000))
    scale = float(10.0 ** rng.uniform(-2, 1.5)); origin = rng.uniform(-20, 20, 3)
    kind = int(rng.integers(0, 5))
    if kind == 0:      # uniform boxes
        q = rng.uniform(0, 1, (n, 3)); t = rng.uniform(0, 1, (m, 3))
    elif kind == 1:    # two noisy sheets (surface-like)
        uv = rng.uniform(0, 1, (n, 2)); q = np.stack([uv[:, 0], uv[:, 1], 0.1 * np.sin(6 * uv[:, 0]) + rng.normal(0, 1e-3, n)], 1)
        uv = rng.uniform(0, 1, (m, 2)); t = np.stack([uv[:, 0], uv[:, 1], 0.1 * np.sin(6 * uv[:, 0]) + rng.normal(0, 1e-3, m)], 1)
    elif kind == 2:    # clusters with duplicates among the targets
        c = rng.uniform(0, 1, (8, 3))
        q = c[rng.integers(0, 8, n)] + rng.normal(0, 0.01, (n, 3))
        t = c[rng.integers(0, 8, m)] + rng.normal(0, 0.01, (m, 3))
        t[rng.integers(0, m, m // 3 + 1)] = t[rng.integers(0, m, m // 3 + 1)]
    elif kind == 3:    # integer lattice: masses of exact ties
        q = rng.integers(0, 6, (n, 3)) / 4.0 + 0.125; t = rng.integers(0, 6, (m, 3)) / 4.0
    else:              # queries are copies of targets (distance 0) mixed with far outliers
        t = rng.uniform(0, 1, (m, 3)); q = t[rng.integers(0, m, n)].copy()
        far = rng.random(n) < 0.2; q[far] += rng.uniform(2, 5, (int(far.sum()), 3))
    return orc.make_points(q * scale + origin), orc.make_points(t * scale + origin)


@pytest.mark.parametrize("flt", ["0", "1", "2", "3"])
def test_nn_randomized_differential(ctx, orc, monkeypatch, flt):
    """40 random problems per filter (AUTO, DIRECT, WARP, CENTRED) against the oracle: indices and distances bit-equal."""
    monkeypatch.setenv("ICPB_NN_FILTER", flt)
    rng = np.random.default_rng(1000 + int(flt))
    for _ in range(40):
        data, target = _random_case(rng, orc)
        _check(ctx, orc, data, target)


# ---- the error band of the FP32 filter, stressed from both sides --------------------------

def _band_sweep_clouds(orc, offset):
    """4,096 queries next to the bisector plane of two targets that sit in DIFFERENT 32-target groups: the relative
    difference of their two squared distances sweeps -1.3e-5 .. +1.3e-5 in steps of ~6e-9 -- through zero, through the
    filter's own rounding noise and through the band |W2 - W1| = 18uA + 115uD (~7e-6 D) on either side, so that
    "the best group alone decides", "the best two groups decide" and "full exact rescan" are all taken, each within an ulp
    of its threshold.  Fillers are metres away.  `offset` moves everything to large coordinates (coarser ulps)."""
    rng = np.random.default_rng(11)
    n = 4096
    k = np.arange(n, dtype=np.float64) - n / 2
    q = np.stack([k * 5e-10, rng.uniform(-0.2, 0.2, n), rng.uniform(-0.2, 0.2, n)], 1)
    t = rng.uniform(2.0, 3.0, (96, 3)) * rng.choice([-1.0, 1.0], (96, 3))
    t[3] = (0.3, 0.0, 0.0)        # group 0
    t[70] = (-0.3, 0.0, 0.0)      # group 2
    off = np.asarray(offset, np.float64)
    return orc.make_points((q + off).astype(np.float32)), orc.make_points((t + off).astype(np.float32))


@pytest.mark.parametrize("offset", [(0, 0, 0), (0.001, -0.002, 0.0005), (5, 5, 5)])
@pytest.mark.parametrize("mode", ["direct", "warp", "centred", "grid"])
def test_nn_band_is_right_on_both_sides(ctx, orc, mode, offset):
    import icpb200
    data, target = _band_sweep_clouds(orc, offset)
    ridx, rdist = orc.nn(data, target, n_threads=8)
    assert 0.3 < (ridx == 3).mean() < 0.7 and set(np.unique(ridx)) == {3, 70}   # the sweep really crosses the bisector
    dc, tc = ctx.cloud_from_points(data), ctx.cloud_from_points(target)
    kw = dict(nn_mode=icpb200.NN_GRID) if mode == "grid" else dict(
        nn_filter={"direct": icpb200.FILTER_DIRECT, "warp": icpb200.FILTER_WARP, "centred": icpb200.FILTER_CENTRED}[mode])
    res, it, dt = ctx.icp_register(dc, tc, 0, 0.0, 0.75, icpb200.SOLVE_REFERENCE, trace=True, **kw)
    assert np.array_equal(it[0], ridx), f"{(it[0] != ridx).sum()} index mismatches"
    assert np.array_equal(dt[0].view(np.uint32), rdist.view(np.uint32))
    dc.close(); tc.close()


def test_nn_warp_filter_is_deterministic(ctx, orc):
    """The query order behind the warp-centred filter is a stable radix sort: 20 runs of a 60k-query search give the
    same indices, distances AND the same number of exact rescans (round 1 used an atomic cursor: 11..17 rescans)."""
    import icpb200
    rng = np.random.default_rng(3)
    base = rng.uniform(4, 6, (60000, 3)).astype(np.float32)
    data = orc.make_points(base + rng.normal(0, 0.02, base.shape).astype(np.float32))
    target = orc.make_points(np.concatenate([base[::2], base[::2] + np.float32(1e-4)]))
    dc0, tc = ctx.cloud_from_points(data), ctx.cloud_from_points(target)
    dc = ctx.cloud(len(data))
    first = None
    for run in range(20):
        dc.copy_from(dc0)
        res, it, dt = ctx.icp_register(dc, tc, 0, 0.0, 0.75, icpb200.SOLVE_REFERENCE, trace=True, nn_filter=icpb200.FILTER_WARP)
        sig = (it[0].tobytes(), dt[0].tobytes(), res["exact_rescans"])
        if first is None:
            first = sig
            ridx, rdist = orc.nn(data[:3000], target, n_threads=8)
            assert np.array_equal(it[0][:3000], ridx) and np.array_equal(dt[0][:3000].view(np.uint32), rdist.view(np.uint32))
        assert sig == first, f"run {run} differs from run 0 (rescans {res['exact_rescans']} vs {first[2]})"
    dc.close(); dc0.close(); tc.close()
