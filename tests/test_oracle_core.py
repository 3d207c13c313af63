"""The oracle against independent restatements (numpy / pure Python / exact rationals) and its own
invariants.  These pin the reference semantics recorded in SURVEY.md section 8a."""
from fractions import Fraction

import numpy as np
import pytest


def _np_nn(data_xyz, target_xyz):
    """icp.cpp:566-620 in numpy: float diffs, double squares summed left to right, one rounding to
    float, float sqrt, first minimum wins."""
    idx = np.zeros(len(data_xyz), np.int32)
    dist = np.zeros(len(data_xyz), np.float32)
    t = target_xyz.astype(np.float32)
    for i, a in enumerate(data_xyz.astype(np.float32)):
        d = (a[None, :] - t).astype(np.float32).astype(np.float64)
        s = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32)
        r = np.sqrt(s)      # float32 sqrt, correctly rounded
        idx[i] = int(np.argmin(r))
        dist[i] = r[idx[i]]
    return idx, dist


def test_nn_vs_numpy_random_and_ties(orc):
    rng = np.random.default_rng(0)
    tx = rng.uniform(3, 8, (700, 3)).astype(np.float32)
    tx = np.concatenate([tx, tx[:100]])        # duplicates: lowest index must win
    dx = np.concatenate([rng.uniform(3, 8, (300, 3)), tx[50:80] + 1e-4]).astype(np.float32)
    idx, dist = orc.nn(orc.make_points(dx), orc.make_points(tx))
    ridx, rdist = _np_nn(dx, tx)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
    g = np.arange(5, dtype=np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    lat = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    q = lat[:40] + 0.5
    idx, dist = orc.nn(orc.make_points(q), orc.make_points(lat))
    ridx, rdist = _np_nn(q, lat)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)


def test_nn_threads_do_not_change_results(orc, pair10k):
    data, target = pair10k
    a = orc.nn(data[:500], target[:3000], 1)
    b = orc.nn(data[:500], target[:3000], 8)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_distance_is_sqrt_of_once_rounded_double_sum(orc):
    """N1: values where an all-float evaluation differs from the reference's double-then-float one."""
    rng = np.random.default_rng(1)
    a = rng.uniform(-8, 8, (20000, 3)).astype(np.float32)
    b = rng.uniform(-8, 8, (20000, 3)).astype(np.float32)
    d = (a - b).astype(np.float32)
    want = np.sqrt(((d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2)
                    + d[:, 2].astype(np.float64) ** 2).astype(np.float32))
    allf = np.sqrt(((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32))
    assert (want != allf).any()            # the distinction is real
    pa, pb = orc.make_points(a), orc.make_points(b)
    idx, dist = orc.nn(pa[:1], pb[:1])
    assert dist[0] == want[0]


def test_canon_reduce_is_the_documented_tree(orc):
    rng = np.random.default_rng(2)
    for n in [1, 31, 256, 257, 1000, 70000]:
        x = rng.standard_normal((n, 2)) * 10.0 ** rng.integers(-3, 6, (n, 2))
        got = orc.canon_reduce(x)

        def r256(v):
            v = v.copy()
            g = []
            for wv in range(8):
                s = v[32 * wv: 32 * wv + 32].copy()
                off = 16
                while off >= 1:
                    s[:off] = s[:off] + s[off: 2 * off]
                    off //= 2
                g.append(s[0])
            t = g[0]
            for wv in range(1, 8):
                t = t + g[wv]
            return t
        for q in range(2):
            chunks = (n + 255) // 256
            pad = np.zeros(chunks * 256); pad[:n] = x[:, q]
            part = [r256(pad[c * 256:(c + 1) * 256]) for c in range(chunks)]
            slots = np.zeros(256)
            for t in range(256):
                acc = 0.0
                for c in range(t, chunks, 256):
                    acc = acc + part[c]
                slots[t] = acc
            assert got[q] == r256(slots)
        assert np.allclose(got, x.sum(0), rtol=1e-9, atol=1e-6)


def test_voxel_coords_edges(orc):
    """M2 (map.cpp:55-85): truncation toward zero and clamping, incl. negative fractions and boundaries."""
    cell = float(np.float32(10.0) / np.float32(300.0))
    dims = (300, 300, 300)
    c32 = np.float32(cell)
    cases = [(-0.01, 0), (-5.0, 0), (0.0, 0), (float(c32), None), (9.9999, 299), (10.0, 299), (1e6, 299),
             (float(np.nextafter(c32, np.float32(0))), 0), (float(c32 * np.float32(7)), None)]
    for v, want in cases:
        got = orc.voxel_coords((v, 0.0, 0.0), cell, dims)[0]
        ref = int(np.float32(v) / c32)
        ref = min(max(ref, 0), 299)
        assert got == ref
        if want is not None:
            assert got == want


@pytest.mark.parametrize("rule,delta", [(0, 25), (0, 180), (1, 25), (1, 180), (0, 0)])
def test_endpoint_rules_are_order_independent_powers(orc, rule, delta):
    """M3: k hits on one voxel give f^k(c) whatever the order (map.cpp:104-113, 139-149, 249-253)."""
    def f(c):
        if rule == 0:
            return 255 if c > 255 - delta else c + delta
        return 255 if c >= 180 - delta else (c + delta) & 255
    dims, cell = (4, 4, 4), 1.0
    rng = np.random.default_rng(3)
    for _ in range(20):
        start = rng.integers(0, 256, dims).astype(np.uint8)
        vox = rng.integers(0, 4, (60, 3))
        pts = orc.make_points(vox + 0.5)
        g1 = start.copy(); orc.map_update_endpoints(g1, dims, cell, pts, rule, delta, 180)
        g2 = start.copy(); orc.map_update_endpoints(g2, dims, cell, pts[::-1].copy(), rule, delta, 180)
        assert np.array_equal(g1, g2)
        want = start.copy()
        for v in vox:
            want[tuple(v)] = f(int(want[tuple(v)]))
        assert np.array_equal(g1, want)


def _walk_fraction(o, e):
    """Independent statement of the M4 walk: sort wall crossings t = (2i+1)/(2 n_k) exactly, ties x<y<z."""
    n = [abs(e[k] - o[k]) for k in range(3)]
    sgn = [(e[k] > o[k]) - (e[k] < o[k]) for k in range(3)]
    events = []
    for k in range(3):
        for i in range(n[k]):
            events.append((Fraction(2 * i + 1, 2 * n[k]), k))
    events.sort()
    p = list(o)
    out = []
    for _, k in events:
        p[k] += sgn[k]
        out.append(tuple(p))
    return out


def test_ray_walk_matches_exact_rational_ordering(orc):
    rng = np.random.default_rng(4)
    dims, cell = (24, 24, 24), 1.0
    for _ in range(200):
        o = rng.integers(0, 24, 3)
        e = rng.integers(0, 24, 3)
        grid = np.full(dims, 200, np.uint8)
        v = orc.map_integrate_rays(grid, dims, cell, orc.make_points(e[None, :] + 0.5), o + 0.5, 25, 25)
        path = _walk_fraction(tuple(int(x) for x in o), tuple(int(x) for x in e))
        assert v == max(len(path) - 1, 0)
        want = np.full(dims, 200, np.uint8)
        for p in path[:-1]:
            want[p] -= 25
        want[tuple(e)] = 225
        assert np.array_equal(grid, want)
        if path:
            assert path[-1] == tuple(int(x) for x in e)


def test_ray_decrement_clamps_and_skips_zero(orc):
    dims, cell = (16, 4, 4), 1.0
    grid = np.zeros(dims, np.uint8)
    grid[3, 1, 1] = 10
    grid[5, 1, 1] = 200
    orc.map_integrate_rays(grid, dims, cell, orc.make_points([[12.5, 1.5, 1.5]]), (0.5, 1.5, 1.5), 25, 25)
    assert grid[3, 1, 1] == 0 and grid[5, 1, 1] == 175 and grid[12, 1, 1] == 25 and grid[0, 1, 1] == 0


def test_z_slab_union_equals_full(orc):
    rng = np.random.default_rng(5)
    dims, cell = (20, 20, 20), 0.5
    pts = orc.make_points(rng.uniform(0, 10, (3000, 3)))
    origin = (5.2, 4.9, 5.0)
    start = rng.integers(0, 80, dims).astype(np.uint8)
    full = start.copy(); orc.map_integrate_rays(full, dims, cell, pts, origin, 25, 25)
    acc = start.copy()
    for g in range(4):
        orc.map_integrate_rays(acc, dims, cell, pts, origin, 25, 25, g * 5, (g + 1) * 5)
    assert np.array_equal(acc, full)


def test_icp_reference_mode_bookkeeping(orc, pair10k):
    data, target = pair10k
    data, target = data[:1500], target[:2000]
    res, out, it, dt = orc.icp(data, target, 4, 0.0, 0.75, orc.SOLVE_REFERENCE, trace=True)
    assert res["iterations"] == 4 and res["nn_passes"] == 5
    # rigid column 3 is the LAST offset only (icp.cpp:266-268); cameraPosition accumulates all of them
    assert np.array_equal(res["rigid"][:3, 3], res["offset"])
    assert abs(np.linalg.det(res["rigid"][:3, :3].astype(np.float64)) - 1) < 1e-5
    # composed pose reproduces the transformed cloud to float accuracy
    moved = orc.xyz_of(data).astype(np.float64) @ res["pose_R"].T + res["pose_t"]
    assert np.abs(moved - orc.xyz_of(out)).max() < 2e-5


def test_icp_kabsch_recovers_known_motion(orc):
    from icpb200 import synth
    rng = np.random.default_rng(6)
    base = rng.uniform(3, 7, (1500, 3))
    R = synth.rot_axis_angle([0.3, -0.5, 0.8], np.deg2rad(2.0))
    t = np.array([0.02, -0.01, 0.015])
    moved = (base - 5.0) @ R.T + 5.0 + t
    res, out, _, _ = orc.icp(orc.make_points(base), orc.make_points(moved), 30, 0.0, 0.75, orc.SOLVE_KABSCH, n_threads=4)
    assert np.abs(orc.xyz_of(out) - moved).max() < 1e-4


def test_backproject_reference_quirk_and_rules(orc):
    """P1: y uses CX/FX like x (pointcloud.cpp:38-39); zero pixels are skipped before the subsample draw."""
    depth = np.array([[0, 5000, 10000], [2500, 0, 7500]], np.uint16)
    pts, cc, cr = orc.backproject(depth)
    assert len(pts) == 4
    K = orc.kinect_v1()
    z = np.float32(5000) / np.float32(5000.0)
    x = (np.float32(1) - np.float32(K.cx_u)) * z / np.float32(K.fx_u)
    y = (np.float32(0) - np.float32(K.cx_u)) * z / np.float32(K.fx_u)
    assert pts[0]["x"] == x and pts[0]["y"] == y and pts[0]["z"] == z
    assert np.allclose(cc, cr, atol=1e-6)
    s, _, _ = orc.backproject(depth, rule=orc.SUB_STRIDE, rule_arg=2)
    assert np.array_equal(s.view(np.uint8), pts[::2].copy().view(np.uint8))
    st, _, _ = orc.backproject(depth, rule=orc.SUB_STREAM, keep_stream=np.array([0, 1, 1, 0], np.uint8))
    assert np.array_equal(st.view(np.uint8), pts[1:3].copy().view(np.uint8))


def test_row_index_multiply_shift_is_the_exact_floor():
    """cloud.cu `row_of`: p / w as (p * (2^40 // w + 1)) >> 40.  The launcher enables it when w >= 2, the largest p it
    can see is below 2^24 and p * w < 2^40; inside that envelope it must be the exact floor for every p."""
    rng = np.random.default_rng(5)
    widths = [2, 3, 7, 8, 100, 512, 640, 641, 1920, 4095, 4096, 65535]
    for w in widths:
        magic = (1 << 40) // w + 1
        reach = min(1 << 24, (1 << 40) // w)
        p = np.unique(np.concatenate([
            np.arange(0, min(reach, 70000), dtype=np.uint64),
            rng.integers(0, reach, 200000, dtype=np.uint64),
            np.arange(max(reach - 70000, 0), reach, dtype=np.uint64),
            (np.arange(1, min(reach // w, 60000) + 1, dtype=np.uint64) * np.uint64(w)) - np.uint64(1),  # last pixel of a row
        ]))
        p = p[p < reach]
        got = (p * np.uint64(magic)) >> np.uint64(40)      # p < 2^24, magic <= 2^39 + 1: no overflow in 64 bits
        assert np.array_equal(got, p // np.uint64(w)), w
