#!/usr/bin/env python
"""Short single-purpose run for ncu: a few association passes at a chosen size (no torch import)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "icp-slam-prototype_b200", "python"))
import numpy as np  # noqa: E402
import icpb200  # noqa: E402
from icpb200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=0, help="0 = full resolution")
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--grid", type=float, default=-1.0, help="use ICPB_NN_GRID with this cell size (0 = default)")
ap.add_argument("--noprof", action="store_true", help="leave the per-launch event timing off")
ap.add_argument("--map", action="store_true", help="also run back-projection + map integration")
ap.add_argument("--cm", type=float, default=2.0, help="map cell in cm (2 -> 300x300x250, 1 -> 600x600x500)")
a = ap.parse_args()

d0, d1, col, _ = synth.frame_pair()
ctx = icpb200.Context(0)
K = icpb200.reference_intrinsics_v1()
cam = np.array([5, 5, 5], np.float32)
h, w = d0.shape
tgt, dat, work = ctx.cloud(w * h), ctx.cloud(w * h), ctx.cloud(w * h)
tgt.from_depth(d0, col, K); tgt.transform(None, cam)
dat.from_depth(d1, col, K); dat.transform(None, cam)
if a.points:
    tgt.upload(synth.subsample_exact(tgt.download(), a.points, 2))
    dat.upload(synth.subsample_exact(dat.download(), a.points, 1))
ctx.set_profiling(not a.noprof)
for r in range(a.repeat):
    work.copy_from(dat)
    res, _, _ = ctx.icp_register(work, tgt, a.iters, 0.0, 0.75, 0, nn_mode=1 if a.grid >= 0 else 0, grid_cell=max(a.grid, 0.0))
    print(f"n={dat.n} m={tgt.n} passes={res['nn_passes']} gpu_ms={res['gpu_ms']:.3f} "
          f"nn_partial_ms={res['nn_partial_ms']:.3f} qpt={res['nn_qpt']} splits={res['nn_splits']} cell={res['grid_cell_used']:.3f} "
          f"rescans={res['exact_rescans']}")
    if not a.noprof and a.grid >= 0:
        g_ms, g_k = ctx.profile_read(icpb200.PROF_NN_GRID)
        f_ms, f_k = ctx.profile_read(icpb200.PROF_NN_FINALIZE)
        print(f"  spans: nn_grid {g_ms:.3f} ms / {g_k} passes = {1e3 * g_ms / max(g_k, 1):.1f} us, "
              f"nn_finalize {f_ms:.3f} ms / {f_k} = {1e3 * f_ms / max(f_k, 1):.1f} us, rest {res['gpu_ms'] - g_ms - f_ms:.3f} ms")
if a.map:
    from icpb200 import synth as _s
    cell = a.cm / 100.0
    dims = (int(round(6 / cell)), int(round(6 / cell)), int(round(5 / cell)))
    m = ctx.map(dims, cell)
    poses = _s.trajectory(3, step_deg=0.8, step_m=0.02)
    wc = ctx.cloud(w * h)
    for f, (R, t) in enumerate(poses):
        dpt = _s.render_depth(R, t, _s.KINECT_V1, seed=f)
        wc.from_depth(dpt, None, K); wc.transform(R.astype(np.float32), t.astype(np.float32))
        ctx.timer_start()
        v = m.integrate_rays(wc, tuple(float(x) for x in t), 25, 25)
        print("map", dims, "frame", f, "rays", wc.n, "voxels visited", v, "ms", round(ctx.timer_stop(), 3))
ctx.close()
