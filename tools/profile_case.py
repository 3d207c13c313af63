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
ap.add_argument("--map", action="store_true", help="also run back-projection + map integration")
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
ctx.set_profiling(True)
for r in range(a.repeat):
    work.copy_from(dat)
    res, _, _ = ctx.icp_register(work, tgt, a.iters, 0.0, 0.75, 0, nn_mode=1 if a.grid >= 0 else 0, grid_cell=max(a.grid, 0.0))
    print(f"n={dat.n} m={tgt.n} passes={res['nn_passes']} gpu_ms={res['gpu_ms']:.3f} "
          f"nn_partial_ms={res['nn_partial_ms']:.3f} qpt={res['nn_qpt']} splits={res['nn_splits']} cell={res['grid_cell_used']:.3f} "
          f"rescans={res['exact_rescans']}")
if a.map:
    m = ctx.map((300, 300, 250), 0.02)
    v = m.integrate_rays(dat, (5.0, 5.0, 5.0), 25, 25)
    print("map voxels visited", v)
ctx.close()
