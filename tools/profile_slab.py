#!/usr/bin/env python
"""Per-slab cost of the ray walk on ONE GPU: the z-slab handles of a multi-GPU run, one after the other.

  python tools/profile_slab.py 0,289,348,440,500 [frames]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "icp-slam-prototype_b200", "python"))
import numpy as np  # noqa: E402
import icpb200  # noqa: E402
from icpb200 import synth  # noqa: E402

bounds = [int(x) for x in sys.argv[1].split(",")]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dims, cell = (600, 600, 500), 0.01
ctx = icpb200.Context(0)
K = icpb200.reference_intrinsics_v1()
poses = synth.trajectory(32, step_deg=0.8, step_m=0.02)[:frames]
clouds = []
for f, (R, t) in enumerate(poses):
    c = ctx.cloud(640 * 480)
    c.from_depth(synth.render_depth(R, t, synth.KINECT_V1, seed=f), None, K)
    c.transform(R.astype(np.float32), t.astype(np.float32))
    clouds.append(c)
for lo, hi in zip(bounds[:-1], bounds[1:]):
    m = ctx.map(dims, cell, lo, hi)
    for rep in range(2):   # second pass: occupied bricks in place
        ctx.set_profiling(rep == 1)
        for c, (R, t) in zip(clouds, poses):
            m.integrate_rays(c, tuple(float(x) for x in t), 25, 25, False)
        ctx.sync()
    ms, k = ctx.profile_read(icpb200.PROF_MAP_RAYS)
    me, _ = ctx.profile_read(icpb200.PROF_MAP_ENDPOINTS)
    ctx.set_profiling(False)
    print(f"slab [{lo},{hi}) rays {1e3 * ms / k:.1f} us/frame endpoints {1e3 * me / k:.1f} us/frame")
    m.close()
ctx.close()
