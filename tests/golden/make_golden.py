#!/usr/bin/env python
"""Generates tests/golden/*.npz.  Run in the authoring container, where /root/reference exists:

    make -C oracle/refshim && python tests/golden/make_golden.py

Every `ref_*` array is produced by the reference's OWN icp.cpp / pointcloud.cpp / map.cpp (compiled
unmodified by path into oracle/_ref, see oracle/refshim).  `orc_*` arrays are the oracle's canonical
(FP64 block-ordered) results, which the CUDA path must reproduce bit for bit.  The fixtures are small
so that they can be committed; tests/test_golden.py checks the oracle and the GPU against them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "icp-slam-prototype_b200", "python"))
from oracle import oracle as orc  # noqa: E402
from oracle import ref  # noqa: E402
from icpb200 import synth  # noqa: E402

assert ref.available(), "build oracle/_ref first: make -C oracle/refshim"
rng = np.random.default_rng(20261018)

# ---- frame pair -> small clouds
d0, d1, col, _ = synth.frame_pair()
cam = np.array([5, 5, 5], np.float32)
p0 = orc.translate(orc.backproject(d0, col)[0], cam)
p1 = orc.translate(orc.backproject(d1, col)[0], cam)

# 1. nearest neighbour: real surface points + duplicates + a lattice with exact ties
tgt = synth.subsample_exact(p0, 900, 11)
g = np.arange(0, 5, dtype=np.float32) * 0.25 + 4.0
X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
lat = orc.make_points(np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1))
tgt = np.concatenate([tgt, tgt[:100], lat])
tgt["c0"] = np.arange(len(tgt)) % 251
dat = np.concatenate([synth.subsample_exact(p1, 250, 12), orc.make_points(lat_q := (np.stack(
    [X.ravel(), Y.ravel(), Z.ravel()], 1)[:40] + np.float32(0.125)))])
b, d = ref.nearest(dat, tgt)
idx, dist = orc.nn(dat, tgt)
assert np.array_equal(d, dist) and np.array_equal(b.view(np.uint8), tgt[idx].view(np.uint8))
np.savez_compressed(os.path.join(HERE, "nn_small.npz"), data=dat, target=tgt, ref_nearest=b, ref_dist=d, idx=idx)

# 2. back-projection of a 64x48 crop with the reference's rand() stream (srand(1))
crop = np.ascontiguousarray(d0[200:248, 300:364])
ccol = np.ascontiguousarray(col[200:248, 300:364])
crop[rng.random(crop.shape) < 0.1] = 0
pts, dec, center = ref.backproject(crop, ccol, seed=1)
full, _, _ = ref.backproject(crop, ccol, seed=1)
np.savez_compressed(os.path.join(HERE, "backproject_small.npz"), depth=crop, bgr=ccol, ref_points=pts,
                    ref_decisions=dec, ref_center=center, orc_all_points=orc.backproject(crop, ccol)[0])

# 3. certainty grid, reference macros (300^3, cell 10/300): sparse dump after repeated updates
mp = synth.subsample_exact(p1, 3000, 13)
out = {}
for name, kind, rule, delta in [("A180", "cloud", 0, 180), ("A25", "assoc", 0, 25), ("C25", "nonassoc", 1, 25)]:
    ref.map_reset()
    for _ in range(3):
        ref.map_update(mp, delta, kind)
    w = ref.map_world()
    nz = np.flatnonzero(w)
    out[f"ref_{name}_idx"] = nz.astype(np.int32)
    out[f"ref_{name}_val"] = w.ravel()[nz]
ref.map_reset()
np.savez_compressed(os.path.join(HERE, "map_small.npz"), points=mp, **out)

# 4. registration loop, 5 iterations, 800 x 1000 points
di = synth.subsample_exact(p1, 800, 14)
ti = synth.subsample_exact(p0, 1000, 15)
r, rout = ref.icp_allpoints(di, ti, 5, 0.0)
fix = dict(data=di, target=ti, ref_rigid=r["rigid"], ref_cam_rotation=r["cam_rotation"],
           ref_cam_position=r["cam_position"], ref_out=rout, ref_mse=np.float32(r["mse"]), ref_n_assoc=r["n_assoc"])
for mode, tag in [(orc.SOLVE_REFERENCE, "refmode"), (orc.SOLVE_KABSCH, "kabsch")]:
    o, oout, it, dt = orc.icp(di, ti, 5, 0.0, 0.75, mode, trace=True)
    fix.update({f"orc_{tag}_idx": it, f"orc_{tag}_dist": dt, f"orc_{tag}_out": oout, f"orc_{tag}_pose_R": o["pose_R"],
                f"orc_{tag}_pose_t": o["pose_t"], f"orc_{tag}_rigid": o["rigid"], f"orc_{tag}_mse": np.float32(o["mse"])})
np.savez_compressed(os.path.join(HERE, "icp_small.npz"), **fix)

# 5. ray integration (builder-defined M4 semantics; oracle only)
dims, cell = (48, 48, 40), 0.125
rp = synth.subsample_exact(p1, 1500, 16)
start = rng.integers(0, 120, dims).astype(np.uint8)
grid = start.copy()
visited = orc.map_integrate_rays(grid, dims, cell, rp, (5.0, 5.0, 5.0), 25, 25)
np.savez_compressed(os.path.join(HERE, "rays_small.npz"), points=rp, start=start, orc_grid=grid, orc_visited=visited,
                    origin=np.array([5.0, 5.0, 5.0], np.float32), dims=np.array(dims), cell=np.float32(cell))
print("golden fixtures written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
