"""Seeded synthetic Kinect-shaped depth frames.

An axis-aligned 6 x 6 x 5 m room with a few boxes, seen by a pin-hole camera
on a smooth seeded trajectory.  Depth is the z-depth of the first hit,
quantised round(z * 5000) to uint16 (pointcloud.cpp:37 scale), zeroed outside
[1000, 25000] (SLAM.hpp:15-16) and with a seeded 5 % Bernoulli dropout.  The
same bytes feed the CUDA path and the CPU oracle.
"""
import numpy as np

MASTER_SEED = 20261018

KINECT_V1 = dict(w=640, h=480, fx=468.60, fy=468.61, cx=318.27, cy=243.99)   # pointcloud.hpp:7-10
KINECT_V2 = dict(w=512, h=424, fx=363.58, fy=363.53, cx=250.32, cy=212.55)   # SLAM.cpp:26-29

ROOM = np.array([6.0, 6.0, 5.0])


def rot_axis_angle(axis, angle_rad):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle_rad) * K + (1 - np.cos(angle_rad)) * (K @ K)


def scene_boxes(seed=MASTER_SEED):
    """Four seeded boxes standing inside the room (lo, hi corners)."""
    rng = np.random.default_rng(seed ^ 0xB0C5)
    boxes = []
    for _ in range(4):
        size = rng.uniform([0.4, 0.4, 0.4], [1.2, 1.5, 1.2])
        lo = rng.uniform([0.3, 0.3, 0.3], ROOM - size - 0.3)
        boxes.append((lo, lo + size))
    return boxes


def render_depth(R_wc, t_wc, sensor=KINECT_V1, seed=MASTER_SEED, dropout=0.05, boxes=None):
    """Ray-cast one depth frame.  R_wc, t_wc: camera-to-world pose."""
    w, h = sensor["w"], sensor["h"]
    if boxes is None:
        boxes = scene_boxes()
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    dc = np.stack([(u - sensor["cx"]) / sensor["fx"], (v - sensor["cy"]) / sensor["fy"], np.ones_like(u)], axis=-1)
    d = dc @ np.asarray(R_wc).T            # world direction; ray parameter t == camera z-depth
    o = np.asarray(t_wc, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        # room: camera is inside, first wall hit is the smallest positive exit distance
        t_lo = (0.0 - o) * inv
        t_hi = (ROOM - o) * inv
        t_exit = np.where(d > 0, t_hi, np.where(d < 0, t_lo, np.inf)).min(axis=-1)
        depth = t_exit
        for lo, hi in boxes:
            t0 = (lo - o) * inv
            t1 = (hi - o) * inv
            tn = np.minimum(t0, t1).max(axis=-1)
            tf = np.maximum(t0, t1).min(axis=-1)
            hit = (tn <= tf) & (tn > 0)
            depth = np.where(hit & (tn < depth), tn, depth)
    q = np.rint(depth * 5000.0)
    q = np.where((q < 1000) | (q > 25000) | ~np.isfinite(q), 0, q).astype(np.uint16)
    if dropout > 0:
        rng = np.random.default_rng(seed ^ 0xD409)
        q[rng.random(q.shape) < dropout] = 0
    return q


def render_color(sensor=KINECT_V1, seed=MASTER_SEED):
    rng = np.random.default_rng(seed ^ 0xC0102)
    return rng.integers(0, 256, size=(sensor["h"], sensor["w"], 3), dtype=np.uint8)


def trajectory(n_frames, seed=MASTER_SEED, step_deg=0.4, step_m=0.01):
    """Smooth seeded camera path inside the room: list of (R_wc, t_wc)."""
    rng = np.random.default_rng(seed ^ 0x7243)
    R = rot_axis_angle([0, 1, 0], 0.3)
    t = np.array([3.0, 2.5, 1.2])
    axis = rng.standard_normal(3)
    vel = rng.standard_normal(3)
    poses = []
    for _ in range(n_frames):
        poses.append((R.copy(), t.copy()))
        axis = axis + 0.2 * rng.standard_normal(3)
        vel = vel + 0.2 * rng.standard_normal(3)
        R = R @ rot_axis_angle(axis, np.deg2rad(step_deg))
        t = np.clip(t + step_m * vel / np.linalg.norm(vel), [1.5, 1.5, 0.8], [4.5, 4.5, 2.0])
    return poses


def frame_pair(sensor=KINECT_V1, seed=MASTER_SEED, angle_deg=5.0, shift_m=0.05, dropout=0.05):
    """Config 1/2 frame pair: second camera = first moved by angle_deg about a
    seeded axis plus shift_m along a seeded direction.  Returns (depth_prev,
    depth_cur, color, (R_rel, t_rel))."""
    rng = np.random.default_rng(seed ^ 0xFA12)
    R0 = rot_axis_angle([0, 1, 0], 0.35)
    t0 = np.array([2.6, 2.4, 1.0])
    axis = rng.standard_normal(3)
    dirn = rng.standard_normal(3)
    dirn /= np.linalg.norm(dirn)
    R_rel = rot_axis_angle(axis, np.deg2rad(angle_deg))
    t_rel = shift_m * dirn
    R1 = R0 @ R_rel
    t1 = t0 + R0 @ t_rel
    d0 = render_depth(R0, t0, sensor, seed, dropout)
    d1 = render_depth(R1, t1, sensor, seed + 1, dropout)
    return d0, d1, render_color(sensor, seed), (R_rel, t_rel)


def subsample_exact(points, count, seed):
    """Seeded choice of exactly `count` points, raster order kept (config 1)."""
    n = len(points)
    if n <= count:
        return points
    rng = np.random.default_rng(seed ^ 0x5B5A)
    keep = np.sort(rng.choice(n, size=count, replace=False))
    return points[keep]
