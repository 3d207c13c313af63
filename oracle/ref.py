"""ctypes binding of oracle/_ref/libicpref.so: the reference's OWN icp.cpp / pointcloud.cpp / map.cpp
compiled unmodified by path against oracle/refshim (TEST INFRASTRUCTURE, never shipped or timed as product)."""
import ctypes as C
import os

import numpy as np

from .oracle import POINT_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libicpref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_distance.restype = C.c_float
        _lib.ref_mse.restype = C.c_float
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def distance(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return float(lib().ref_distance(_p(a), _p(b)))


def backproject(depth, bgr=None, seed=1):
    """Returns (points, decisions, center): decisions[k] is the reference's rand()%40==0 draw for non-zero pixel k."""
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = depth.shape
    out = np.zeros(h * w, dtype=POINT_DTYPE)
    dec = np.zeros(int((depth != 0).sum()) + 1, dtype=np.uint8)
    center = np.zeros(3, dtype=np.float32)
    if bgr is not None:
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    n = lib().ref_backproject(_p(depth), _p(bgr), w, h, C.c_uint(seed), _p(out), _p(dec), _p(center))
    return out[:n].copy(), dec[:-1].copy(), center


def rotate(pts, R):
    pts = np.ascontiguousarray(pts).copy()
    R = np.ascontiguousarray(R, dtype=np.float32).reshape(9)
    lib().ref_rotate(_p(pts), len(pts), _p(R))
    return pts


def translate(pts, t):
    pts = np.ascontiguousarray(pts).copy()
    t = np.ascontiguousarray(t, dtype=np.float32).reshape(3)
    lib().ref_translate(_p(pts), len(pts), _p(t))
    return pts


def nn_assoc(data, target):
    data = np.ascontiguousarray(data); target = np.ascontiguousarray(target)
    a = np.zeros(len(data), dtype=POINT_DTYPE); b = np.zeros(len(data), dtype=POINT_DTYPE)
    e = np.zeros(len(data), dtype=np.float32)
    k = lib().ref_nn_assoc(_p(data), len(data), _p(target), len(target), _p(a), _p(b), _p(e))
    return a[:k].copy(), b[:k].copy(), e[:k].copy()


def nearest(data, target):
    data = np.ascontiguousarray(data); target = np.ascontiguousarray(target)
    b = np.zeros(len(data), dtype=POINT_DTYPE); d = np.zeros(len(data), dtype=np.float32)
    lib().ref_nearest(_p(data), len(data), _p(target), len(target), _p(b), _p(d))
    return b, d


def mse(errors):
    errors = np.ascontiguousarray(errors, dtype=np.float32)
    return float(lib().ref_mse(_p(errors), len(errors)))


def calculate_offset(a, b):
    """icp::calculateOffset (icp.cpp:314-344) over the pairs (a[i], b[i])."""
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    out = np.zeros(3, dtype=np.float32)
    lib().ref_calculate_offset(_p(a), _p(b), len(a), _p(out))
    return out


def make_rotation(x, y, z):
    out = np.zeros(9, dtype=np.float32)
    lib().ref_make_rotation(C.c_float(x), C.c_float(y), C.c_float(z), _p(out))
    return out.reshape(3, 3)


def voxel(p):
    p = np.ascontiguousarray(p, dtype=np.float32).reshape(3)
    v = np.zeros(3, dtype=np.int32)
    lib().ref_voxel(_p(p), _p(v))
    return tuple(int(x) for x in v)


def map_reset():
    lib().ref_map_reset()


def map_update(pts, delta, kind):
    pts = np.ascontiguousarray(pts)
    fn = {"cloud": lib().ref_map_update_cloud, "nonassoc": lib().ref_map_update_nonassoc,
          "assoc": lib().ref_map_update_assoc}[kind]
    fn(_p(pts), len(pts), int(delta))


def map_cloud(kind, capacity=1 << 20):
    """mapCloud.keypoints (kind 0) / mapCloud.points (kind 1) of the reference's global map."""
    out = np.zeros(capacity, dtype=POINT_DTYPE)
    n = lib().ref_map_cloud(int(kind), _p(out), capacity)
    return out[:n].copy()


def map_world():
    out = np.zeros((300, 300, 300), dtype=np.uint8)
    lib().ref_map_world(_p(out))
    return out


def icp_allpoints(data, target, max_iterations, threshold):
    data = np.ascontiguousarray(data).copy(); target = np.ascontiguousarray(target)
    rigid = np.zeros(16, np.float32); camR = np.zeros(9, np.float32); camP = np.zeros(3, np.float32)
    mse_ = C.c_float(0); na = C.c_int(0)
    it = lib().ref_icp_allpoints(_p(data), len(data), _p(target), len(target), int(max_iterations),
                                 C.c_float(threshold), _p(rigid), _p(camR), _p(camP), C.byref(mse_), C.byref(na))
    return {"iterations": it, "rigid": rigid.reshape(4, 4), "cam_rotation": camR.reshape(3, 3),
            "cam_position": camP, "mse": mse_.value, "n_assoc": na.value}, data


# ---- 8f-4: the reference's own Quaternion class; q = [w, x, y, z]
def quat_from_rot(R):
    R = np.ascontiguousarray(R, dtype=np.float32).reshape(9); q = np.zeros(4, np.float32)
    lib().ref_quat_from_rot(_p(R), _p(q))
    return q


def quat_mul(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
    o = np.zeros(4, np.float32)
    lib().ref_quat_mul(_p(a), _p(b), _p(o))
    return o


def quat_inverse(a):
    a = np.ascontiguousarray(a, dtype=np.float32); o = np.zeros(4, np.float32)
    lib().ref_quat_inverse(_p(a), _p(o))
    return o


class _quiet_stdout:
    """The reference prints its MSE and map size to std::cout (icp.cpp:264,279): keep that out of the caller's stdout."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *exc):
        lib().ref_flush_stdout()
        os.dup2(self._saved, 1)
        os.close(self._null); os.close(self._saved)


def get_transformation(depth_cur, depth_prev, bgr, keypoints_xy, max_iterations=16, threshold=1e-4, seed=1):
    """The reference's own icp::getTransformation (live key-point variant) on its process-global map / camera pose.
    Returns (rigid 4x4, cameraRotation, cameraPosition) after the call."""
    depth_cur = np.ascontiguousarray(depth_cur, dtype=np.uint16); depth_prev = np.ascontiguousarray(depth_prev, dtype=np.uint16)
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    kp = np.ascontiguousarray(keypoints_xy, dtype=np.float32).reshape(-1, 2)
    h, w = depth_cur.shape
    rigid = np.zeros(16, np.float32); camR = np.zeros(9, np.float32); camP = np.zeros(3, np.float32)
    with _quiet_stdout():
        lib().ref_get_transformation(_p(depth_cur), _p(depth_prev), _p(bgr), w, h, _p(kp), len(kp), int(max_iterations),
                                     C.c_float(threshold), C.c_uint(seed), _p(rigid), _p(camR), _p(camP))
    return rigid.reshape(4, 4), camR.reshape(3, 3), camP


def nearest_mt(data, target, threads=1):
    """getNearestPoint for every query on `threads` OpenMP threads (bench.py's CPU baseline); returns the distances."""
    data = np.ascontiguousarray(data); target = np.ascontiguousarray(target)
    d = np.zeros(len(data), dtype=np.float32)
    lib().ref_nearest_mt(_p(data), len(data), _p(target), len(target), _p(d), int(threads))
    return d
