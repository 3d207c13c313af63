"""ctypes binding of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  See oracle/icp_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libicp_oracle.so")

POINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("c0", "u1"), ("c1", "u1"), ("c2", "u1"), ("pad", "u1")]
)
assert POINT_DTYPE.itemsize == 16

SUB_NONE, SUB_STRIDE, SUB_HASH, SUB_STREAM = 0, 1, 2, 3
SOLVE_REFERENCE, SOLVE_KABSCH = 0, 1
RULE_A, RULE_C = 0, 1


class Intrinsics(C.Structure):
    _fields_ = [("fx_u", C.c_float), ("cx_u", C.c_float), ("fx_v", C.c_float), ("cx_v", C.c_float),
                ("depth_scale", C.c_float)]


class IcpParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("threshold", C.c_float), ("max_nn_distance", C.c_float),
                ("solve_mode", C.c_int), ("last_translation", C.c_float * 3), ("n_threads", C.c_int)]


class IcpResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("nn_passes", C.c_int), ("n_assoc", C.c_int), ("mse", C.c_float),
                ("rigid", C.c_float * 16), ("cam_rotation", C.c_float * 9), ("cam_position", C.c_float * 3),
                ("offset", C.c_float * 3), ("pose_R", C.c_double * 9), ("pose_t", C.c_double * 3),
                ("small_assoc_exit", C.c_int)]


def build(force=False):
    """Compile the oracle with gcc (idempotent)."""
    src = os.path.join(_HERE, "icp_oracle.c")
    hdr = os.path.join(_HERE, "icp_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_hash32.restype = C.c_uint32
        _lib.orc_hash32.argtypes = [C.c_uint32, C.c_uint32]
        _lib.orc_distance.restype = C.c_float
        _lib.orc_det33f.restype = C.c_double
        _lib.orc_map_integrate_rays.restype = C.c_longlong
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def kinect_v1():
    """pointcloud.hpp:7-10 with the reference's CX/FX-on-both-axes quirk (pointcloud.cpp:38-39)."""
    return Intrinsics(468.60, 318.27, 468.60, 318.27, 5000.0)


def kinect_v2():
    """SLAM.cpp:26-29 values, same quirk."""
    return Intrinsics(363.58, 250.32, 363.58, 250.32, 5000.0)


def make_points(xyz, color=None):
    xyz = np.asarray(xyz, dtype=np.float32)
    pts = np.zeros(xyz.shape[0], dtype=POINT_DTYPE)
    pts["x"], pts["y"], pts["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    if color is not None:
        pts["c0"], pts["c1"], pts["c2"] = color[:, 0], color[:, 1], color[:, 2]
    return pts


def xyz_of(pts):
    return np.stack([pts["x"], pts["y"], pts["z"]], axis=1)


def backproject(depth, bgr=None, K=None, rule=SUB_NONE, rule_arg=1, seed=0, keep_stream=None):
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = depth.shape
    K = K or kinect_v1()
    out = np.zeros(h * w, dtype=POINT_DTYPE)
    n = C.c_int(0)
    cc = (C.c_double * 3)()
    cr = (C.c_float * 3)()
    bgr_p = None
    if bgr is not None:
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        bgr_p = _p(bgr)
    ks_p = None
    if keep_stream is not None:
        keep_stream = np.ascontiguousarray(keep_stream, dtype=np.uint8)
        ks_p = _p(keep_stream)
    rc = lib().orc_backproject(_p(depth), bgr_p, w, h, C.byref(K), rule, C.c_uint32(rule_arg), C.c_uint32(seed),
                               ks_p, _p(out), C.byref(n), cc, cr)
    assert rc == 0
    return out[: n.value].copy(), np.array(cc[:]), np.array(cr[:], dtype=np.float32)


def rotate(pts, R):
    pts = pts.copy()
    R = np.ascontiguousarray(R, dtype=np.float32).reshape(9)
    lib().orc_rotate(_p(pts), len(pts), _p(R))
    return pts


def translate(pts, t):
    pts = pts.copy()
    t = np.ascontiguousarray(t, dtype=np.float32).reshape(3)
    lib().orc_translate(_p(pts), len(pts), _p(t))
    return pts


def normals(depth):
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = depth.shape
    out = np.zeros((h, w, 3), dtype=np.float32)
    lib().orc_normals(_p(depth), w, h, _p(out))
    return out


def depth_filter(depth, min_d=1000, max_d=25000):
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = depth.shape
    out = np.zeros((h, w), dtype=np.uint16)
    lib().orc_depth_filter(_p(depth), w, h, int(min_d), int(max_d), _p(out))
    return out


def nn(data, target, n_threads=1):
    data = np.ascontiguousarray(data)
    target = np.ascontiguousarray(target)
    idx = np.zeros(len(data), dtype=np.int32)
    dist = np.zeros(len(data), dtype=np.float32)
    lib().orc_nn(_p(data), len(data), _p(target), len(target), _p(idx), _p(dist), int(n_threads))
    return idx, dist


def canon_reduce(terms):
    terms = np.ascontiguousarray(terms, dtype=np.float64)
    n, k = terms.shape
    out = np.zeros(k, dtype=np.float64)
    lib().orc_canon_reduce(_p(terms), n, k, _p(out))
    return out


def svd3(A):
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(9)
    U = np.zeros(9); w = np.zeros(3); Vt = np.zeros(9)
    lib().orc_svd3(_p(A), _p(U), _p(w), _p(Vt))
    return U.reshape(3, 3), w, Vt.reshape(3, 3)


def gemm33f(A, B):
    A = np.ascontiguousarray(A, dtype=np.float32).reshape(9)
    B = np.ascontiguousarray(B, dtype=np.float32).reshape(9)
    Cm = np.zeros(9, dtype=np.float32)
    lib().orc_gemm33f(_p(A), _p(B), _p(Cm))
    return Cm.reshape(3, 3)


def inv33f(A):
    A = np.ascontiguousarray(A, dtype=np.float32).reshape(9)
    D = np.zeros(9, dtype=np.float32)
    lib().orc_inv33f(_p(A), _p(D))
    return D.reshape(3, 3)


def det33f(A):
    A = np.ascontiguousarray(A, dtype=np.float32).reshape(9)
    return float(lib().orc_det33f(_p(A)))


def icp(data, target, max_iterations=20, threshold=0.0, max_nn_distance=0.75, solve_mode=SOLVE_REFERENCE,
        last_translation=(0, 0, 0), n_threads=1, trace=False):
    """Returns (result dict, transformed data, idx_trace or None, dist_trace or None)."""
    data = np.ascontiguousarray(data).copy()
    target = np.ascontiguousarray(target)
    prm = IcpParams(max_iterations, threshold, max_nn_distance, solve_mode,
                    (C.c_float * 3)(*last_translation), n_threads)
    res = IcpResult()
    it = dt = None
    itp = dtp = None
    if trace:
        it = np.full((max_iterations + 1, len(data)), -1, dtype=np.int32)
        dt = np.zeros((max_iterations + 1, len(data)), dtype=np.float32)
        itp, dtp = _p(it), _p(dt)
    rc = lib().orc_icp(_p(data), len(data), _p(target), len(target), C.byref(prm), C.byref(res), itp, dtp)
    assert rc == 0
    out = {
        "iterations": res.iterations, "nn_passes": res.nn_passes, "n_assoc": res.n_assoc, "mse": res.mse,
        "rigid": np.array(res.rigid[:], dtype=np.float32).reshape(4, 4),
        "cam_rotation": np.array(res.cam_rotation[:], dtype=np.float32).reshape(3, 3),
        "cam_position": np.array(res.cam_position[:], dtype=np.float32),
        "offset": np.array(res.offset[:], dtype=np.float32),
        "pose_R": np.array(res.pose_R[:]).reshape(3, 3), "pose_t": np.array(res.pose_t[:]),
        "small_assoc_exit": res.small_assoc_exit,
    }
    return out, data, it, dt


def _res_dict(res):
    return {
        "iterations": res.iterations, "nn_passes": res.nn_passes, "n_assoc": res.n_assoc, "mse": res.mse,
        "rigid": np.array(res.rigid[:], dtype=np.float32).reshape(4, 4),
        "cam_rotation": np.array(res.cam_rotation[:], dtype=np.float32).reshape(3, 3),
        "cam_position": np.array(res.cam_position[:], dtype=np.float32),
        "offset": np.array(res.offset[:], dtype=np.float32),
        "pose_R": np.array(res.pose_R[:]).reshape(3, 3), "pose_t": np.array(res.pose_t[:]),
        "small_assoc_exit": res.small_assoc_exit,
    }


def icp_carry(data, carry, target, max_iterations=20, threshold=0.0, max_nn_distance=0.75, solve_mode=SOLVE_REFERENCE,
              last_translation=(0, 0, 0), n_threads=1):
    """All-point loop; `carry` follows every motion of `data`.  Returns (result dict, moved data, moved carry)."""
    data = np.ascontiguousarray(data).copy()
    carry = np.ascontiguousarray(carry).copy()
    target = np.ascontiguousarray(target)
    prm = IcpParams(max_iterations, threshold, max_nn_distance, solve_mode, (C.c_float * 3)(*last_translation), n_threads)
    res = IcpResult()
    rc = lib().orc_icp_carry(_p(data), len(data), _p(carry), len(carry), _p(target), len(target), C.byref(prm), C.byref(res))
    assert rc == 0
    return _res_dict(res), data, carry


def icp_keypoints(keypoints, points, map_keypoints, max_iterations=16, threshold=1e-4, max_nn_distance=0.1,
                  solve_mode=SOLVE_REFERENCE, last_translation=(0, 0, 0), n_threads=1):
    """8f-2 (icp.cpp:98,155-258).  Returns (result dict, moved key-points, moved points, non-associations)."""
    kp = np.ascontiguousarray(keypoints).copy()
    pts = np.ascontiguousarray(points).copy()
    mk = np.ascontiguousarray(map_keypoints)
    prm = IcpParams(max_iterations, threshold, max_nn_distance, solve_mode, (C.c_float * 3)(*last_translation), n_threads)
    res = IcpResult()
    non = np.zeros(max(1, (max_iterations + 1) * len(kp)), dtype=POINT_DTYPE)
    n_non = C.c_int(0)
    rc = lib().orc_icp_keypoints(_p(kp), len(kp), _p(pts), len(pts), _p(mk), len(mk), C.byref(prm), C.byref(res),
                                 _p(non), C.byref(n_non))
    assert rc == 0
    return _res_dict(res), kp, pts, non[:n_non.value].copy()


def voxel_coords(p, cell, dims):
    p = np.ascontiguousarray(p, dtype=np.float32).reshape(3)
    d = (C.c_int * 3)(*dims)
    v = (C.c_int * 3)()
    lib().orc_voxel_coords(_p(p), C.c_float(cell), d, v)
    return tuple(v[:])


def map_update_endpoints(grid, dims, cell, pts, rule=RULE_A, delta=25, max_conf=180):
    assert grid.dtype == np.uint8 and grid.flags.c_contiguous
    d = (C.c_int * 3)(*dims)
    pts = np.ascontiguousarray(pts)
    lib().orc_map_update_endpoints(_p(grid), d, C.c_float(cell), _p(pts), len(pts), rule, delta, max_conf)
    return grid


def map_update_tracked(grid, table, dims, cell, pts, variant, delta=25, max_conf=180, map_cloud_size=0):
    """Returns the indices (into pts) of the points appended to the map cloud, in insertion order."""
    assert grid.dtype == np.uint8 and table.dtype == np.int32 and grid.flags.c_contiguous and table.flags.c_contiguous
    d = (C.c_int * 3)(*dims)
    pts = np.ascontiguousarray(pts)
    app = np.zeros(max(len(pts), 1), dtype=np.int32)
    k = lib().orc_map_update_tracked(_p(grid), _p(table), d, C.c_float(cell), _p(pts), len(pts), variant, delta, max_conf,
                                     map_cloud_size, _p(app))
    return app[:k].copy()


def map_integrate_rays(grid, dims, cell, pts, origin, delta_dec=25, delta_inc=25, z_lo=0, z_hi=None):
    assert grid.dtype == np.uint8 and grid.flags.c_contiguous
    d = (C.c_int * 3)(*dims)
    pts = np.ascontiguousarray(pts)
    o = (C.c_float * 3)(*origin)
    if z_hi is None:
        z_hi = dims[2]
    return int(lib().orc_map_integrate_rays(_p(grid), d, C.c_float(cell), _p(pts), len(pts), o, delta_dec,
                                            delta_inc, z_lo, z_hi))


# ---- 8f-4 pose reporting (q = [w, x, y, z]) ------------------------------------------------------------------
def _f(a, n):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    assert a.size == n
    return a


def quat_from_rot(R):
    R = _f(R, 9); q = np.zeros(4, np.float32)
    lib().orc_quat_from_rot(_p(R), _p(q))
    return q


def quat_mul(a, b):
    a = _f(a, 4); b = _f(b, 4); o = np.zeros(4, np.float32)
    lib().orc_quat_mul(_p(a), _p(b), _p(o))
    return o


def quat_inverse(q):
    q = _f(q, 4); o = np.zeros(4, np.float32)
    lib().orc_quat_inverse(_p(q), _p(o))
    return o


def quat_to_euler_deg(q):
    q = _f(q, 4); e = np.zeros(3, np.float32)
    lib().orc_quat_to_euler_deg(_p(q), _p(e))
    return e


def mat_to_euler_deg(R):
    R = _f(R, 9); e = np.zeros(3, np.float32)
    lib().orc_mat_to_euler_deg(_p(R), _p(e))
    return e
