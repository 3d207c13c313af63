"""ctypes binding over the C-ABI of libicpb200.so (include/icpb200.h).

This is harness plumbing for tests/ and bench.py: every compute call goes
through the same extern "C" entry points a C++ host would bind.  There is no
CPU fallback: importing works without a GPU (so the symbol table can be
checked), but any compute call fails loudly when the library or a CUDA
device is missing.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB_PATH = os.environ.get("ICPB_LIB") or os.path.join(_PKG, "lib", "libicpb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "icpb200.h")

POINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("c0", "u1"), ("c1", "u1"), ("c2", "u1"), ("pad", "u1")]
)

OK, ERR_INVALID, ERR_EMPTY, ERR_CUDA, ERR_CAPACITY, ERR_NCCL = 0, 1, 2, 3, 4, 5
COMM_ID_BYTES = 128
SUB_NONE, SUB_STRIDE, SUB_HASH, SUB_STREAM = 0, 1, 2, 3
SOLVE_REFERENCE, SOLVE_KABSCH = 0, 1
RULE_A, RULE_C = 0, 1
NN_BRUTE, NN_GRID, NN_AUTO = 0, 1, 2
FILTER_AUTO, FILTER_DIRECT, FILTER_WARP, FILTER_CENTRED = 0, 1, 2, 3
TRACK_INIT, TRACK_ASSOC, TRACK_NONASSOC = 0, 1, 2
PROF_MAP_RAYS, PROF_MAP_ENDPOINTS, PROF_NN_GRID, PROF_NN_FINALIZE, PROF_LIFT = 0, 1, 2, 3, 4


class IcpbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"icpb status {status}: {message}")
        self.status = status


class Intrinsics(C.Structure):
    _fields_ = [("fx_u", C.c_float), ("cx_u", C.c_float), ("fx_v", C.c_float), ("cx_v", C.c_float),
                ("depth_scale", C.c_float)]


class IcpParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("threshold", C.c_float), ("max_nn_distance", C.c_float),
                ("solve_mode", C.c_int), ("last_translation", C.c_float * 3),
                ("idx_trace", C.c_void_p), ("dist_trace", C.c_void_p), ("nn_mode", C.c_int), ("grid_cell", C.c_float),
                ("nn_filter", C.c_int)]


class IcpResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("nn_passes", C.c_int), ("n_assoc", C.c_int), ("mse", C.c_float),
                ("rigid", C.c_float * 16), ("cam_rotation", C.c_float * 9), ("cam_position", C.c_float * 3),
                ("offset", C.c_float * 3), ("pose_R", C.c_double * 9), ("pose_t", C.c_double * 3),
                ("small_assoc_exit", C.c_int), ("exact_rescans", C.c_int), ("gpu_ms", C.c_float),
                ("kernel_launches", C.c_int), ("nn_partial_ms", C.c_float), ("nn_partial_launches", C.c_int),
                ("nn_qpt", C.c_int), ("nn_splits", C.c_int), ("nn_mode_used", C.c_int), ("grid_cell_used", C.c_float),
                ("nn_filter_used", C.c_int), ("n_nonassoc", C.c_int), ("grid_pairs", C.c_longlong), ("nn_grid_ms", C.c_float)]

    def to_dict(self):
        return {
            "iterations": self.iterations, "nn_passes": self.nn_passes, "n_assoc": self.n_assoc, "mse": self.mse,
            "rigid": np.array(self.rigid[:], dtype=np.float32).reshape(4, 4),
            "cam_rotation": np.array(self.cam_rotation[:], dtype=np.float32).reshape(3, 3),
            "cam_position": np.array(self.cam_position[:], dtype=np.float32),
            "offset": np.array(self.offset[:], dtype=np.float32),
            "pose_R": np.array(self.pose_R[:]).reshape(3, 3), "pose_t": np.array(self.pose_t[:]),
            "small_assoc_exit": self.small_assoc_exit, "exact_rescans": self.exact_rescans,
            "gpu_ms": self.gpu_ms, "kernel_launches": self.kernel_launches,
            "nn_partial_ms": self.nn_partial_ms, "nn_partial_launches": self.nn_partial_launches,
            "nn_qpt": self.nn_qpt, "nn_splits": self.nn_splits, "nn_mode_used": self.nn_mode_used,
            "grid_cell_used": self.grid_cell_used, "nn_filter_used": self.nn_filter_used,
            "n_nonassoc": self.n_nonassoc, "grid_pairs": self.grid_pairs, "nn_grid_ms": self.nn_grid_ms,
        }


_lib = None


def load():
    """Load libicpb200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IcpbError(ERR_CUDA, f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                                  "(make -C icp-slam-prototype_b200); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.icpb_status_string.restype = C.c_char_p
    lib.icpb_last_error.restype = C.c_char_p
    lib.icpb_last_error.argtypes = [C.c_void_p]
    lib.icpb_ctx_stream.restype = C.c_void_p
    lib.icpb_ctx_stream.argtypes = [C.c_void_p]
    lib.icpb_cloud_device_ptr.restype = C.c_void_p
    lib.icpb_cloud_device_ptr.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def reference_intrinsics_v1():
    K = Intrinsics()
    load().icpb_intrinsics_reference_v1(C.byref(K))
    return K


def reference_intrinsics_v2():
    K = Intrinsics()
    load().icpb_intrinsics_reference_v2(C.byref(K))
    return K


def device_count():
    n = C.c_int(0)
    rc = load().icpb_device_count(C.byref(n))
    return n.value if rc == OK else 0


class Context:
    def __init__(self, device=0, stream=None):
        self.lib = load()
        h = C.c_void_p()
        if stream is None:
            rc = self.lib.icpb_ctx_create(int(device), C.byref(h))
        else:
            rc = self.lib.icpb_ctx_create_on_stream(int(device), C.c_void_p(stream), C.byref(h))
        if rc != OK:
            raise IcpbError(rc, self.lib.icpb_last_error(None).decode())
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != OK:
            raise IcpbError(rc, self.lib.icpb_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.icpb_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def sync(self):
        self.check(self.lib.icpb_ctx_sync(self.h))

    def timer_start(self):
        self.check(self.lib.icpb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        self.check(self.lib.icpb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def set_profiling(self, enabled=True):
        self.check(self.lib.icpb_ctx_set_profiling(self.h, int(bool(enabled))))

    def profile_read(self, kernel):
        """(summed ms, launches) of `kernel` (PROF_*) since the last read; profiling mode only."""
        ms = C.c_float(0)
        k = C.c_int(0)
        self.check(self.lib.icpb_ctx_profile_read(self.h, int(kernel), C.byref(ms), C.byref(k)))
        return ms.value, k.value

    def launch_count(self):
        n = C.c_longlong(0)
        self.check(self.lib.icpb_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def measure_fp32_peak(self, repeats=5):
        tf = C.c_double(0)
        ms = C.c_float(0)
        self.check(self.lib.icpb_measure_fp32_peak(self.h, repeats, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    # ---- clouds
    def cloud(self, capacity):
        return Cloud(self, capacity)

    def cloud_from_points(self, pts, capacity=None):
        c = Cloud(self, capacity or max(len(pts), 1))
        c.upload(pts)
        return c

    def backproject_batch_device(self, d_depth, d_bgr, frames, w, h, K, d_points, capacity_per_frame, d_counts):
        self.check(self.lib.icpb_backproject_batch_device(
            self.h, C.c_void_p(d_depth), C.c_void_p(d_bgr) if d_bgr else None, int(frames), int(w), int(h), C.byref(K),
            C.c_void_p(d_points), int(capacity_per_frame), C.c_void_p(d_counts)))

    # ---- image stages
    def frame_lift_band_device(self, d_depth, w, h, row0, row1, K, R, t, d_band, band_capacity):
        """Sync-free: rows [row0, row1) of a device-resident depth frame -> world-space band (header row + points)."""
        Rp = None if R is None else (C.c_float * 9)(*np.asarray(R, np.float32).reshape(9))
        tp = None if t is None else (C.c_float * 3)(*np.asarray(t, np.float32).reshape(3))
        self.check(self.lib.icpb_frame_lift_band_device(self.h, C.c_void_p(d_depth), int(w), int(h), int(row0), int(row1),
                                                        C.byref(K), Rp, tp, C.c_void_p(d_band), int(band_capacity)))

    def normals_batch_device(self, d_depth, frames, w, h, d_normals):
        self.check(self.lib.icpb_normals_batch_device(self.h, C.c_void_p(d_depth), int(frames), int(w), int(h),
                                                      C.c_void_p(d_normals)))

    def normals(self, depth):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        h, w = depth.shape
        out = np.zeros((h, w, 3), dtype=np.float32)
        self.check(self.lib.icpb_normals_from_depth(self.h, _p(depth), w, h, _p(out)))
        return out

    def depth_filter(self, depth, min_d=1000, max_d=25000):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        h, w = depth.shape
        out = np.zeros((h, w), dtype=np.uint16)
        self.check(self.lib.icpb_depth_filter(self.h, _p(depth), w, h, int(min_d), int(max_d), _p(out)))
        return out

    # ---- registration
    def nn_search(self, data, target):
        idx = np.zeros(data.n, dtype=np.int32)
        dist = np.zeros(data.n, dtype=np.float32)
        resc = C.c_int(0)
        self.check(self.lib.icpb_nn_search(self.h, data.h, target.h, _p(idx), _p(dist), C.byref(resc)))
        return idx, dist, resc.value

    def icp_register(self, data, target, max_iterations=20, threshold=0.0, max_nn_distance=0.75,
                     solve_mode=SOLVE_REFERENCE, last_translation=(0, 0, 0), trace=False, nn_mode=NN_BRUTE,
                     grid_cell=0.0, nn_filter=FILTER_AUTO):
        it = dt = None
        prm = IcpParams(max_iterations, threshold, max_nn_distance, solve_mode, (C.c_float * 3)(*last_translation),
                        None, None, nn_mode, grid_cell, nn_filter)
        if trace:
            it = np.full((max_iterations + 1, data.n), -1, dtype=np.int32)
            dt = np.zeros((max_iterations + 1, data.n), dtype=np.float32)
            prm.idx_trace = it.ctypes.data
            prm.dist_trace = dt.ctypes.data
        res = IcpResult()
        self.check(self.lib.icpb_icp_register(self.h, data.h, target.h, C.byref(prm), C.byref(res)))
        return res.to_dict(), it, dt

    def icp_register_carry(self, data, target, carry, max_iterations=20, threshold=0.0, max_nn_distance=0.75,
                           solve_mode=SOLVE_REFERENCE, nn_mode=NN_BRUTE):
        """icpb_icp_register_carry: the all-point loop; `carry` receives every motion `data` receives."""
        prm = IcpParams(max_iterations, threshold, max_nn_distance, solve_mode, (C.c_float * 3)(0, 0, 0), None, None,
                        nn_mode, 0.0, FILTER_AUTO)
        res = IcpResult()
        self.check(self.lib.icpb_icp_register_carry(self.h, data.h, target.h, carry.h if carry is not None else None,
                                                    C.byref(prm), C.byref(res)))
        return res.to_dict()

    def icp_register_keypoints(self, keypoints, points, map_keypoints, max_iterations=16, threshold=1e-4,
                               max_nn_distance=0.1, solve_mode=SOLVE_REFERENCE, last_translation=(0, 0, 0),
                               non_associations=None):
        """8f-2: the reference's live loop (icp.cpp:98,155-258).  Clouds are moved in place."""
        prm = IcpParams(max_iterations, threshold, max_nn_distance, solve_mode, (C.c_float * 3)(*last_translation),
                        None, None, NN_BRUTE, 0.0, FILTER_AUTO)
        res = IcpResult()
        self.check(self.lib.icpb_icp_register_keypoints(
            self.h, keypoints.h, points.h if points is not None else None, map_keypoints.h, C.byref(prm), C.byref(res),
            non_associations.h if non_associations is not None else None))
        return res.to_dict()

    def icp_register_batch(self, datas, targets, max_iterations=20, threshold=0.0, max_nn_distance=0.75,
                           solve_mode=SOLVE_REFERENCE, nn_mode=NN_BRUTE, grid_cell=0.0):
        n = len(datas)
        prm = IcpParams(max_iterations, threshold, max_nn_distance, solve_mode, (C.c_float * 3)(0, 0, 0), None, None,
                        nn_mode, grid_cell, FILTER_AUTO)
        dh = (C.c_void_p * n)(*[d.h for d in datas])
        th = (C.c_void_p * n)(*[t.h for t in targets])
        res = (IcpResult * n)()
        self.check(self.lib.icpb_icp_register_batch(self.h, dh, th, n, C.byref(prm), res))
        return [r.to_dict() for r in res]

    def icp_register_async(self, data, target, max_iterations=20, threshold=0.0, max_nn_distance=0.75,
                           solve_mode=SOLVE_REFERENCE, nn_mode=NN_BRUTE, grid_cell=0.0):
        """icpb_icp_register_async: returns a Pending as soon as the loop is enqueued; .wait() -> result dict."""
        prm = IcpParams(max_iterations, threshold, max_nn_distance, solve_mode, (C.c_float * 3)(0, 0, 0), None, None,
                        nn_mode, grid_cell, FILTER_AUTO)
        h = C.c_void_p()
        self.check(self.lib.icpb_icp_register_async(self.h, data.h, target.h, C.byref(prm), C.byref(h)))
        return Pending(self, h, 1, (data, target))

    def icp_register_batch_async(self, datas, targets, max_iterations=20, threshold=0.0, max_nn_distance=0.75,
                                 solve_mode=SOLVE_REFERENCE, nn_mode=NN_BRUTE, grid_cell=0.0):
        n = len(datas)
        prm = IcpParams(max_iterations, threshold, max_nn_distance, solve_mode, (C.c_float * 3)(0, 0, 0), None, None,
                        nn_mode, grid_cell, FILTER_AUTO)
        dh = (C.c_void_p * n)(*[d.h for d in datas])
        th = (C.c_void_p * n)(*[t.h for t in targets])
        h = C.c_void_p()
        self.check(self.lib.icpb_icp_register_batch_async(self.h, dh, th, n, C.byref(prm), C.byref(h)))
        return Pending(self, h, n, (list(datas), list(targets)))

    # ---- map
    def map(self, dims, cell, z_lo=0, z_hi=None):
        return Map(self, dims, cell, z_lo, dims[2] if z_hi is None else z_hi)


class Cloud:
    def __init__(self, ctx, capacity):
        self.ctx = ctx
        h = C.c_void_p()
        ctx.check(ctx.lib.icpb_cloud_create(ctx.h, int(capacity), C.byref(h)))
        self.h = h
        self.capacity = capacity

    def close(self):
        if self.h:
            self.ctx.lib.icpb_cloud_destroy(self.h)
            self.h = None

    @property
    def n(self):
        n = C.c_int(0)
        self.ctx.check(self.ctx.lib.icpb_cloud_size(self.h, C.byref(n)))
        return n.value

    def upload(self, pts):
        pts = np.ascontiguousarray(pts)
        assert pts.dtype == POINT_DTYPE
        self.ctx.check(self.ctx.lib.icpb_cloud_upload(self.h, _p(pts), len(pts)))

    def upload_xyz(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self.ctx.check(self.ctx.lib.icpb_cloud_upload_xyz(self.h, _p(xyz), len(xyz)))

    def upload_device(self, ptr, n):
        self.ctx.check(self.ctx.lib.icpb_cloud_upload_device(self.h, C.c_void_p(ptr), int(n)))

    def download_device(self, ptr, capacity):
        self.ctx.check(self.ctx.lib.icpb_cloud_download_device(self.h, C.c_void_p(ptr), int(capacity)))

    def pack_band_device(self, ptr, band_capacity):
        self.ctx.check(self.ctx.lib.icpb_cloud_pack_band_device(self.h, C.c_void_p(ptr), int(band_capacity)))

    def assemble_bands_device(self, ptr, world, band_capacity):
        self.ctx.check(self.ctx.lib.icpb_cloud_assemble_bands_device(self.h, C.c_void_p(ptr), int(world), int(band_capacity)))
        return self.n

    def device_ptr(self):
        return self.ctx.lib.icpb_cloud_device_ptr(self.h)

    def download(self):
        n = self.n
        out = np.zeros(max(n, 1), dtype=POINT_DTYPE)
        nn = C.c_int(0)
        self.ctx.check(self.ctx.lib.icpb_cloud_download(self.h, _p(out), len(out), C.byref(nn)))
        return out[:n]

    def copy_from(self, other):
        self.ctx.check(self.ctx.lib.icpb_cloud_copy(self.h, other.h))

    def from_depth(self, depth, bgr=None, K=None, rule=SUB_NONE, rule_arg=1, seed=0, keep_stream=None):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        h, w = depth.shape
        K = K or reference_intrinsics_v1()
        if bgr is not None:
            bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        ks_len = 0
        if keep_stream is not None:
            keep_stream = np.ascontiguousarray(keep_stream, dtype=np.uint8)
            ks_len = len(keep_stream)
        self.ctx.check(self.ctx.lib.icpb_cloud_from_depth(self.h, _p(depth), _p(bgr), w, h, C.byref(K), rule,
                                                         C.c_uint32(rule_arg), C.c_uint32(seed), _p(keep_stream),
                                                         ks_len))
        return self.n

    def from_depth_device(self, d_depth, w, h, d_bgr=None, K=None, rule=SUB_NONE, rule_arg=1, seed=0):
        K = K or reference_intrinsics_v1()
        self.ctx.check(self.ctx.lib.icpb_cloud_from_depth_device(
            self.h, C.c_void_p(d_depth), C.c_void_p(d_bgr) if d_bgr else None, w, h, C.byref(K), rule,
            C.c_uint32(rule_arg), C.c_uint32(seed), None, 0))
        return self.n

    def transform(self, R=None, t=None):
        Rp = None if R is None else np.ascontiguousarray(R, dtype=np.float32).reshape(9)
        tp = None if t is None else np.ascontiguousarray(t, dtype=np.float32).reshape(3)
        self.ctx.check(self.ctx.lib.icpb_cloud_transform(self.h, _p(Rp), _p(tp)))

    def center(self):
        c = (C.c_double * 3)()
        self.ctx.check(self.ctx.lib.icpb_cloud_center(self.h, c))
        return np.array(c[:])



class Pending:
    """A registration in flight (icpb_pending).  Keeps its clouds alive until the wait."""

    def __init__(self, ctx, h, count, keep):
        self.ctx, self.h, self.count, self._keep = ctx, h, count, keep

    def ready(self):
        r = C.c_int(0)
        self.ctx.check(self.ctx.lib.icpb_icp_pending_ready(self.h, C.byref(r)))
        return bool(r.value)

    def wait(self):
        if self.h is None:
            raise RuntimeError("already waited for")
        res = (IcpResult * self.count)()
        h, self.h = self.h, None
        self.ctx.check(self.ctx.lib.icpb_icp_pending_wait(h, res))
        self._keep = None
        out = [r.to_dict() for r in res]
        return out[0] if self.count == 1 else out

class Map:
    def __init__(self, ctx, dims, cell, z_lo, z_hi):
        self.ctx = ctx
        self.dims = tuple(int(d) for d in dims)
        self.z_lo, self.z_hi = int(z_lo), int(z_hi)
        h = C.c_void_p()
        d = (C.c_int * 3)(*self.dims)
        ctx.check(ctx.lib.icpb_map_create(ctx.h, d, C.c_float(cell), self.z_lo, self.z_hi, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.ctx.lib.icpb_map_destroy(self.h)
            self.h = None

    def clear(self):
        self.ctx.check(self.ctx.lib.icpb_map_clear(self.h))

    def update_endpoints(self, cloud, rule=RULE_A, delta=25, max_conf=180):
        self.ctx.check(self.ctx.lib.icpb_map_update_endpoints(self.h, cloud.h, rule, delta, max_conf))

    def update_tracked(self, cloud, variant, delta, max_conf, map_cloud):
        n = C.c_int(0)
        self.ctx.check(self.ctx.lib.icpb_map_update_tracked(self.h, cloud.h, variant, delta, max_conf, map_cloud.h, C.byref(n)))
        return n.value

    def has_entry(self, p):
        pp = (C.c_float * 3)(*p)
        e = C.c_int(0)
        self.ctx.check(self.ctx.lib.icpb_map_has_entry(self.h, pp, C.byref(e)))
        return bool(e.value)

    def integrate_rays(self, cloud, origin, delta_dec=25, delta_inc=25, count_visits=True):
        o = (C.c_float * 3)(*origin)
        v = C.c_longlong(0)
        self.ctx.check(self.ctx.lib.icpb_map_integrate_rays(self.h, cloud.h, o, delta_dec, delta_inc,
                                                           C.byref(v) if count_visits else None))
        return v.value

    def integrate_rays_profiled(self, cloud, origin, layer_work, delta_dec=25, delta_inc=25):
        """Integration + per-z-layer work histogram of the ray walk, accumulated into layer_work (uint64[dims[2]])."""
        o = (C.c_float * 3)(*origin)
        assert layer_work.dtype == np.uint64 and len(layer_work) == self.dims[2]
        self.ctx.check(self.ctx.lib.icpb_map_integrate_rays_profiled(self.h, cloud.h, o, delta_dec, delta_inc, _p(layer_work)))

    def integrate_bands_device(self, d_bands, world, band_capacity, origin, delta_dec=25, delta_inc=25):
        """Sync-free: `world` device-resident bands -> ray decrements + endpoint increments on this slab."""
        o = (C.c_float * 3)(*origin)
        self.ctx.check(self.ctx.lib.icpb_map_integrate_bands_device(self.h, C.c_void_p(d_bands), int(world),
                                                                    int(band_capacity), o, delta_dec, delta_inc))

    def voxel_coords(self, p):
        pp = (C.c_float * 3)(*p)
        v = (C.c_int * 3)()
        self.ctx.check(self.ctx.lib.icpb_map_voxel_coords(self.h, pp, v))
        return tuple(v[:])

    def size_bytes(self):
        s = C.c_longlong(0)
        self.ctx.check(self.ctx.lib.icpb_map_size_bytes(self.h, C.byref(s)))
        return s.value

    def download(self):
        """Returns the slab as [X, Y, z_hi - z_lo] uint8 (reference order world[x][y][z])."""
        out = np.zeros((self.dims[0], self.dims[1], self.z_hi - self.z_lo), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.icpb_map_download(self.h, _p(out), out.size))
        return out

    def upload(self, grid):
        grid = np.ascontiguousarray(grid, dtype=np.uint8)
        self.ctx.check(self.ctx.lib.icpb_map_upload(self.h, _p(grid), grid.size))


# ---- multi-GPU (include/icpb200.h "multi-GPU"): NCCL communicator + z-slab map, all inside the library ---------------
def comm_unique_id():
    """128 bytes from ncclGetUniqueId (call on rank 0, hand to every rank by any transport)."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    rc = load().icpb_comm_unique_id(buf)
    if rc != OK:
        raise IcpbError(rc, load().icpb_last_error(None).decode())
    return bytes(buf)


def slab_bounds_from_work(work, world):
    """z-slab boundaries [0, ..., layers] with equal shares of the per-layer work histogram."""
    work = np.ascontiguousarray(work, dtype=np.uint64)
    b = (C.c_int * (world + 1))()
    rc = load().icpb_slab_bounds_from_work(_p(work), len(work), int(world), b)
    if rc != OK:
        raise IcpbError(rc, "icpb_slab_bounds_from_work")
    return list(b)


class Comm:
    def __init__(self, ctx, world, rank, unique_id):
        self.ctx, self.world, self.rank = ctx, int(world), int(rank)
        h = C.c_void_p()
        idb = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        ctx.check(ctx.lib.icpb_comm_create(ctx.h, self.world, self.rank, idb, C.byref(h)))
        self.h = h

    @staticmethod
    def from_torch(ctx):
        """Bootstrap over an initialised torch.distributed group: rank 0's unique id is broadcast as an object."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return Comm(ctx, world, rank, box[0])

    def close(self):
        if self.h:
            self.ctx.lib.icpb_comm_destroy(self.h)
            self.h = None

    def shard_range(self, n_items):
        lo, hi = C.c_longlong(0), C.c_longlong(0)
        self.ctx.check(self.ctx.lib.icpb_comm_shard_range(self.h, C.c_longlong(n_items), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def allgather_host(self, arr):
        """arr: contiguous numpy array, same shape on every rank -> [world, *arr.shape]."""
        arr = np.ascontiguousarray(arr)
        out = np.zeros((self.world,) + arr.shape, dtype=arr.dtype)
        self.ctx.check(self.ctx.lib.icpb_comm_allgather_host(self.h, _p(arr), _p(out), C.c_longlong(arr.nbytes)))
        return out


class SlabMapC:
    """icpb_slabmap: the rank's z-slab of a shared certainty map; comm=None for a single rank."""

    def __init__(self, ctx, comm, dims, cell, w, h, bounds=None):
        self.ctx, self.comm = ctx, comm
        self.dims = tuple(int(d) for d in dims)
        hnd = C.c_void_p()
        d = (C.c_int * 3)(*self.dims)
        b = None if bounds is None else (C.c_int * len(bounds))(*[int(x) for x in bounds])
        ctx.check(ctx.lib.icpb_slabmap_create(ctx.h, comm.h if comm is not None else None, d, C.c_float(cell), b, int(w),
                                              int(h), C.byref(hnd)))
        self.h = hnd
        mh, lo, hi = C.c_void_p(), C.c_int(0), C.c_int(0)
        ctx.check(ctx.lib.icpb_slabmap_local(self.h, C.byref(mh), C.byref(lo), C.byref(hi)))
        self.z_lo, self.z_hi = lo.value, hi.value
        self.map = Map.__new__(Map)   # a view of the slab's map handle (owned by the slab map)
        self.map.ctx, self.map.dims, self.map.z_lo, self.map.z_hi, self.map.h = ctx, self.dims, lo.value, hi.value, mh

    def integrate_sequence_device(self, d_depths, frames, K, Rs, ts, delta_dec=25, delta_inc=25, frames_per_exchange=8):
        Rs = np.ascontiguousarray(Rs, dtype=np.float32).reshape(frames, 9)
        ts = np.ascontiguousarray(ts, dtype=np.float32).reshape(frames, 3)
        self.ctx.check(self.ctx.lib.icpb_slabmap_integrate_sequence_device(
            self.h, C.c_void_p(d_depths), int(frames), C.byref(K), _p(Rs), _p(ts), int(delta_dec), int(delta_inc),
            int(frames_per_exchange)))

    def download(self):
        return self.map.download()

    def close(self):
        if self.h:
            self.map.h = None
            self.ctx.lib.icpb_slabmap_destroy(self.h)
            self.h = None


# ---- pose reporting: scalar host functions of the library; q = [w, x, y, z], degrees -----------
def _f32(a, n):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    if a.size != n:
        raise ValueError(f"expected {n} floats")
    return a


def pose_quat_from_rotation(R):
    R = _f32(R, 9); q = np.zeros(4, np.float32)
    if load().icpb_pose_quat_from_rotation(_p(R), _p(q)):
        raise IcpbError(ERR_INVALID, "icpb_pose_quat_from_rotation")
    return q


def pose_quat_mul(a, b):
    a = _f32(a, 4); b = _f32(b, 4); o = np.zeros(4, np.float32)
    if load().icpb_pose_quat_mul(_p(a), _p(b), _p(o)):
        raise IcpbError(ERR_INVALID, "icpb_pose_quat_mul")
    return o


def pose_quat_inverse(q):
    q = _f32(q, 4); o = np.zeros(4, np.float32)
    if load().icpb_pose_quat_inverse(_p(q), _p(o)):
        raise IcpbError(ERR_INVALID, "icpb_pose_quat_inverse")
    return o


def pose_quat_to_euler_deg(q):
    q = _f32(q, 4); e = np.zeros(3, np.float32)
    if load().icpb_pose_quat_to_euler_deg(_p(q), _p(e)):
        raise IcpbError(ERR_INVALID, "icpb_pose_quat_to_euler_deg")
    return e


def pose_matrix_to_euler_deg(R):
    R = _f32(R, 9); e = np.zeros(3, np.float32)
    if load().icpb_pose_matrix_to_euler_deg(_p(R), _p(e)):
        raise IcpbError(ERR_INVALID, "icpb_pose_matrix_to_euler_deg")
    return e
