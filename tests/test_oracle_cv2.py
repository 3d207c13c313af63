"""Pin the oracle's restatement of the OpenCV arithmetic that is NOT under /root/reference
(OpenCV 3.2 world lib, build/SLAM.exe.vcxproj:154,183) against cv2 4.13 present in this image,
and the Kabsch mode against the reference's own rigid_transform_3D.py when it is readable."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

REF_PY = "/root/reference/rigid_transform_3D.py"


def test_gemm33f_bit_exact_vs_cv2(orc):
    rng = np.random.default_rng(0)
    for _ in range(500):
        A = rng.standard_normal((3, 3)).astype(np.float32)
        B = rng.standard_normal((3, 3)).astype(np.float32)
        assert np.array_equal(orc.gemm33f(A, B), cv2.gemm(A, B, 1.0, None, 0.0))


def test_rotate_bit_exact_vs_cv2_gemm(orc):
    """PointCloud::rotate = R * M^T through cv::Mat operator* (pointcloud.cpp:323-325)."""
    rng = np.random.default_rng(1)
    R = rng.standard_normal((3, 3)).astype(np.float32)
    xyz = (rng.standard_normal((4000, 3)) * 4 + 5).astype(np.float32)
    want = cv2.gemm(R, np.ascontiguousarray(xyz.T), 1.0, None, 0.0).T
    got = orc.xyz_of(orc.rotate(orc.make_points(xyz), R))
    assert np.array_equal(got, want)


def test_inv_and_det_bit_exact_vs_cv2(orc):
    rng = np.random.default_rng(2)
    for k in range(500):
        A = rng.standard_normal((3, 3)).astype(np.float32)
        if k % 2 == 0:
            A = np.linalg.qr(A.astype(np.float64))[0].astype(np.float32)
        assert orc.det33f(A) == cv2.determinant(A)
        assert np.array_equal(orc.inv33f(A), cv2.invert(A)[1])


def test_svd3_matches_lapack_and_cv2(orc):
    rng = np.random.default_rng(3)
    for _ in range(300):
        A = rng.standard_normal((3, 3)) * rng.uniform(0.1, 100)
        U, w, Vt = orc.svd3(A)
        assert np.allclose(U @ np.diag(w) @ Vt, A, atol=1e-12 * np.abs(A).max())
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-13) and np.allclose(Vt @ Vt.T, np.eye(3), atol=1e-13)
        assert np.all(np.diff(w) <= 0)
        assert np.allclose(w, np.linalg.svd(A, compute_uv=False), rtol=1e-12)
        # R = V U^T is what the solve consumes (icp.cpp:218); compare with cv2's float32 SVD
        w32, u32, vt32 = cv2.SVDecomp(A.astype(np.float32))
        assert np.abs(Vt.T @ U.T - (vt32.T @ u32.T)).max() < 5e-5


def _cv2_reference_solve(a, b):
    """icp.cpp:199-246 with cv2 standing in for OpenCV 3.2 (float32 throughout)."""
    a = a.astype(np.float32); b = b.astype(np.float32)
    M = cv2.gemm(b, a, 1.0, None, 0.0, flags=cv2.GEMM_1_T)
    w, u, vt = cv2.SVDecomp(M)
    R = cv2.gemm(vt, u, 1.0, None, 0.0, flags=cv2.GEMM_1_T | cv2.GEMM_2_T)
    if cv2.determinant(R) < 0:
        R[:, 2] *= -1
    Rinv = cv2.invert(R)[1]
    offset = np.zeros(3, np.float32)
    for k in range(len(a)):
        offset += a[k] - b[k]
    offset /= np.float32(len(a))
    return R, Rinv, offset


def test_reference_solve_first_iteration_vs_cv2(orc):
    """One solve of the oracle (max_iterations=1) against the cv2-flavoured solve: pose within 1e-5."""
    rng = np.random.default_rng(4)
    from icpb200 import synth
    base = rng.uniform(3, 7, (2000, 3))
    Rt = synth.rot_axis_angle([0.2, 1.0, -0.4], np.deg2rad(1.5))
    data = orc.make_points((base @ Rt.T) + [0.01, -0.02, 0.005])
    target = orc.make_points(base)
    res, out, it, dt = orc.icp(data, target, 1, 0.0, 0.75, orc.SOLVE_REFERENCE, trace=True)
    idx = it[0]
    R, Rinv, offset = _cv2_reference_solve(orc.xyz_of(data), orc.xyz_of(target)[idx])
    assert np.abs(res["rigid"][:3, :3] - R).max() < 1e-5
    assert np.abs(res["cam_rotation"] - Rinv).max() < 1e-5
    assert np.abs(res["offset"] - offset).max() < 1e-5


@pytest.mark.skipif(not os.path.exists(REF_PY), reason="reference checkout not present")
def test_kabsch_mode_vs_rigid_transform_3D_py(orc):
    """Run the reference's own rigid_transform_3D.py (numpy>=2 needs the `mat` alias it star-imports)."""
    import contextlib, io
    src = open(REF_PY).read().split("# Test with random data")[0]
    ns = {"mat": np.asmatrix}
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(src, REF_PY, "exec"), ns)
    rng = np.random.default_rng(5)
    from icpb200 import synth
    base = rng.uniform(4, 6, (1500, 3))
    Rt = synth.rot_axis_angle([1, 1, 0.2], np.deg2rad(2.0))
    moved = (base - 5) @ Rt.T + 5 + [0.02, 0.0, -0.01]
    data, target = orc.make_points(base), orc.make_points(moved)
    res, out, it, dt = orc.icp(data, target, 1, 0.0, 0.75, orc.SOLVE_KABSCH, trace=True)
    A = np.asmatrix(orc.xyz_of(data).astype(np.float64))
    B = np.asmatrix(orc.xyz_of(target)[it[0]].astype(np.float64))
    with contextlib.redirect_stdout(io.StringIO()):
        R, t = ns["rigid_transform_3D"](A, B)
    assert np.abs(res["pose_R"] - np.asarray(R)).max() < 1e-6
    assert np.abs(res["pose_t"] - np.asarray(t).ravel()).max() < 1e-5


def test_depth_filter_vs_cv2_morphology(orc):
    """filterDepthImage (SLAM.cpp:553-573): threshold, then 5x5 rect dilate + erode anchored at (3,3)."""
    rng = np.random.default_rng(6)
    for (h, w) in [(48, 64), (7, 9), (424, 512)]:
        img = rng.integers(0, 30000, (h, w)).astype(np.uint16)
        img[rng.random((h, w)) < 0.3] = 0
        thr = img.copy()
        thr[(thr > 25000) | (thr < 1000)] = 0
        el = cv2.getStructuringElement(cv2.MORPH_RECT, (5, 5), (3, 3))
        want = cv2.erode(cv2.dilate(thr, el, anchor=(3, 3)), el, anchor=(3, 3))
        assert np.array_equal(orc.depth_filter(img, 1000, 25000), want)


def test_normals_match_float64_formula(orc):
    """getNormalMap (SLAM.cpp:412-430) interior pixels; borders are the defined zeros."""
    rng = np.random.default_rng(7)
    d = rng.integers(0, 20000, (30, 40)).astype(np.uint16)
    n = orc.normals(d)
    f = d.astype(np.float64)
    dzdx = (f[2:, 1:-1] - f[:-2, 1:-1]) / 2
    dzdy = (f[1:-1, 2:] - f[1:-1, :-2]) / 2
    v = np.stack([-dzdx, -dzdy, np.ones_like(dzdx)], -1)
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    assert np.abs(n[1:-1, 1:-1] - v).max() < 1e-7
    assert not n[0].any() and not n[-1].any() and not n[:, 0].any() and not n[:, -1].any()


def test_normals_bit_exact_against_the_cv_normalize_formula(orc):
    """cv::normalize(Vec3f) as OpenCV 3.2 defines it (matx.hpp: `double nv = norm(v); return v * (nv ? 1./nv : 0.)`,
    norm = sqrt of the squares summed left to right in double, the product rounded to float per component), written out
    in IEEE double with numpy: the oracle's normals must equal it bit for bit, not just to 1e-7 (SLAM.cpp:412-430)."""
    rng = np.random.default_rng(11)
    for (h, w) in [(30, 40), (64, 48), (5, 4)]:
        d = rng.integers(0, 65536, (h, w)).astype(np.uint16)
        d[rng.random((h, w)) < 0.2] = 0
        n = orc.normals(d)
        f = d.astype(np.float32)
        dzdx = (f[2:, 1:-1] - f[:-2, 1:-1]) / np.float32(2)          # float arithmetic, :421-422
        dzdy = (f[1:-1, 2:] - f[1:-1, :-2]) / np.float32(2)
        v = np.stack([-dzdx, -dzdy, np.ones_like(dzdx)], -1).astype(np.float64)
        nv = np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2])
        want = (v * (1.0 / nv)[..., None]).astype(np.float32)
        assert np.array_equal(n[1:-1, 1:-1].view(np.uint32), want.view(np.uint32))
