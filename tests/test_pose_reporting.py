"""SURVEY.md 8f-4 -- pose reporting (rotation matrix -> quaternion -> Euler degrees; SLAM.cpp:284-293).

Scalar per-frame host code in the reference and here: the library's icpb_pose_* entry points need no device, so this
whole file runs on the CPU.  Pins: the oracle against the reference's OWN Quaternion class (quaternion.cpp compiled by
path, Route B) bit for bit; the oracle's Euler formulas (SLAM.cpp is not compilable here) against scipy; the library
against the oracle bit for bit."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import icpb200
from oracle import ref


def _rotations(n, seed):
    R = Rotation.random(n, random_state=seed).as_matrix().astype(np.float32)
    extra = [np.eye(3, dtype=np.float32)]
    for ax in range(3):                      # 180-degree turns: w == 0, one of x / y / z dominates
        d = -np.ones(3, np.float32); d[ax] = 1
        extra.append(np.diag(d))
    for deg in (90, 120, 179.99, -90):       # each of the four "largest component" branches near its boundary
        extra.append(Rotation.from_euler("xyz", [deg, 20, -35], degrees=True).as_matrix().astype(np.float32))
        extra.append(Rotation.from_euler("zyx", [deg, -50, 10], degrees=True).as_matrix().astype(np.float32))
    return np.concatenate([R, np.stack(extra)], 0)


def _bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_quaternion_matches_the_references_own_class(orc):
    Rs = _rotations(2000, 1)
    for R in Rs:
        assert np.array_equal(_bits(orc.quat_from_rot(R)), _bits(ref.quat_from_rot(R)))
    rng = np.random.default_rng(2)
    for _ in range(2000):
        a, b = rng.standard_normal(4).astype(np.float32), rng.standard_normal(4).astype(np.float32)
        assert np.array_equal(_bits(orc.quat_mul(a, b)), _bits(ref.quat_mul(a, b)))
        assert np.array_equal(_bits(orc.quat_inverse(a)), _bits(ref.quat_inverse(a)))


def test_oracle_quaternion_is_the_rotation(orc):
    """Known answer: the quaternion represents the same rotation as the matrix (sign aside)."""
    for R in _rotations(500, 3):
        q = orc.quat_from_rot(R).astype(np.float64)        # w, x, y, z
        Rq = Rotation.from_quat([q[1], q[2], q[3], q[0]]).as_matrix()
        assert np.abs(Rq - R).max() < 5e-4   # the diagonal-sqrt construction loses digits near 180 degrees
        assert abs(np.linalg.norm(q) - 1) < 1e-6


def test_oracle_euler_formulas_against_scipy(orc):
    """toEulerianAngle is the roll / pitch / yaw (intrinsic z-y'-x'') extraction; transformationMatToEulerianAngle
    reads the same convention off the TRANSPOSED matrix (it uses t(1,2), t(0,2), t(0,1))."""
    for R in _rotations(500, 4):
        q = orc.quat_from_rot(R)
        e = orc.quat_to_euler_deg(q).astype(np.float64)
        qd = q.astype(np.float64)                          # the formula is pinned on the SAME (float) quaternion
        want = Rotation.from_quat([qd[1], qd[2], qd[3], qd[0]]).as_euler("ZYX", degrees=True)[::-1]  # -> x, y, z
        if abs(abs(want[1]) - 90) < 0.5:
            continue                                                                              # gimbal lock
        d = (e - want + 180) % 360 - 180
        assert np.abs(d).max() < 2e-3, (e, want)
        em = orc.mat_to_euler_deg(R).astype(np.float64)
        want_t = Rotation.from_matrix(R.astype(np.float64).T).as_euler("ZYX", degrees=True)[::-1]
        if abs(abs(want_t[1]) - 90) < 0.5:
            continue
        d = (em - want_t + 180) % 360 - 180
        assert np.abs(d).max() < 2e-3, (em, want_t)


def test_library_pose_functions_bit_equal_to_the_oracle(orc):
    rng = np.random.default_rng(5)
    for R in _rotations(2000, 6):
        q = icpb200.pose_quat_from_rotation(R)
        assert np.array_equal(_bits(q), _bits(orc.quat_from_rot(R)))
        assert np.array_equal(_bits(icpb200.pose_quat_to_euler_deg(q)), _bits(orc.quat_to_euler_deg(q)))
        assert np.array_equal(_bits(icpb200.pose_matrix_to_euler_deg(R)), _bits(orc.mat_to_euler_deg(R)))
        b = rng.standard_normal(4).astype(np.float32)
        assert np.array_equal(_bits(icpb200.pose_quat_mul(q, b)), _bits(orc.quat_mul(q, b)))
        assert np.array_equal(_bits(icpb200.pose_quat_inverse(b)), _bits(orc.quat_inverse(b)))


def test_reporting_chain_of_the_frame_loop(orc):
    """SLAM.cpp:284-293: rotation accumulates icpRotation per frame; the printed angles follow from Quaternion(rotation)."""
    rot = np.eye(3, dtype=np.float32)
    step = Rotation.from_euler("xyz", [0.4, -0.7, 0.2], degrees=True).as_matrix().astype(np.float32)
    for f in range(1, 40):
        rot = (rot @ step).astype(np.float32)
        e = icpb200.pose_quat_to_euler_deg(icpb200.pose_quat_from_rotation(rot))
        want = Rotation.from_matrix(rot.astype(np.float64)).as_euler("ZYX", degrees=True)[::-1]
        assert np.abs(((e - want + 180) % 360) - 180).max() < 5e-3
