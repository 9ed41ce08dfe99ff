"""Wire codec (include/scl_wire.h): the pose arithmetic of distributedMapping.h:1129-1158 / 1244-1259 and the
message payloads. PCL / GTSAM / tf are absent from /root/reference (parity unpinned): checked against an independent
scipy restatement and algebraic identities. Host-only functions of the C-ABI library: no GPU needed."""
import ctypes as C

import numpy as np
from scipy.spatial.transform import Rotation as Rot

from scl_slam_b200 import engine


def _mat(p):   # (x, y, z, roll, pitch, yaw) -> 4x4, Rz(yaw) Ry(pitch) Rx(roll)
    T = np.eye(4)
    T[:3, :3] = Rot.from_euler("ZYX", [p[5], p[4], p[3]]).as_matrix()
    T[:3, 3] = p[:3]
    return T


def _same_rotation(q1, q2, tol):
    q1, q2 = np.asarray(q1), np.asarray(q2)
    return min(np.abs(q1 - q2).max(), np.abs(q1 + q2).max()) < tol


def test_pose6_to_transform_is_tf_setrpy():
    rng = np.random.default_rng(1)
    for _ in range(50):
        p = np.concatenate([rng.normal(0, 20, 3), rng.uniform(-np.pi, np.pi, 3)]).astype(np.float32)
        t = engine.wire_pose6_to_transform(p)
        assert np.allclose(t[:3], p[:3].astype(np.float64))
        q = Rot.from_euler("ZYX", [float(p[5]), float(p[4]), float(p[3])]).as_quat()      # x y z w
        assert _same_rotation(t[3:], q, 1e-12) and abs(np.linalg.norm(t[3:]) - 1.0) < 1e-12


def test_loop_between_equals_independent_restatement():
    rng = np.random.default_rng(2)
    for flavour in (False, True):
        for _ in range(100):
            cur = np.concatenate([rng.normal(0, 30, 3), rng.uniform(-0.3, 0.3, 2), rng.uniform(-np.pi, np.pi, 1)]).astype(np.float32)
            pre = np.concatenate([rng.normal(0, 30, 3), rng.uniform(-0.3, 0.3, 2), rng.uniform(-np.pi, np.pi, 1)]).astype(np.float32)
            corr = np.concatenate([rng.normal(0, 0.5, 3), rng.uniform(-0.1, 0.1, 3)])
            T = _mat(corr).astype(np.float32)
            got = engine.wire_loop_between(T, cur, pre, flavour)
            frm = T.astype(np.float64) @ _mat(cur.astype(np.float64))
            bet = np.linalg.inv(frm) @ _mat(pre.astype(np.float64))
            assert np.allclose(got[:3], bet[:3, 3], atol=2e-4), (got[:3], bet[:3, 3])        # float32 composition upstream of the doubles
            assert _same_rotation(got[3:], Rot.from_matrix(bet[:3, :3]).as_quat(), 2e-6)
            assert abs(np.linalg.norm(got[3:]) - 1.0) < 1e-9


def test_loop_between_identities():
    ident = np.eye(4, dtype=np.float32)
    p = np.array([1.5, -2.0, 0.3, 0.01, -0.02, 0.7], np.float32)
    t = engine.wire_loop_between(ident, p, p)                     # a keyframe against itself: identity
    assert np.allclose(t[:3], 0, atol=1e-5) and _same_rotation(t[3:], [0, 0, 0, 1], 1e-6)
    q = np.array([4.0, 1.0, -0.2, 0.0, 0.0, -1.1], np.float32)
    ab, ba = engine.wire_loop_between(ident, p, q), engine.wire_loop_between(ident, q, p)      # between(a, b) = between(b, a)^-1
    Rab, Rba = Rot.from_quat(ab[3:]), Rot.from_quat(ba[3:])
    assert _same_rotation((Rab * Rba).as_quat(), [0, 0, 0, 1], 1e-6)
    assert np.allclose(Rab.apply(ba[:3]) + np.array(ab[:3]), 0, atol=1e-5)


def test_message_round_trips():
    lib = engine.load_library()
    vals = np.arange(1200, dtype=np.float32) * 0.25
    cur = engine.SclTransform(1, 2, 3, 0, 0, 0, 1)
    pre6 = np.array([0.5, 0.25, 0.0, 0.0, 0.0, 0.3], np.float32)
    m = engine.SclGlobalDescriptor()
    lib.scl_wire_make_global_descriptor(42, vals.ctypes.data, 1200, C.byref(cur), 1, pre6.ctypes.data, C.byref(m))
    assert m.index == 42 and m.n_values == 1200 and m.cur_pose.as_tuple() == cur.as_tuple()
    assert np.allclose(m.pre_pose.as_tuple(), engine.wire_pose6_to_transform(pre6))
    need = lib.scl_wire_encode_global_descriptor(C.byref(m), None, 0)
    assert need == 4 + 112 + 4 + 4800
    buf = (C.c_ubyte * need)()
    assert lib.scl_wire_encode_global_descriptor(C.byref(m), buf, need - 1) == -1
    assert lib.scl_wire_encode_global_descriptor(C.byref(m), buf, need) == need
    d = engine.SclGlobalDescriptor()
    assert lib.scl_wire_decode_global_descriptor(buf, need, C.byref(d)) == need
    assert d.index == 42 and d.n_values == 1200 and d.pre_pose.as_tuple() == m.pre_pose.as_tuple()
    assert np.array_equal(np.ctypeslib.as_array(d.values, (1200,)), vals)
    assert lib.scl_wire_decode_global_descriptor(buf, need - 4, C.byref(d)) == -1
    # loop_info
    T = np.eye(4, dtype=np.float32)
    a = np.array([1, 2, 3, 0, 0, 0.5], np.float32)
    b = np.array([2, 2, 3, 0, 0, 0.4], np.float32)
    li = engine.SclLoopInfo()
    lib.scl_wire_make_loop_info(1, 300, 12, 0.17, T.ctypes.data, a.ctypes.data, b.ctypes.data, C.byref(li))
    assert (li.robot0, li.robot1, li.index0, li.index1) == (1, 1, 300, 12) and abs(li.noise - 0.17) < 1e-7
    assert np.allclose(li.bet_pose.as_tuple(), engine.wire_loop_between(T, a, b, False))
    buf2 = (C.c_ubyte * 76)()
    assert lib.scl_wire_encode_loop_info(C.byref(li), buf2, 76) == 76
    lj = engine.SclLoopInfo()
    assert lib.scl_wire_decode_loop_info(buf2, 76, C.byref(lj)) == 76
    assert (lj.robot0, lj.index0, lj.index1, lj.noise) == (li.robot0, li.index0, li.index1, li.noise) and lj.bet_pose.as_tuple() == li.bet_pose.as_tuple()
    assert lib.scl_wire_decode_loop_info(buf2, 75, C.byref(lj)) == -1
