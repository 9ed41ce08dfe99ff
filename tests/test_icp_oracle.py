"""CPU checks of the verification oracle (oracle/icp_oracle.cpp; PCL is not in this image, so it stays "parity unpinned"
against PCL itself — DESIGN.md §2): planted rigid offsets are recovered, the fitness is the brute-force mean squared
nearest-neighbour distance, and the RANSAC restatement accepts a matching pair and rejects a displaced one."""
import numpy as np

import oracle_lib


def _rigid(yaw, pitch, roll, t):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T


def _pad(xyz):
    out = np.zeros((xyz.shape[0], 4), np.float32)
    out[:, :3] = xyz
    return out


def _scene(seed, n=4000):
    """Well-separated points (a jittered lattice), so that a small offset keeps every nearest neighbour the right one."""
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(np.arange(20), np.arange(20), np.arange(10), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    pts = g * 2.0 + rng.uniform(-0.3, 0.3, size=g.shape)
    return pts[rng.permutation(len(pts))[:n]]


def test_icp_recovers_planted_transform():
    tgt = _scene(1)
    T_true = _rigid(0.03, -0.01, 0.015, (0.25, -0.18, 0.1))
    # source = target moved by the inverse, so that ICP's answer (source -> target) is T_true
    Ti = np.linalg.inv(T_true)
    src = tgt[:1500] @ Ti[:3, :3].T + Ti[:3, 3]
    T, fit, conv, it = oracle_lib.icp(_pad(src), _pad(tgt))
    assert conv and 1 <= it <= 50
    assert np.linalg.norm(T[:3, 3] - T_true[:3, 3]) < 1e-3
    assert np.arccos(np.clip((np.trace(T[:3, :3].T.astype(np.float64) @ T_true[:3, :3]) - 1) / 2, -1, 1)) < 1e-4
    assert fit < 1e-6
    assert np.allclose(T[3], [0, 0, 0, 1])
    assert abs(np.linalg.det(T[:3, :3].astype(np.float64)) - 1) < 1e-5


def test_icp_fitness_is_mean_squared_nn_distance():
    rng = np.random.default_rng(5)
    tgt = _scene(2)
    src = tgt[:900] + rng.normal(0, 0.05, size=(900, 3))
    T, fit, conv, it = oracle_lib.icp(_pad(src), _pad(tgt))
    moved = (src @ T[:3, :3].T.astype(np.float64) + T[:3, 3]).astype(np.float32)
    idx, d2 = oracle_lib.nn_bruteforce(_pad(moved), _pad(tgt))
    # independent brute force in numpy on a subset
    sub = moved[:100].astype(np.float64)
    dd = ((sub[:, None, :] - tgt[None, :, :].astype(np.float32).astype(np.float64)) ** 2).sum(-1)
    assert np.array_equal(dd.argmin(1), idx[:100])
    assert np.allclose(dd.min(1), d2[:100], rtol=1e-4, atol=1e-7)
    assert abs(fit - float(np.mean(d2.astype(np.float64)))) <= 1e-4 * fit + 1e-7
    assert 0.001 < fit < 0.02                       # three axes of sigma 0.05, minus what the fit absorbs


def test_icp_max_correspondence_distance_gates_pairs():
    tgt = _scene(3)
    src = tgt[:600].copy()
    src[:40] += 500.0                                # outliers farther than the gate: must not drag the estimate
    T, fit, conv, it = oracle_lib.icp(_pad(src), _pad(tgt), max_corr_dist=5.0)
    assert conv
    assert np.linalg.norm(T[:3, 3]) < 1e-3 and np.allclose(T[:3, :3], np.eye(3), atol=1e-4)


def test_ransac_accepts_match_and_rejects_mismatch():
    tgt = _scene(4)
    T_true = _rigid(0.01, 0.0, 0.0, (0.06, -0.04, 0.02))
    Ti = np.linalg.inv(T_true)
    src = tgt[:1200] @ Ti[:3, :3].T + Ti[:3, 3]
    T, nc, ni, ok = oracle_lib.verify_ransac(_pad(src), _pad(tgt), min_inlier_ratio=0.45, seed=3)
    assert ok and nc == 1200 and ni >= 0.9 * nc
    assert np.linalg.norm(T[:3, 3] - T_true[:3, 3]) < 0.02
    T2, nc2, ni2, ok2 = oracle_lib.verify_ransac(_pad(src), _pad(tgt), min_inlier_ratio=0.45, seed=3)
    assert np.array_equal(T, T2) and ni == ni2      # reproducible for a given seed
    rng = np.random.default_rng(9)
    other = rng.uniform(0, 40, size=(1200, 3))      # unrelated geometry
    _, nc3, ni3, ok3 = oracle_lib.verify_ransac(_pad(other), _pad(tgt), min_inlier_ratio=0.75, seed=3)
    assert not ok3 and ni3 < 0.75 * nc3
