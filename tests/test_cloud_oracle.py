"""CPU tests of the cloud-preparation oracle (oracle/cloud_oracle.cpp): pcl::VoxelGrid<PointXYZI> and
loopFindNearKeyframes restated (distributedMapping.h:996-998, 1163-1186). PCL is absent from /root/reference, so the
oracle is checked against a hand-computed case and an independent numpy float32 restatement of PCL's published algorithm."""
import numpy as np

import oracle_lib


def _numpy_voxel_grid(pts, leaf):
    """voxel_grid.hpp applyFilter in numpy float32, sums in input order (small inputs only: python loop)."""
    f = np.float32
    ok = np.isfinite(pts[:, :3]).all(1)
    p = pts[ok]
    if len(p) == 0:
        return np.empty((0, 4), f)
    inv = f(1.0) / f(leaf)
    mn, mx = p[:, :3].min(0), p[:, :3].max(0)
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div = max_b - min_b + 1
    ijk = (np.floor(p[:, :3] * inv) - min_b.astype(f)).astype(np.int32)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    out = []
    i = 0
    while i < len(order):
        j = i
        s = np.zeros(4, f)
        while j < len(order) and idx[order[j]] == idx[order[i]]:
            s = (s + p[order[j], :4]).astype(f)
            j += 1
        out.append(s / f(j - i))
        i = j
    return np.array(out, f)


def test_voxel_grid_hand_case():
    pts = np.array([[0.1, 0.1, 0.1, 1.0], [0.9, 0.9, 0.9, 3.0],       # leaf (0,0,0)
                    [1.5, 0.5, 0.5, 5.0],                             # leaf (1,0,0)
                    [0.5, 1.5, 0.5, 7.0],                             # leaf (0,1,0)
                    [np.nan, 0.0, 0.0, 9.0],                          # skipped
                    [0.5, 0.5, 1.5, 2.0], [0.25, 0.75, 1.25, 4.0]], np.float32)   # leaf (0,0,1)
    got = oracle_lib.voxel_grid_pcl(pts, 1.0)
    exp = np.array([[0.5, 0.5, 0.5, 2.0], [1.5, 0.5, 0.5, 5.0], [0.5, 1.5, 0.5, 7.0], [0.375, 0.625, 1.375, 3.0]], np.float32)
    assert np.array_equal(got, exp)


def test_voxel_grid_equals_numpy_restatement():
    rng = np.random.default_rng(11)
    pts = np.concatenate([rng.normal(0, 6, (1500, 3)), rng.uniform(0, 100, (1500, 1))], 1).astype(np.float32)
    pts[::97, 1] = np.inf
    for leaf in (0.4, 1.0, 2.5):
        got = oracle_lib.voxel_grid_pcl(pts, leaf)
        exp = _numpy_voxel_grid(pts, leaf)
        assert got.shape == exp.shape and np.array_equal(got.view(np.uint32), exp.view(np.uint32)), leaf


def test_voxel_grid_edges():
    assert oracle_lib.voxel_grid_pcl(np.empty((0, 4), np.float32), 0.4).shape == (0, 4)
    allnan = np.full((5, 4), np.nan, np.float32)
    assert oracle_lib.voxel_grid_pcl(allnan, 0.4).shape == (0, 4)
    one = np.tile(np.array([[3.1, -2.2, 0.3, 8.0]], np.float32), (9, 1))
    got = oracle_lib.voxel_grid_pcl(one, 0.4)
    assert got.shape == (1, 4) and np.allclose(got[0], one[0], rtol=1e-6)
    # a leaf grid whose linear index would overflow an int: PCL warns and returns the input unchanged
    far = np.array([[0, 0, 0, 1], [1e6, 1e6, 1e6, 2]], np.float32)
    assert np.array_equal(oracle_lib.voxel_grid_pcl(far, 0.01), far)


def test_assemble_submap():
    rng = np.random.default_rng(12)
    clouds = [np.concatenate([rng.normal(0, 5, (n, 3)), rng.uniform(0, 1, (n, 1))], 1).astype(np.float32) for n in (40, 0, 25)]
    ident = np.zeros((3, 6), np.float32)
    assert np.array_equal(oracle_lib.assemble_submap(clouds, ident, 0.0), np.concatenate(clouds))
    yaw90 = np.array([[1.0, 2.0, 3.0, 0.0, 0.0, np.pi / 2]], np.float32)
    got = oracle_lib.assemble_submap([np.array([[1.0, 0.0, 0.0, 5.0]], np.float32)], yaw90, 0.0)
    assert np.allclose(got[0], [1.0, 3.0, 3.0, 5.0], atol=1e-6)
    poses = rng.normal(0, 1, (3, 6)).astype(np.float32)
    world = oracle_lib.assemble_submap(clouds, poses, 0.0)
    assert np.array_equal(oracle_lib.assemble_submap(clouds, poses, 0.4).view(np.uint32), oracle_lib.voxel_grid_pcl(world, 0.4).view(np.uint32))
