"""GPU: the row-key candidate search behind include/scl_rowkey.h (the kNN stage of the reference's lidar_iris_descriptor,
descriptor.h:1047-1059, 1087-1267) against the CPU oracle (oracle/rowkey_oracle.cpp, itself checked against the reference's
own text) and the fixtures generated from that text — ids, biases, kNN lists and float distances bit for bit."""
import os

import numpy as np
import pytest
import torch

import rowkey_scenario as sc
from oracle_lib import IrisOracle, iris_compare
from test_rowkey_oracle import PARAMS, few_neighbours_saves

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rk_mod():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from scl_slam_b200 import build, rowkey
    build.build()
    return rowkey


def _gpu_run(rk_mod, saves, robot_num, this_id, batch_saves=False):
    e = rk_mod.LidarIrisRowKeysB200(rows=PARAMS["rows"], numExcludeRecent=PARAMS["num_exclude_recent"], numCandidates=PARAMS["num_candidates"],
                                    distThres=PARAMS["dist_thres"], robotNum=robot_num, thisID=this_id)
    feat, tag = {}, {}
    counts = [0] * robot_num
    for g, (key, r, i, f) in enumerate(saves):
        assert e.save(key, r, i) == g
        feat[(r, counts[r])] = f; tag[(r, counts[r])] = g
        counts[r] += 1

    def compare(ra, la, rb, lb):
        return iris_compare(feat[(ra, la)], tag[(ra, la)], feat[(rb, lb)], tag[(rb, lb)])
    K = PARAMS["num_candidates"]
    out = {}
    rows = []
    for p in range(counts[this_id]):
        i, b = e.detectIntraLoopClosureID(p, compare)
        n, idx, d2 = e.intra_candidates(p)
        rows.append((i, b, n, idx, d2))
    out["intra"] = rows
    rows = []
    for g in range(len(saves)):
        i, b = e.detectInterLoopClosureID(g, compare)
        n, idx, gk, d2 = e.inter_candidates(g)
        rows.append((i, b, n, idx, d2))
    out["inter"] = rows
    out["index"] = np.array([e.getIndex(g) for g in range(len(saves))], np.int32)
    out["sizes"] = np.array([e.getSize(-1)] + [e.getSize(r) for r in range(robot_num)], np.int32)
    assert e.getIndex(len(saves)) == (-1, -1)
    return out


def _check(got, exp):
    for part in ("intra", "inter"):
        for q, row in enumerate(got[part]):
            i, b, n, idx, d2 = row
            assert i == exp[part]["id"][q] and b == exp[part]["bias"][q], (part, q, i, b, exp[part]["id"][q], exp[part]["bias"][q])
            assert n == exp[part]["n"][q], (part, q)
            if n:
                assert np.array_equal(idx, exp[part]["cand"][q]), (part, q, idx, exp[part]["cand"][q])
                assert np.array_equal(d2.view(np.uint32), exp[part]["d2"][q].view(np.uint32)), (part, q)
    assert np.array_equal(got["index"], exp["index"]) and np.array_equal(got["sizes"], exp["sizes"])


@pytest.mark.parametrize("seed,this_id", [(1, 0), (2, 1), (3, 2)])
def test_detect_equals_oracle(rk_mod, seed, this_id):
    saves = sc.make(seed)
    exp = sc.run(lambda **kw: IrisOracle(kind="port", **kw), saves, 3, this_id, **PARAMS)
    _check(_gpu_run(rk_mod, saves, 3, this_id), exp)


def test_fewer_neighbours_than_candidates(rk_mod):
    saves = few_neighbours_saves()
    exp = sc.run(lambda **kw: IrisOracle(kind="port", **kw), saves, 2, 0, **PARAMS)
    got = _gpu_run(rk_mod, saves, 2, 0)
    _check(got, exp)
    assert (got["inter"][3][3] == -1).sum() == 3 and got["inter"][3][0] == 13


def test_golden_on_gpu(rk_mod):
    """tests/golden/rowkey_golden.npz was generated from the reference's own text (tests/golden/make_golden_rowkey.py)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "rowkey_golden.npz"))
    for t in ("a", "b"):
        seed, this_id, robots = [int(v) for v in g[t + "_meta"]]
        saves = [(g[t + "_keys"][i], int(g[t + "_robot"][i]), int(g[t + "_idx"][i]), float(g[t + "_feat"][i])) for i in range(len(g[t + "_robot"]))]
        exp = {part: {k: g[f"{t}_{part}_{k}"] for k in ("id", "bias", "n", "cand", "d2")} for part in ("intra", "inter")}
        exp["index"] = g[t + "_index"]; exp["sizes"] = g[t + "_sizes"]
        _check(_gpu_run(rk_mod, saves, robots, this_id), exp)


def _walk_keys(n, rows, seed):
    rng = np.random.default_rng(seed)
    base = rng.uniform(0.5, 6.0, rows).astype(np.float32)
    out = np.empty((n, rows), np.float32)
    for c0 in range(0, n, 4096):                              # many trajectories of 4096 places around different bases
        m = min(4096, n - c0)
        out[c0:c0 + m] = np.abs(base + rng.normal(0, 1.0, rows) + np.cumsum(rng.normal(0, 0.05, (m, rows)), axis=0)).astype(np.float32)
    return out


@pytest.mark.parametrize("rows", [80, 20])
def test_batch_tensor_core_equals_exact_and_oracle(rk_mod, rows):
    """Batches against 20 000 + 18 000 keys of two other robots (the inter key set of this robot) and against this robot's
    own first 30 000 keys: the tensor-core prefilter + exact re-rank (mode 2) gives the lists of the exact kernel (mode 1),
    and both give the oracle's on a sample; the concatenation offsets and newLocal2Global come out right."""
    K = 10
    e = rk_mod.LidarIrisRowKeysB200(rows=rows, numCandidates=K, robotNum=3, thisID=1)
    o = IrisOracle(rows=rows, num_candidates=K, robot_num=3, this_id=1)
    sizes = {0: 20000, 1: 33000, 2: 18000}
    keys = {r: _walk_keys(n, rows, 100 + r) for r, n in sizes.items()}
    for r in (2, 0, 1):                                        # global keys follow the save order, not the robot order
        e.save_batch(keys[r], r)
    assert e.getSize(-1) == sum(sizes.values()) and e.getSize(2) == 18000 and e.getIndex(18000) == (0, 0)
    rng = np.random.default_rng(7)
    nq = 300
    pick = rng.integers(0, 20000, nq)
    q = (keys[0][pick] + rng.normal(0, 0.02, (nq, rows))).astype(np.float32)
    q[:5] = keys[0][pick[:5]]                                  # exact copies of stored keys: the self-match rule applies
    res = {m: e.knn_batch(q, from_robot=1, K=K, knn_mode=m) for m in (1, 2, 0)}
    st = e.knn_stats()
    assert st["tc_queries"] == 2 * 2 * nq, st               # modes 2 and 0 (automatic: >= 16 384 keys), two key sets each
    for m in (2, 0):
        for a, b in zip(res[1], res[m]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), m
    idx, gk, d2 = res[1]
    assert (idx[:, 0] == pick).mean() > 0.9                    # robot 0 comes first in the concatenation (descriptor.h:1164-1179)
    off = np.where(idx < 20000, 18000 + idx, idx - 20000)      # saved order: robot 2 (0..17999), robot 0 (18000..37999)
    assert np.array_equal(gk, off)
    # the oracle on a sample: robot 0's and robot 2's lists merged by (d2, concatenated index)
    s = 24
    i0, d0 = o_knn(o, keys[0], q[:s], K); i2, d2o = o_knn(o, keys[2], q[:s], K)
    for j in range(s):
        cat = sorted([(d0[j, t], int(i0[j, t])) for t in range(K) if i0[j, t] >= 0] + [(d2o[j, t], 20000 + int(i2[j, t])) for t in range(K) if i2[j, t] >= 0])[:K]
        assert [c[1] for c in cat] == list(idx[j]) and np.array_equal(np.array([c[0] for c in cat], np.float32).view(np.uint32), d2[j].view(np.uint32)), j
    # the intra key set: this robot's first 30 000 keys
    qi = (keys[1][rng.integers(0, 30000, nq)] + rng.normal(0, 0.02, (nq, rows))).astype(np.float32)
    a = e.knn_batch(qi, from_robot=-1, n_limit=30000, K=K, knn_mode=1)
    b = e.knn_batch(qi, from_robot=-1, n_limit=30000, K=K, knn_mode=2)
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    assert a[0].max() < 30000 and np.array_equal(a[1], 38000 + a[0])
    # smooth 4096-place trajectories are the hard case of the prefilter (every key of the dozen tiles around the query's
    # place passes the union bound): the re-rank's streaming cut must keep them off the exact kernel
    assert e.knn_stats()["fallback_queries"] <= 5 * nq // 10, e.knn_stats()


def o_knn(o, keys, q, K):
    """The oracle's linear scan over an explicit key matrix (a throw-away handle holding just these keys)."""
    t = IrisOracle(rows=keys.shape[1], num_candidates=K, robot_num=1, this_id=0)
    for c0 in range(0, keys.shape[0], 1):
        t.save(keys[c0], 0, c0, 0.0)
    return t.knn_batch(q, 0, keys.shape[0], K, threads=4)
