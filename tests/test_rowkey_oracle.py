"""CPU: the restatement of the row-key candidate stage (oracle/rowkey_oracle.cpp) against the reference's own text of it
(oracle/_ref/libiris_ref.so, built where /root/reference exists) and against the fixtures generated from that text."""
import os

import numpy as np
import pytest

import rowkey_scenario as sc
from oracle_lib import IrisOracle, have_iris_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARAMS = dict(rows=80, num_exclude_recent=30, num_candidates=10, dist_thres=0.32)


def _same(a, b):
    for part in ("intra", "inter"):
        for k in ("id", "n", "cand"):
            assert np.array_equal(a[part][k], b[part][k]), (part, k)
        assert np.array_equal(a[part]["bias"], b[part]["bias"]), part
        assert np.array_equal(a[part]["d2"].view(np.uint32), b[part]["d2"].view(np.uint32)), part
    assert np.array_equal(a["index"], b["index"]) and np.array_equal(a["sizes"], b["sizes"])


@pytest.mark.skipif(not have_iris_ref(), reason="oracle/_ref/libiris_ref.so is built only where /root/reference exists")
@pytest.mark.parametrize("seed,this_id", [(1, 0), (2, 1), (3, 2)])
def test_restatement_equals_reference_text(seed, this_id):
    saves = sc.make(seed)
    port = sc.run(lambda **kw: IrisOracle(kind="port", **kw), saves, 3, this_id, **PARAMS)
    ref = sc.run(lambda **kw: IrisOracle(kind="ref", **kw), saves, 3, this_id, **PARAMS)
    _same(port, ref)
    if this_id != 2:
        assert (port["intra"]["id"] >= 0).sum() > 10 and (port["inter"]["id"] >= 0).sum() > 10  # the scenario does close loops
        assert (port["intra"]["n"] == 0).sum() >= 41                                               # the early return of :1094-1097
    else:
        assert (port["inter"]["n"] == 0).any() and (port["intra"]["n"] == 0).all()                 # robot 2 holds 7 keys: fewer than numCandidates + 1 (:1194-1197)


def few_neighbours_saves(rows=80):
    """Robot 1 holds 12 keys of which 5 equal robot 0's key 3 exactly: an inter query of that key finds 7 neighbours, libnabo
    leaves the other three slots at -1 / +inf and the candidate loop must skip them (descriptor.h:1214-1218)."""
    rng = np.random.default_rng(5)
    k0 = rng.uniform(0.5, 6.0, (6, rows)).astype(np.float32)
    k1 = rng.uniform(0.5, 6.0, (12, rows)).astype(np.float32)
    k1[[1, 4, 5, 8, 11]] = k0[3]
    saves = [(k0[i], 0, i, float(i)) for i in range(6)] + [(k1[i], 1, i, 3.1 if i == 7 else 50.0 + i) for i in range(12)]
    return saves


@pytest.mark.skipif(not have_iris_ref(), reason="oracle/_ref/libiris_ref.so is built only where /root/reference exists")
def test_fewer_neighbours_than_candidates():
    saves = few_neighbours_saves()
    port = sc.run(lambda **kw: IrisOracle(kind="port", **kw), saves, 2, 0, **PARAMS)
    ref = sc.run(lambda **kw: IrisOracle(kind="ref", **kw), saves, 2, 0, **PARAMS)
    _same(port, ref)
    assert port["inter"]["n"][3] == 10 and (port["inter"]["cand"][3] == -1).sum() == 3 and np.isinf(port["inter"]["d2"][3][-3:]).all()
    assert port["inter"]["id"][3] == 6 + 7                                                         # feature 3.1 against 3.0: the loop is found among the seven


def test_restatement_replays_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "rowkey_golden.npz"))
    for tag in ("a", "b"):
        seed, this_id, robots = [int(v) for v in g[tag + "_meta"]]
        saves = sc.make(seed)
        assert np.array_equal(np.stack([s[0] for s in saves]).view(np.uint32), g[tag + "_keys"].view(np.uint32))   # the generator has not drifted
        port = sc.run(lambda **kw: IrisOracle(kind="port", **kw), saves, robots, this_id, **PARAMS)
        for part in ("intra", "inter"):
            for k in ("id", "n", "cand", "bias"):
                assert np.array_equal(port[part][k], g[f"{tag}_{part}_{k}"]), (tag, part, k)
            assert np.array_equal(port[part]["d2"].view(np.uint32), g[f"{tag}_{part}_d2"].view(np.uint32))
        assert np.array_equal(port["index"], g[tag + "_index"])
