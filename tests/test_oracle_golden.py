"""The CPU restatement (oracle/liboracle.so) against the golden fixtures that were generated from
the reference's own class text + vendored nanoflann (oracle/_ref, tests/golden/make_golden.py).
Integer results (bins, candidate ids, loop ids, shifts) must be bit-exact; doubles are compared
bit-exact too because the stand-in Eigen and the restatement both sum in index order."""
import numpy as np
import pytest

from oracle_lib import Oracle


@pytest.mark.parametrize("ci", [0, 1, 2])
@pytest.mark.parametrize("rs", [(20, 60), (40, 120)])
def test_make_scancontext_matches_reference(golden, ci, rs):
    R, S = rs
    o = Oracle(num_ring=R, num_sector=S)
    got = o.make_scancontext(golden[f"cloud{ci}_pts"])
    assert np.array_equal(got.view(np.uint32), golden[f"cloud{ci}_desc_{R}x{S}"].view(np.uint32))


def _load_case(golden, tag, kind="port"):
    R, S, K, excl = [int(v) for v in golden[tag + "_params"]]
    o = Oracle(num_ring=R, num_sector=S, num_candidates=K, num_exclude_recent=excl, kind=kind)
    db = golden[tag + "_db"]
    for i in range(db.shape[0]):
        o.saveDescriptorAndKey(db[i], i % 3, i // 3)
    return o, db


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_keys_and_distances(golden, tag):
    o, db = _load_case(golden, tag)
    n = db.shape[0]
    assert o.getSize() == n
    assert o.getIndex(4) == (1, 1) and o.getIndex(-1) == (-1, -1) and o.getIndex(n) == (-1, -1)
    rk = np.stack([o.ring_key(i) for i in range(n)])
    assert np.array_equal(rk.view(np.uint32), golden[tag + "_ring_keys"].view(np.uint32))
    sk = np.stack([o.sector_key(i) for i in range(0, n, 7)])
    assert np.array_equal(sk.view(np.uint64), golden[tag + "_sector_keys"].view(np.uint64))
    pairs = golden[tag + "_pairs"]
    res = [o.distance(int(a), int(b)) for a, b in pairs]
    assert np.array_equal(np.array([r[1] for r in res]), golden[tag + "_pair_shift"])
    d = np.array([r[0] for r in res])
    assert np.array_equal(d.view(np.uint64), golden[tag + "_pair_dist"].view(np.uint64))
    assert np.array_equal(np.array([o.fast_align(int(a), int(b)) for a, b in pairs]), golden[tag + "_pair_align"])
    # Q8: an empty descriptor gives (1e7, 0)
    assert o.distance(7, 9) == (10000000.0, 0)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_detect_intra_inter(golden, tag):
    o, db = _load_case(golden, tag)
    n = db.shape[0]
    intra = [o.detectIntraLoopClosureID(i) for i in range(n)]
    inter = [o.detectInterLoopClosureID(i) for i in range(n)]
    assert np.array_equal(np.array([r[0] for r in intra]), golden[tag + "_intra_id"])
    assert np.array_equal(np.array([r[1] for r in intra], np.float32), golden[tag + "_intra_second"])
    assert np.array_equal(np.array([r[0] for r in inter]), golden[tag + "_inter_id"])
    assert np.array_equal(np.array([r[1] for r in inter], np.float32), golden[tag + "_inter_second"])
    assert (golden[tag + "_intra_id"] >= 0).sum() > 0, "fixture must contain detected loops"


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_knn_matches_nanoflann(golden, tag):
    """Brute-force kNN of the restatement vs the reference's vendored nanoflann tree
    (descriptor.h:1716): same float distances, same ids modulo equal-distance classes."""
    o, db = _load_case(golden, tag)
    n_db = int(golden[tag + "_knn_ndb"])
    for qi, cur in enumerate(golden[tag + "_knn_q"]):
        found, ids, d2 = o.knn(int(cur), n_db, 10, 0)
        assert found == 10
        assert np.array_equal(d2.view(np.uint32), golden[tag + "_knn_d2"][qi].view(np.uint32))
        gid = golden[tag + "_knn_ids"][qi]
        for dv in np.unique(d2):
            assert set(ids[d2 == dv]) == set(gid[golden[tag + "_knn_d2"][qi] == dv]) or (d2 == dv).sum() > 1
