"""The in-process sharded engine (scl_create_sharded, include/scl_engine.h) on two real GPUs against the single-device
engine on the same inputs: same candidates, distances, shifts and winners, batch by batch and call by call. Skipped on
boxes with one GPU (the exchange kernels of different ranks must run on different devices: B200_PROFILING.md)."""
import numpy as np
import pytest
import torch

from scl_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_in_process_equals_single_device(world):
    from scl_slam_b200 import engine
    if torch.cuda.device_count() < world:
        pytest.skip("not enough GPUs")
    n, nq, K = 70000, 250, 10                       # 35k keys per shard at world 2: the tensor-core kNN on every shard
    db = synth.desc_db(n, seed=101).numpy()
    db[21] = db[20]
    q = synth.desc_queries(torch.from_numpy(db), nq, seed=102)[0].numpy()
    one = engine.ScanContextB200(numCandidates=K)
    one.insert_batch(db)
    sh = engine.ShardedScanContextB200(list(range(world)), numCandidates=K, max_q=256, max_k=K)
    sh.insert_batch(db[:30000])
    sh.insert_batch(db[30000:])
    assert sh.getSize() == n and np.array_equal(_bits(sh.desc(12345)), _bits(db[12345]))
    for n_db in (n, n - 1001):
        exp = one.query_batch(q_desc=q, K=K, n_db=n_db)
        for rep in range(3):                          # the lanes rotate: every repetition runs on another one
            got = sh.query_batch(q, K=K, n_db=n_db)
            for k in exp:
                assert np.array_equal(got[k], exp[k], equal_nan=True), (world, n_db, rep, k)
    # ragged batch (padding inside), queries that are database entries (self-skip rule)
    ids = np.arange(500, 537, dtype=np.int32)
    exp = one.query_batch(q_ids=ids, K=K, n_db=n)
    got = sh.query_batch(db[ids], K=K, n_db=n, q_ids=ids)
    for k in ("cand_ids", "cand_shift", "best_id", "best_shift"):
        assert np.array_equal(got[k], exp[k]), k
    assert np.array_equal(_bits(got["cand_dist"]), _bits(exp["cand_dist"]))
    # the reference's own calls
    for cur in (n - 1, n - 57, 40000):
        assert sh.detectIntraLoopClosureID(cur) == one.detectIntraLoopClosureID(cur)
        assert sh.detectInterLoopClosureID(cur) == one.detectInterLoopClosureID(cur)
    assert sh.getIndex(-1) == (-1, -1) and sh.getIndex(777) == one.getIndex(777)
