"""N>1 host logic on CPU: two gloo ranks each hold `key mod 2 == rank` of a seeded database
(the CPU oracle stands in for the per-rank engine), exchange their local top-K records with one
all_gather, and the merge rule must reproduce the unsharded oracle result exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, nq, K, ret):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_lib import Oracle
    from scl_slam_b200 import sharding, synth
    db = synth.desc_db(n, seed=71).numpy().reshape(n, -1)
    q = synth.desc_queries(torch.from_numpy(db.reshape(n, 20, 60)), nq, seed=72)[0].numpy().reshape(nq, -1)
    rows = sharding.local_rows(n, rank, world)
    n_local = sharding.local_count(n, rank, world)
    assert len(rows) == n_local
    o = Oracle(num_candidates=K)
    o.bulk_load(np.concatenate([db[rows], q]))
    n_search = sharding.local_search_bound(n - 101, rank, world)          # a global "exclude recent" bound
    loc = o.query_batch(np.arange(n_local, n_local + nq), n_search, K, 0)
    ids = np.where(loc["cand_ids"] >= 0, loc["cand_ids"].astype(np.int64) * world + rank, -1).astype(np.int32)
    gathered = {}
    for name, arr in (("ids", ids), ("d2", loc["cand_d2"]), ("dist", loc["cand_dist"]), ("shift", loc["cand_shift"])):
        t = torch.from_numpy(np.ascontiguousarray(arr))
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        gathered[name] = np.stack([x.numpy() for x in outs])
    merged = sharding.merge_shards_numpy(gathered["ids"], gathered["d2"], gathered["dist"], gathered["shift"])
    if rank == 0:
        full = Oracle(num_candidates=K)
        full.bulk_load(np.concatenate([db, q]))
        exp = full.query_batch(np.arange(n, n + nq), n - 101, K, 0)
        ok = all(np.array_equal(merged[k], exp[k], equal_nan=True) for k in exp)
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def _worker_two_phase(rank, world, port, n, nq, K, ret):
    """The protocol bench.py runs at N > 1: kNN per shard -> exchange (id, d2) -> global top-K -> SC distance for the owned
    candidates only -> exchange (dist, shift) -> owner pick + winner scan."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_lib import Oracle
    from scl_slam_b200 import sharding, synth
    db = synth.desc_db(n, seed=73).numpy().reshape(n, -1)
    q = synth.desc_queries(torch.from_numpy(db.reshape(n, 20, 60)), nq, seed=74)[0].numpy().reshape(nq, -1)
    rows = sharding.local_rows(n, rank, world)
    n_local = sharding.local_count(n, rank, world)
    o = Oracle(num_candidates=K)
    o.bulk_load(np.concatenate([db[rows], q]))
    n_search = sharding.local_search_bound(n - 57, rank, world)
    loc = o.query_batch(np.arange(n_local, n_local + nq), n_search, K, 0)
    ids = np.where(loc["cand_ids"] >= 0, loc["cand_ids"].astype(np.int64) * world + rank, -1).astype(np.int32)

    def gather(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr))
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return np.stack([x.numpy() for x in outs])
    g_ids, g_d2 = sharding.merge_topk_numpy(gather(ids), gather(loc["cand_d2"]))
    own_dist = np.full((nq, K), np.nan); own_shift = np.zeros((nq, K), np.int32)
    for qi in range(nq):
        for k in range(K):
            i = int(g_ids[qi, k])
            if i >= 0 and i % world == rank:
                own_dist[qi, k], own_shift[qi, k] = o.distance(n_local + qi, i // world)
    fin = sharding.combine_owned_numpy(g_ids, gather(own_dist), gather(own_shift))
    if rank == 0:
        full = Oracle(num_candidates=K)
        full.bulk_load(np.concatenate([db, q]))
        exp = full.query_batch(np.arange(n, n + nq), n - 57, K, 0)
        ok = np.array_equal(g_ids, exp["cand_ids"]) and np.array_equal(g_d2, exp["cand_d2"]) and \
            all(np.array_equal(fin[k], exp[k], equal_nan=True) for k in fin)
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_phase_exchange_equals_unsharded_gloo(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker_two_phase, args=(r, world, port, 1203, 48, 10, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True


def test_two_rank_merge_equals_unsharded():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1501, 64, 10, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True


def test_local_rows_partition():
    from scl_slam_b200 import sharding
    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 4, 8):
            rows = [sharding.local_rows(n, r, world) for r in range(world)]
            assert sorted(np.concatenate(rows).tolist()) == list(range(n))
            assert [len(x) for x in rows] == [sharding.local_count(n, r, world) for r in range(world)]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_merge_rules_equal_unsharded_with_ties(world):
    """Host-side rules only (no processes): heavy d2 ties, short shards and NaN distances. The one-phase merge and the two-phase
    (global top-K, owner's distance, winner scan) rules give the same answer as ranking the whole database at once."""
    from scl_slam_b200 import sharding
    rng = np.random.default_rng(100 + world)
    Q, K, n = 37, 10, 23                                           # n < world * K: some ranks hold fewer than K keys
    d2 = rng.integers(0, 6, size=(Q, n)).astype(np.float32)        # few distinct values: ties decided by the lower id
    dist = rng.uniform(0, 1, size=(Q, n))
    dist[rng.uniform(size=(Q, n)) < 0.1] = np.nan                  # empty overlap: never selected (descriptor.h:1534,1561)
    dist[:, ::5] = np.round(dist[:, ::5], 1)                       # SC-distance ties: nearest ring key wins (strict <)
    shift = rng.integers(0, 60, size=(Q, n)).astype(np.int32)
    q_ids = rng.integers(0, n, size=Q).astype(np.int32)            # a query that is itself in the database is skipped as winner
    all_ids = np.full((world, Q, K), -1, np.int32); all_d2 = np.full((world, Q, K), np.finfo(np.float32).max, np.float32)
    all_dist = np.full((world, Q, K), np.nan); all_shift = np.zeros((world, Q, K), np.int32)
    for w in range(world):
        own = sharding.local_rows(n, w, world)
        for q in range(Q):
            order = sorted(own, key=lambda i: (d2[q, i], i))[:K]
            for k, i in enumerate(order):
                all_ids[w, q, k], all_d2[w, q, k], all_dist[w, q, k], all_shift[w, q, k] = i, d2[q, i], dist[q, i], shift[q, i]
    one = sharding.merge_shards_numpy(all_ids, all_d2, all_dist, all_shift, q_ids)
    g_ids, g_d2 = sharding.merge_topk_numpy(all_ids, all_d2)
    own_dist = np.full((world, Q, K), np.nan); own_shift = np.zeros((world, Q, K), np.int32)
    for q in range(Q):
        for k in range(K):
            i = int(g_ids[q, k])
            if i >= 0:
                own_dist[i % world, q, k], own_shift[i % world, q, k] = dist[q, i], shift[q, i]
    two = sharding.combine_owned_numpy(g_ids, own_dist, own_shift, q_ids)
    for q in range(Q):                                             # the unsharded ranking
        order = sorted(range(n), key=lambda i: (d2[q, i], i))[:K]
        assert g_ids[q].tolist() == order and one["cand_ids"][q].tolist() == order
        best, best_i, best_s = 1e7, -1, 0
        for i in order:
            if dist[q, i] < best and i != q_ids[q]:
                best, best_i, best_s = dist[q, i], i, shift[q, i]
        for res in (one, two):
            assert res["best_id"][q] == best_i and res["best_shift"][q] == best_s and res["best_dist"][q] == best
    assert np.array_equal(g_d2, one["cand_d2"])
    for k in ("cand_dist", "cand_shift", "best_id", "best_dist", "best_shift"):
        assert np.array_equal(one[k], two[k], equal_nan=True), k
