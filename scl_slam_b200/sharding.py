"""Database sharding across GPUs: rows by `key mod world`, queries replicated,
one all-gather of the per-rank top-K records, then a merge that reproduces the unsharded
result exactly. This module holds the host-side rules; the device-side merge is
`merge_shards_kernel` (csrc/k4_scdist.cu, C-ABI scl_merge_shards_dev)."""
import numpy as np


def local_count(n_global, rank, world):
    """How many of the global keys 0..n_global-1 live on `rank` (key mod world == rank)."""
    return (n_global - rank + world - 1) // world if n_global > rank else 0


def local_rows(n_global, rank, world):
    """Global keys owned by `rank`, in local-key order (local key l <-> global l*world + rank)."""
    return np.arange(rank, n_global, world, dtype=np.int64)


def local_search_bound(n_db_global, rank, world):
    """A global search range [0, n_db_global) is the local range [0, this) on `rank` — e.g. the
    reference's "exclude the most recent 100" rule (descriptor.h:1627,1696) stays an index bound."""
    return local_count(n_db_global, rank, world)


def merge_shards_numpy(all_ids, all_d2, all_dist, all_shift, q_ids=None):
    """Reference semantics of the merge: inputs [world, Q, K] (each rank's candidates in its own
    kNN order, id -1 = missing). Global top-K by (d2, id); winner = strict-< scan of the SC
    distance in that order, the query itself skipped (descriptor.h:1721-1737)."""
    world, Q, K = all_ids.shape
    out = dict(cand_ids=np.full((Q, K), -1, np.int32), cand_d2=np.full((Q, K), np.finfo(np.float32).max, np.float32),
               cand_dist=np.full((Q, K), np.nan), cand_shift=np.zeros((Q, K), np.int32),
               best_id=np.full(Q, -1, np.int32), best_dist=np.full(Q, 1e7), best_shift=np.zeros(Q, np.int32))
    for q in range(Q):
        recs = [(all_d2[w, q, k], all_ids[w, q, k], all_dist[w, q, k], all_shift[w, q, k])
                for w in range(world) for k in range(K) if all_ids[w, q, k] >= 0]
        recs.sort(key=lambda r: (r[0], r[1]))
        self_id = -1 if q_ids is None else q_ids[q]
        for k, (d2, i, dist, sh) in enumerate(recs[:K]):
            out["cand_ids"][q, k], out["cand_d2"][q, k], out["cand_dist"][q, k], out["cand_shift"][q, k] = i, d2, dist, sh
            if dist < out["best_dist"][q] and i != self_id:
                out["best_dist"][q], out["best_id"][q], out["best_shift"][q] = dist, i, sh
    return out


def merge_topk_numpy(all_ids, all_d2):
    """Two-phase exchange, point 1 (merge_topk_kernel / xchg_merge_topk_kernel): inputs [world, Q, K] per-rank lists in
    their own kNN order; output the global top-K by (d2, id), identical on every rank."""
    world, Q, K = all_ids.shape
    ids = np.full((Q, K), -1, np.int32)
    d2 = np.full((Q, K), np.finfo(np.float32).max, np.float32)
    for q in range(Q):
        recs = sorted((all_d2[w, q, k], all_ids[w, q, k]) for w in range(world) for k in range(K) if all_ids[w, q, k] >= 0)
        for k, (d, i) in enumerate(recs[:K]):
            ids[q, k], d2[q, k] = i, d
    return ids, d2


def combine_owned_numpy(cand_ids, all_dist, all_shift, q_ids=None):
    """Two-phase exchange, point 2 (combine_owned_kernel / xchg_combine_kernel): every candidate's SC distance comes from
    the rank that owns it (id mod world); then the strict-< winner scan in global kNN order (descriptor.h:1721-1737)."""
    world, Q, K = all_dist.shape
    out = dict(cand_dist=np.full((Q, K), np.nan), cand_shift=np.zeros((Q, K), np.int32),
               best_id=np.full(Q, -1, np.int32), best_dist=np.full(Q, 1e7), best_shift=np.zeros(Q, np.int32))
    for q in range(Q):
        self_id = -1 if q_ids is None else q_ids[q]
        for k in range(K):
            i = cand_ids[q, k]
            if i < 0:
                continue
            dist, sh = all_dist[i % world, q, k], all_shift[i % world, q, k]
            out["cand_dist"][q, k], out["cand_shift"][q, k] = dist, sh
            if dist < out["best_dist"][q] and i != self_id:
                out["best_dist"][q], out["best_id"][q], out["best_shift"][q] = dist, i, sh
    return out

