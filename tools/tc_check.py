"""Bring-up check of the BF16x3 tensor-core kNN (mode 3) against the exact kernel (mode 1), then K3 timing at the bench workload."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_slam_b200 import synth, engine

MODE = int(os.environ.get("TC_MODE", "2"))

def run(R, S, n, Q, K, metric=0, n_db=None, ordered=False):
    dev = torch.device("cuda:0")
    db = synth.desc_db(n, R, S, seed=3, device=dev)
    if ordered:   # a database ordered like a trajectory: neighbouring keys are alike, whole stretches are far from any one query
        db = db[torch.argsort(db.reshape(n, -1).mean(1))].contiguous()
    q, src, shift = synth.desc_queries(db[: min(n, 1 << 16)], Q, seed=4)
    e = engine.ScanContextB200(numRing=R, numSector=S, numCandidates=K)
    torch.cuda.synchronize()      # the engine runs on its own stream: the generated data must be complete before it reads it
    e.insert_batch_dev(db)
    torch.cuda.synchronize()
    qh = q.cpu().numpy()
    n_db = n if n_db is None else n_db
    e.set_knn_mode(1)
    a = e.query_batch(q_desc=qh, K=K, n_db=n_db, metric=metric)
    e.set_knn_mode(MODE, True)
    t = time.time()
    b = e.query_batch(q_desc=qh, K=K, n_db=n_db, metric=metric)
    st = e.knn_stats()
    same = {k: bool(np.array_equal(a[k], b[k], equal_nan=True)) for k in a}
    print(f"R={R} S={S} n={n} n_db={n_db} Q={Q} K={K} metric={metric} ordered={ordered}: {same} stats={st} t={time.time()-t:.3f}s", flush=True)
    if not all(same.values()):
        bad = np.where((a["cand_ids"] != b["cand_ids"]).any(1))[0]
        print("  mismatching queries", len(bad), bad[:5])
        for i in bad[:2]:
            print("   exact", a["cand_ids"][i], a["cand_d2"][i]); print("   tc   ", b["cand_ids"][i], b["cand_d2"][i])
    return all(same.values())

def timing():
    dev = torch.device("cuda:0"); N, Q, K = 1 << 20, 1024, 10
    e = engine.ScanContextB200(numCandidates=K); e.set_stream(torch.cuda.current_stream().cuda_stream); e.reserve(N)
    for c0 in range(0, N, 1 << 17): e.insert_batch_dev(synth.desc_db(1 << 17, device=dev, start=c0))
    q = synth.desc_queries(synth.desc_db(1 << 16, device=dev), Q)[0]
    ids = torch.empty((Q, K), dtype=torch.int32, device=dev); d2 = torch.empty((Q, K), device=dev)
    res = {}
    for mode in (1, MODE):
        e.set_knn_mode(mode, False)
        for _ in range(3): e.knn_batch_dev(q, Q, K, N, 0, ids, d2)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): e.knn_batch_dev(q, Q, K, N, 0, ids, d2)
        b.record(); torch.cuda.synchronize()
        res[mode] = ids.cpu().numpy().copy()
        print(f"mode {mode}: K2+K3 {a.elapsed_time(b) / 10 * 1000:.1f} us per batch of {Q} queries, N={N}", flush=True)
    print("ids equal between modes:", bool(np.array_equal(res[1], res[MODE])))
    e.set_knn_mode(MODE, True)
    e.knn_batch_dev(q, Q, K, N, 0, ids, d2); torch.cuda.synchronize()
    print("stats", e.knn_stats())

if __name__ == "__main__":
    ok = True
    if "--time-only" not in sys.argv:
        ok &= run(20, 60, 40000, 256, 10)
        ok &= run(20, 60, 300000, 1024, 10)
        ok &= run(20, 60, 70001, 130, 3, metric=1)
        ok &= run(20, 60, 70001, 300, 10, n_db=69901)
        ok &= run(40, 120, 50000, 200, 10)
        ok &= run(20, 60, 200000, 2500, 10)
        ok &= run(20, 60, 300000, 1024, 10, ordered=True)
        print("TC_CHECK", "OK" if ok else "FAIL", flush=True)
    timing()
