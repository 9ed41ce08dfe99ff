"""Bring-up check of the tensor-core kNN: mode 2 vs mode 1 on the same engine, plus fallback count."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_slam_b200 import synth, engine

def run(R, S, n, Q, K, metric=0):
    dev = torch.device("cuda:0")
    db = synth.desc_db(n, R, S, seed=3, device=dev)
    q, src, shift = synth.desc_queries(db[: min(n, 1 << 16)], Q, seed=4)
    e = engine.ScanContextB200(numRing=R, numSector=S, numCandidates=K)
    e.insert_batch_dev(db)
    torch.cuda.synchronize()
    qh = q.cpu().numpy()
    e.set_knn_mode(1)
    a = e.query_batch(q_desc=qh, K=K, n_db=n, metric=metric)
    e.set_knn_mode(2, True)
    t = time.time()
    b = e.query_batch(q_desc=qh, K=K, n_db=n, metric=metric)
    st = e.knn_stats()
    same = {k: bool(np.array_equal(a[k], b[k], equal_nan=True)) for k in a}
    print(f"R={R} S={S} n={n} Q={Q} K={K} metric={metric}: {same} stats={st} t={time.time()-t:.3f}s", flush=True)
    if not all(same.values()):
        bad = np.where((a["cand_ids"] != b["cand_ids"]).any(1))[0]
        print("  first mismatching queries", bad[:5])
        for i in bad[:2]:
            print("   exact", a["cand_ids"][i], a["cand_d2"][i]); print("   tc   ", b["cand_ids"][i], b["cand_d2"][i])
    return all(same.values())

ok = True
ok &= run(20, 60, 40000, 256, 10)
ok &= run(20, 60, 300000, 1024, 10)
ok &= run(20, 60, 70001, 130, 3, metric=1)
ok &= run(40, 120, 50000, 200, 10)
print("TC_CHECK", "OK" if ok else "FAIL")
