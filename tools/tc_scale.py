"""K2+K3 time of the BF16x3 path against the database size: slope (per-tile cost) and intercept (fixed cost)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_slam_b200 import synth, engine
dev = torch.device("cuda:0"); N, K = 1 << 20, 10
e = engine.ScanContextB200(numCandidates=K); e.set_stream(torch.cuda.current_stream().cuda_stream); e.reserve(N)
for c0 in range(0, N, 1 << 17): e.insert_batch_dev(synth.desc_db(1 << 17, device=dev, start=c0))
e.set_knn_mode(int(os.environ.get("TC_MODE", "2")), False)
for Q in (1024, 256):
    q = synth.desc_queries(synth.desc_db(1 << 16, device=dev), Q)[0]
    ids = torch.empty((Q, K), dtype=torch.int32, device=dev); d2 = torch.empty((Q, K), device=dev)
    for n_db in (1 << 20, 1 << 19, 1 << 18, 1 << 17, 1 << 16):
        for _ in range(3): e.knn_batch_dev(q, Q, K, n_db, 0, ids, d2)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): e.knn_batch_dev(q, Q, K, n_db, 0, ids, d2)
        b.record(); torch.cuda.synchronize()
        print(f"Q={Q} n_db={n_db}: K2+K3 {a.elapsed_time(b) / 10 * 1000:.1f} us", flush=True)
