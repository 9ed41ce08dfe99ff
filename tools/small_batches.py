"""K2+K3+K4 latency of small host-buffer batches by key (auto mode) on a 1 M-keyframe database."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_slam_b200 import synth, engine
dev = torch.device("cuda:0"); N = 1 << 20
e = engine.ScanContextB200(numCandidates=10); e.reserve(N)
for c0 in range(0, N, 1 << 17): e.insert_batch_dev(synth.desc_db(1 << 17, device=dev, start=c0))
torch.cuda.synchronize()
for Q in (1, 4, 8, 9, 16, 32, 64, 128):
    ids = np.arange(N - Q, N, dtype=np.int32)
    for mode in (0, 1):
        e.set_knn_mode(mode)
        for _ in range(3): r = e.query_batch(q_ids=ids, K=10, n_db=N - 200, metric=0)
        t0 = time.perf_counter()
        for _ in range(20): r = e.query_batch(q_ids=ids, K=10, n_db=N - 200, metric=0)
        dt = (time.perf_counter() - t0) / 20 * 1e6
        if mode == 0: auto = r
        else: same = all(np.array_equal(auto[k], r[k], equal_nan=True) for k in r)
        print(f"Q={Q:4d} mode {mode}: {dt:8.0f} us per call" + ("" if mode == 0 else f"   same results as auto: {same}"), flush=True)
