"""K2+K3 time of the tensor-core path launched call by call and replayed from a CUDA graph (the difference is host launch cost)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_slam_b200 import synth, engine
dev = torch.device("cuda:0"); N, K, Q = 1 << 20, 10, 1024
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    e = engine.ScanContextB200(numCandidates=K); e.set_stream(s.cuda_stream); e.reserve(N)
    for c0 in range(0, N, 1 << 17): e.insert_batch_dev(synth.desc_db(1 << 17, device=dev, start=c0))
    e.set_knn_mode(2, False)
    q = synth.desc_queries(synth.desc_db(1 << 16, device=dev), Q)[0]
    ids = torch.empty((Q, K), dtype=torch.int32, device=dev); d2 = torch.empty((Q, K), device=dev)
    def ev(fn, n=20):
        for _ in range(3): fn()
        s.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(n): fn()
        b.record(s); s.synchronize()
        return a.elapsed_time(b) / n * 1000
    for n_db in (N, N // 2, N // 8):
        f = lambda: e.knn_batch_dev(q, Q, K, n_db, 0, ids, d2)
        t_plain = ev(f)
        ref = ids.clone()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            f()
        t_graph = ev(g.replay)
        print(f"n_db={n_db}: K2+K3 call by call {t_plain:.1f} us, graph replay {t_graph:.1f} us, same ids {bool((ids == ref).all())}", flush=True)
