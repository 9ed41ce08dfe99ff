"""Timing experiments on the tensor-core kNN path (SCL_TC_FLAGS disables parts of it; results are then wrong on purpose)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scl_slam_b200 import synth, engine
dev = torch.device("cuda:0"); N, K, Q = 1 << 20, 10, 1024
e = engine.ScanContextB200(numCandidates=K); e.set_stream(torch.cuda.current_stream().cuda_stream); e.reserve(N)
for c0 in range(0, N, 1 << 17): e.insert_batch_dev(synth.desc_db(1 << 17, device=dev, start=c0))
e.set_knn_mode(2, False)
q = synth.desc_queries(synth.desc_db(1 << 16, device=dev), Q)[0]
ids = torch.empty((Q, K), dtype=torch.int32, device=dev); d2 = torch.empty((Q, K), device=dev)
flags = [int(x) for x in sys.argv[1:]] or [0, 16, 48, 112]
for fl in flags:
    os.environ["SCL_TC_FLAGS"] = str(fl)
    for _ in range(3): e.knn_batch_dev(q, Q, K, N, 0, ids, d2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): e.knn_batch_dev(q, Q, K, N, 0, ids, d2)
    b.record(); torch.cuda.synchronize()
    print(f"flags {fl}: K2+K3 {a.elapsed_time(b) / 20 * 1000:.1f} us", flush=True)
