"""Timing experiments on the tensor-core kNN kernel (SCL_TC_FLAGS disables parts of it; results are then wrong on purpose)."""
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
for _ in range(2): e.knn_batch_dev(q, Q, K, N, 0, ids, d2)
torch.cuda.synchronize()
for fl in (0, 1, 3, 7, 5, 4, 8, 9, 15):
    os.environ["SCL_TC_FLAGS"] = str(fl)
    sys.stderr.write(f"flags {fl}: "); sys.stderr.flush()
    e.knn_batch_dev(q, Q, K, N, 0, ids, d2); torch.cuda.synchronize()
