"""Per-role cycle counters of knn_tc_kernel (SCL_TC_TIMES=1) at the bench workload."""
import os, sys
os.environ["SCL_TC_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scl_slam_b200 import synth, engine
dev = torch.device("cuda:0"); N, Q, K = 1 << 20, 1024, 10
e = engine.ScanContextB200(numCandidates=K); e.set_stream(torch.cuda.current_stream().cuda_stream); e.reserve(N)
for c0 in range(0, N, 1 << 17): e.insert_batch_dev(synth.desc_db(1 << 17, device=dev, start=c0))
q = synth.desc_queries(synth.desc_db(1 << 16, device=dev), Q)[0]
out = dict(best_id=torch.empty(Q, dtype=torch.int32, device=dev))
for _ in range(3): e.query_batch_dev(q, None, Q, K, N, 0, out)
torch.cuda.synchronize()
