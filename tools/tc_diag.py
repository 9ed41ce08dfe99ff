"""Developer aid: runs the tensor-core kNN (mode 2) on a synthetic database and prints how many queries the exact kernel had to
redo; with a library built with -DSCL_DEV_SWITCHES (SCL_B200_LIB=...) and SCL_TC_DEBUG=1 the launcher also prints why.
usage: python tools/tc_diag.py {smooth|runs|random} n_db [no_match_fraction]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scl_slam_b200 import engine, synth  # noqa: E402

kind, n = sys.argv[1], int(sys.argv[2])
nm = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
gen = {"smooth": synth.desc_db_smooth, "runs": synth.desc_db_trajectory, "random": synth.desc_db}[kind]
dev = torch.device("cuda", 0)
Q, K = 1024, 10
e = engine.ScanContextB200(numCandidates=K)
e.set_stream(torch.cuda.current_stream().cuda_stream)
e.set_knn_mode(2, True)
for c0 in range(0, n, 1 << 17):
    m = min(1 << 17, n - c0)
    e.insert_batch_dev(gen(m, 20, 60, seed=5, device=dev, start=c0))
head = gen(min(n, 1 << 16), 20, 60, seed=5, device=dev)
q, src, _ = synth.desc_queries(head, Q, seed=40)
m = int(Q * nm)
if m:
    q[:m] = synth.desc_db(m, 20, 60, seed=977, device=dev)
ids = torch.empty((Q, K), dtype=torch.int32, device=dev); d2 = torch.empty((Q, K), dtype=torch.float32, device=dev)
for rep in range(2):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); e.knn_batch_dev(q.contiguous(), Q, K, n, 0, ids, d2); b.record()
    torch.cuda.synchronize()
    print(kind, n, "no-match", nm, "ms", a.elapsed_time(b), e.knn_stats(), flush=True)
