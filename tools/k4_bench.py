"""K4 alone on the C3 shape (1024 queries x 10 candidates from a 262,144-entry database): stage timing with CUDA events,
both variants. Used for the ncu capture of scdist_kernel (profiles/)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scl_slam_b200 import build, engine, synth  # noqa: E402

build.build()
dev = torch.device("cuda", 0)
N, Q, K = 1 << 18, 1024, 10
e = engine.ScanContextB200(numCandidates=K)
e.set_stream(torch.cuda.current_stream().cuda_stream)
e.reserve(N)
for c0 in range(0, N, 1 << 16):
    e.insert_batch_dev(synth.desc_db(1 << 16, seed=3, device=dev, start=c0))
q, src, shift = synth.desc_queries(synth.desc_db(1 << 16, seed=3, device=dev), Q, seed=4)
out = dict(cand_ids=torch.empty((Q, K), dtype=torch.int32, device=dev), cand_dist=torch.empty((Q, K), dtype=torch.float64, device=dev),
           cand_shift=torch.empty((Q, K), dtype=torch.int32, device=dev), best_id=torch.empty(Q, dtype=torch.int32, device=dev),
           best_shift=torch.empty(Q, dtype=torch.int32, device=dev))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for mode in (0, 1):
    e.set_scdist_mode(mode)
    for _ in range(3):
        e.query_batch_dev(q, None, Q, K, N, 0, out)
    torch.cuda.synchronize()
    e.set_profiling(True)
    for _ in range(iters):
        flush.zero_()
        e.query_batch_dev(q, None, Q, K, N, 0, out)
    torch.cuda.synchronize()
    e.set_profiling(False)
    t = {s: e.stage_time(s) for s in (0, 1, 2)}
    ok = float((out["best_id"].cpu() == src.cpu()).float().mean())
    print(f"mode {mode}: K2 {t[0][0] / t[0][1] * 1e3:.1f} us, K3 {t[1][0] / t[1][1] * 1e3:.1f} us, K4 {t[2][0] / t[2][1] * 1e3:.1f} us; sources recovered {ok:.3f}")
