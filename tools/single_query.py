"""Latency of the reference-shaped single-query calls (detectInterLoopClosureID / detectIntraLoopClosureID) on a large database."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_slam_b200 import synth, engine
dev = torch.device("cuda:0"); N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
e = engine.ScanContextB200(numCandidates=10); e.reserve(N)
for c0 in range(0, N, 1 << 17): e.insert_batch_dev(synth.desc_db(min(1 << 17, N - c0), device=dev, start=c0))
torch.cuda.synchronize()
for name, fn in (("detectInterLoopClosureID", e.detectInterLoopClosureID), ("detectIntraLoopClosureID", e.detectIntraLoopClosureID)):
    for _ in range(3): fn(N - 1)
    t0 = time.perf_counter()
    for i in range(50): r = fn(N - 1 - i)
    print(f"{name}: {(time.perf_counter() - t0) / 50 * 1e6:.0f} us per call, N = {N}, last result {r}", flush=True)
