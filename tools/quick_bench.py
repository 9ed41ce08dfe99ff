"""Scratch timing of the individual stages on one B200 (not the contract bench; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_slam_b200 import synth, engine

def ev_time(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

def main():
    dev = torch.device("cuda:0")
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    Q, K = 1024, 10
    e = engine.ScanContextB200(numCandidates=K)
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    e.reserve(N)
    t = time.time()
    for c0 in range(0, N, 1 << 17):
        m = min(1 << 17, N - c0)
        e.insert_batch_dev(synth.desc_db(m, device=dev, start=c0))
    torch.cuda.synchronize(); print("db fill s", time.time() - t, "N", e.getSize())
    db_head = synth.desc_db(1 << 16, device=dev)
    q, src, shift = synth.desc_queries(db_head, Q)
    out = dict(cand_ids=torch.empty((Q, K), dtype=torch.int32, device=dev), cand_d2=torch.empty((Q, K), device=dev),
               cand_dist=torch.empty((Q, K), dtype=torch.float64, device=dev), cand_shift=torch.empty((Q, K), dtype=torch.int32, device=dev),
               best_id=torch.empty(Q, dtype=torch.int32, device=dev), best_dist=torch.empty(Q, dtype=torch.float64, device=dev),
               best_shift=torch.empty(Q, dtype=torch.int32, device=dev))
    for n_db in [20000, 131072, N]:
        if n_db > N: continue
        ms = ev_time(lambda: e.query_batch_dev(q, None, Q, K, n_db, 0, out))
        acc = (out["best_id"].cpu().numpy() == src.cpu().numpy()).mean() if n_db >= (1 << 16) else float('nan')
        bytes_alg = 80 * n_db + 52892 * Q
        print(f"query n_db={n_db} ms={ms:.3f} q/s={Q/ms*1e3:.0f} GB/s(alg)={bytes_alg/ms/1e6:.1f} recovered={acc:.3f}")
    # descriptor build
    world = synth.make_world(1, 300)
    sc = synth.to_pcl_xyzi(synth.scan(world, (0, 0, 0), synth.lidar_dirs("hdl64"), seed=0))
    B = 64
    pts = torch.from_numpy(np.concatenate([sc] * B)).to(dev)
    offs = np.arange(B + 1, dtype=np.int32) * sc.shape[0]
    outd = torch.empty((B, 20, 60), device=dev)
    e2 = engine.ScanContextB200()
    e2.set_stream(torch.cuda.current_stream().cuda_stream)
    ms = ev_time(lambda: e2.build_batch_dev(pts, offs, 32, insert=False, out_dev=outd), iters=10)
    print(f"build B={B} P={sc.shape[0]} ms={ms:.3f} desc/s={B/ms*1e3:.0f} GB/s(32B/pt)={B*sc.shape[0]*32/ms/1e6:.1f}")

main()
