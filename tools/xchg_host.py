"""Is the N > 1 step host-bound? Times the host side of `steps` sharded query steps (no synchronisation inside) and the device side."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from scl_slam_b200 import synth, engine
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
Q, K, n = 1024, 10, (1 << 20) // world
e = engine.ScanContextB200(numCandidates=K, device=lr); e.set_stream(torch.cuda.current_stream().cuda_stream); e.set_shard(rank, world); e.reserve(n)
for c0 in range(0, n, 1 << 16):
    full = synth.desc_db((1 << 16) * world, seed=3, device=dev, start=c0 * world)
    e.insert_batch_dev(full[rank::world][: 1 << 16].contiguous())
q = synth.desc_queries(synth.desc_db(1 << 16, seed=3, device=dev), Q, seed=4)[0]
QK = Q * K
blob1 = torch.empty(QK * 8, dtype=torch.uint8, device=dev); blob2 = torch.empty(QK * 12, dtype=torch.uint8, device=dev)
loc_ids, loc_d2 = blob1[:QK * 4].view(torch.int32).view(Q, K), blob1[QK * 4:].view(torch.float32).view(Q, K)
own_dist, own_shift = blob2[:QK * 8].view(torch.float64).view(Q, K), blob2[QK * 8:].view(torch.int32).view(Q, K)
out = dict(cand_ids=torch.empty((Q, K), dtype=torch.int32, device=dev), cand_d2=torch.empty((Q, K), dtype=torch.float32, device=dev),
           cand_dist=torch.empty((Q, K), dtype=torch.float64, device=dev), cand_shift=torch.empty((Q, K), dtype=torch.int32, device=dev),
           best_id=torch.empty(Q, dtype=torch.int32, device=dev), best_dist=torch.empty(Q, dtype=torch.float64, device=dev),
           best_shift=torch.empty(Q, dtype=torch.int32, device=dev))
h = e.xchg_create(world, QK); hs = [None] * world; dist.all_gather_object(hs, h); e.xchg_open(world, rank, hs)
seq = [0]
def step():
    e.knn_batch_dev(q, Q, K, n, 0, loc_ids, loc_d2)
    seq[0] += 1
    e.xchg_merge_topk_dev(seq[0], Q, K, blob1, out["cand_ids"], out["cand_d2"])
    e.scdist_owned_dev(q, Q, K, out["cand_ids"], own_dist, own_shift)
    e.xchg_combine_dev(seq[0], Q, K, blob2, out["cand_ids"], out)
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for _ in range(50): step()
b.record(); t_host = (time.perf_counter() - t0) / 50 * 1e6
torch.cuda.synchronize()
print(f"rank {rank}: host {t_host:.1f} us per step to enqueue, device {a.elapsed_time(b) / 50 * 1e3:.1f} us per step", flush=True)
e.xchg_close(); dist.destroy_process_group()
