"""Bring-up check of the peer-memory exchange (csrc/k7_exchange.cu) against the NCCL form: torchrun --nproc-per-node N tools/xchg_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from scl_slam_b200 import synth, engine
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
Q, K, n = 256, 10, 60000
e = engine.ScanContextB200(numCandidates=K, device=lr); e.set_stream(torch.cuda.current_stream().cuda_stream); e.set_shard(rank, world)
full = synth.desc_db(n * world, seed=3, device=dev)
e.insert_batch_dev(full[rank::world].contiguous())
q = synth.desc_queries(full[: 1 << 14], Q, seed=4)[0]
QK = Q * K
blob1 = torch.empty(QK * 8, dtype=torch.uint8, device=dev); blob2 = torch.empty(QK * 12, dtype=torch.uint8, device=dev)
loc_ids, loc_d2 = blob1[:QK * 4].view(torch.int32).view(Q, K), blob1[QK * 4:].view(torch.float32).view(Q, K)
own_dist, own_shift = blob2[:QK * 8].view(torch.float64).view(Q, K), blob2[QK * 8:].view(torch.int32).view(Q, K)
g1 = torch.empty((world, QK * 8), dtype=torch.uint8, device=dev); g2 = torch.empty((world, QK * 12), dtype=torch.uint8, device=dev)
def buf():
    return dict(cand_ids=torch.empty((Q, K), dtype=torch.int32, device=dev), cand_d2=torch.empty((Q, K), dtype=torch.float32, device=dev),
                cand_dist=torch.empty((Q, K), dtype=torch.float64, device=dev), cand_shift=torch.empty((Q, K), dtype=torch.int32, device=dev),
                best_id=torch.empty(Q, dtype=torch.int32, device=dev), best_dist=torch.empty(Q, dtype=torch.float64, device=dev),
                best_shift=torch.empty(Q, dtype=torch.int32, device=dev))
a, b = buf(), buf()
def step(p2p, out, seq):
    e.knn_batch_dev(q, Q, K, e.getSize(), 0, loc_ids, loc_d2); torch.cuda.synchronize(); print(rank, "knn ok", flush=True)
    if p2p: e.xchg_merge_topk_dev(seq, Q, K, blob1, out["cand_ids"], out["cand_d2"])
    else:
        dist.all_gather_into_tensor(g1, blob1); e.merge_topk_dev(world, Q, K, g1, g1[:, QK * 4:], QK * 8, out["cand_ids"], out["cand_d2"])
    torch.cuda.synchronize(); print(rank, "topk ok", flush=True)
    e.scdist_owned_dev(q, Q, K, out["cand_ids"], own_dist, own_shift); torch.cuda.synchronize(); print(rank, "k4 ok", flush=True)
    if p2p: e.xchg_combine_dev(seq, Q, K, blob2, out["cand_ids"], out)
    else:
        dist.all_gather_into_tensor(g2, blob2); e.combine_owned_dev(world, Q, K, out["cand_ids"], g2, g2[:, QK * 8:], QK * 12, out)
    torch.cuda.synchronize(); print(rank, "combine ok", flush=True)
step(False, a, 0)
h = e.xchg_create(world, QK); hs = [None] * world; dist.all_gather_object(hs, h); e.xchg_open(world, rank, hs)
print(rank, "opened", flush=True)
for s in (1, 2, 3):
    step(True, b, s)
    print(rank, "step", s, {k: bool(torch.equal(a[k].view(torch.uint8), b[k].view(torch.uint8))) for k in a}, flush=True)
e.xchg_close()
dist.destroy_process_group()
