"""Generates tests/golden/sc_golden.npz from oracle/_ref/libscl_ref.so — the reference's own
scan_context_descriptor class text (/root/reference/include/descriptor.h:1304-1801) and vendored
nanoflann compiled in this container (oracle/Makefile). Run where /root/reference exists:

    python tests/golden/make_golden.py

Inputs are stored next to the expected outputs so the fixtures do not depend on any
generator staying bit-stable. tests/test_oracle_golden.py replays them through the restatement
(oracle/liboracle.so) on CPU; tests/test_gpu_parity.py replays them through the CUDA engine.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle_lib import Oracle, build_oracle, have_ref  # noqa: E402
from scl_slam_b200 import synth  # noqa: E402


def case(tag, R, S, K, excl, n_db, out, seed):
    o = Oracle(num_ring=R, num_sector=S, num_candidates=K, num_exclude_recent=excl, kind="ref")
    db = synth.desc_db(n_db, R, S, seed=seed).numpy()
    # revisits: the last quarter of the entries are perturbed rotations of early ones
    import torch
    nq = n_db // 4
    q, src, shift = synth.desc_queries(torch.from_numpy(db[: n_db - nq - excl]), nq, seed=seed + 1)
    db[n_db - nq:] = q.numpy()
    db[5] = db[3]            # exact duplicate (ring-key tie / libnabo self-match rule)
    db[7][:] = 0.0           # empty descriptor (NaN distance path, descriptor.h:1534)
    for i in range(n_db):
        o.saveDescriptorAndKey(db[i], i % 3, i // 3)
    out[tag + "_params"] = np.array([R, S, K, excl], np.int32)
    out[tag + "_db"] = db
    out[tag + "_ring_keys"] = np.stack([o.ring_key(i) for i in range(n_db)])
    out[tag + "_sector_keys"] = np.stack([o.sector_key(i) for i in range(0, n_db, 7)])
    rng = np.random.default_rng(seed)
    pairs = rng.integers(0, n_db, size=(64, 2)).astype(np.int32)
    pairs[:4] = [[3, 5], [7, 9], [9, 7], [7, 7]]
    pairs[4:4 + min(16, nq)] = [[n_db - nq + i, int(src[i])] for i in range(min(16, nq))]
    res = [o.distance(int(a), int(b)) for a, b in pairs]
    out[tag + "_pairs"] = pairs
    out[tag + "_pair_dist"] = np.array([r[0] for r in res], np.float64)
    out[tag + "_pair_shift"] = np.array([r[1] for r in res], np.int32)
    out[tag + "_pair_align"] = np.array([o.fast_align(int(a), int(b)) for a, b in pairs], np.int32)
    intra = [o.detectIntraLoopClosureID(i) for i in range(n_db)]
    inter = [o.detectInterLoopClosureID(i) for i in range(n_db)]
    out[tag + "_intra_id"] = np.array([r[0] for r in intra], np.int32)
    out[tag + "_intra_second"] = np.array([r[1] for r in intra], np.float32)
    out[tag + "_inter_id"] = np.array([r[0] for r in inter], np.int32)
    out[tag + "_inter_second"] = np.array([r[1] for r in inter], np.float32)
    knn_q = np.arange(n_db - nq, n_db, max(1, nq // 8)).astype(np.int32)
    ids, d2 = [], []
    for c in knn_q:
        f, i_, d_ = o.knn(int(c), n_db - nq - excl, 10, 0)
        ids.append(i_); d2.append(d_)
    out[tag + "_knn_q"] = knn_q
    out[tag + "_knn_ndb"] = np.int32(n_db - nq - excl)
    out[tag + "_knn_ids"] = np.stack(ids)
    out[tag + "_knn_d2"] = np.stack(d2)


def main():
    build_oracle()
    assert have_ref(), "oracle/_ref/libscl_ref.so missing: /root/reference not available"
    out = {}
    # descriptor build (polar binning) on three small clouds incl. edge-case points
    world = synth.make_world(1, 300)
    traj = synth.trajectory(40, seed=1)
    for ci, (kind, n_az) in enumerate([("vlp16", 200), ("hdl64", 60), ("livox", 3000)]):
        pts = synth.to_pcl_xyzi(synth.scan(world, traj[ci * 7], synth.lidar_dirs(kind, n_az=n_az), seed=ci))
        edge = np.zeros((12, 8), np.float32)
        edge[:, :3] = [[0, 0, 1], [80, 0, 2], [0, 80, 2], [-80, 0, 3], [0, -80, 3], [56.5685, 56.5685, 4],
                       [1e-30, 1e-30, 5], [3, 0, -1000 - 1.65], [3, 0.0001, -2000], [-0.0, 5, 1], [5, -0.0, 1],
                       [79.99999, -0.00001, 6]]
        pts = np.concatenate([pts, edge]).astype(np.float32)
        for (R, S) in [(20, 60), (40, 120)]:
            o = Oracle(num_ring=R, num_sector=S, kind="ref")
            out[f"cloud{ci}_desc_{R}x{S}"] = o.make_scancontext(pts)
        out[f"cloud{ci}_pts"] = pts
    case("a", 20, 60, 3, 30, 150, out, seed=11)
    case("b", 20, 60, 10, 100, 260, out, seed=12)
    case("c", 40, 120, 10, 20, 72, out, seed=13)
    np.savez_compressed(os.path.join(HERE, "sc_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "sc_golden.npz"), sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    main()
