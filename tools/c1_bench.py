"""C1 arm of bench.py alone (per-keyframe build + intra + inter query through the reference's call sequence), then the
same loop with the three calls timed separately.
usage: python tools/c1_bench.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from scl_slam_b200 import build, engine, synth  # noqa: E402

if __name__ == "__main__":
    build.build()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    res = bench.arm_c1(engine, synth, dev, bench.peaks(), cpu=False)
    world = synth.make_world(1, 600)
    traj = synth.trajectory(400, seed=1)
    dirs = synth.lidar_dirs("vlp16", n_az=1800)
    pts, off = synth.scan_batch_torch(world, traj, dirs, seed=0, device=dev)
    h = pts.cpu().numpy()
    clouds = [h[off[i]:off[i + 1]] for i in range(len(off) - 1)]
    e = engine.ScanContextB200(numCandidates=10)
    t = [0.0, 0.0, 0.0]
    for i, c in enumerate(clouds):
        t0 = time.perf_counter(); e.makeAndSaveDescriptorAndKey(c, 0, i)
        t1 = time.perf_counter(); e.detectIntraLoopClosureID(i)
        t2 = time.perf_counter(); e.detectInterLoopClosureID(i)
        t3 = time.perf_counter()
        t[0] += t1 - t0; t[1] += t2 - t1; t[2] += t3 - t2
    res["per_call_us"] = {"makeAndSaveDescriptorAndKey": t[0] / len(clouds) * 1e6, "detectIntraLoopClosureID": t[1] / len(clouds) * 1e6,
                          "detectInterLoopClosureID": t[2] / len(clouds) * 1e6}
    print(json.dumps(res))
