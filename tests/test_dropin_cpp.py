"""The C++ drop-in class (include/descriptor_b200.h) compiles against the reference's abstract
interface and links to the C-ABI library (CPU check); on a GPU it is driven through
unique_ptr<scan_descriptor> like distributed_mapping does and compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "dropin_main")


def _compile():
    from scl_slam_b200 import build
    build.build()
    pkg = os.path.join(ROOT, "scl_slam_b200")
    subprocess.run(["g++", "-std=c++14", "-O2", "-Wall", os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"), "-o", EXE,
                    "-L" + pkg, "-lscl_b200", "-Wl,-rpath," + pkg, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"],
                   check=True)


def test_dropin_class_compiles_and_links():
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
@pytest.mark.parametrize("n_gpus", [0, 1, 2])
def test_dropin_class_runs_like_the_reference(tmp_path, n_gpus):
    """n_gpus = 0: scan_context_descriptor_b200; 1, 2: scan_context_descriptor_b200_sharded over that many devices."""
    import torch
    from oracle_lib import Oracle
    if n_gpus > torch.cuda.device_count():
        pytest.skip("not enough GPUs")
    from scl_slam_b200 import synth
    _compile()
    world = synth.make_world(6, 200)
    traj = synth.trajectory(20, seed=6)
    dirs = synth.lidar_dirs("vlp16", n_az=240)
    clouds = [synth.to_pcl_xyzi(synth.scan(world, traj[i], dirs, seed=i)) for i in range(6)]
    db = synth.desc_db(150, seed=61)
    q, _, _ = synth.desc_queries(db[:80], 40, seed=62)
    wires = np.concatenate([db.numpy()[:110], q.numpy()]).reshape(150, -1)
    path = tmp_path / "scenario.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("iii", len(clouds), wires.shape[0], wires.shape[1]))
        for c in clouds:
            f.write(struct.pack("i", c.shape[0])); f.write(c.tobytes())
        f.write(wires.astype(np.float32).tobytes())
    out = subprocess.run([EXE, str(path)] + ([str(n_gpus)] if n_gpus else []), check=True, capture_output=True, text=True).stdout.splitlines()
    o = Oracle(num_candidates=10, num_exclude_recent=30)
    sums = [float(np.sum(o.makeAndSaveDescriptorAndKey(c, 0, i).astype(np.float64))) for i, c in enumerate(clouds)]
    for i in range(wires.shape[0]):
        o.saveDescriptorAndKey(wires[i], 1, i)
    builds = [l.split() for l in out if l.startswith("build")]
    assert len(builds) == 6
    for b, s in zip(builds, sums):
        assert int(b[2]) == 1200 and abs(float(b[3]) - s) <= 1e-6 * max(1.0, abs(s))
    assert out[6] == f"size {o.getSize()}"
    hits = 0
    for l in [l.split() for l in out if l.startswith("query")]:
        cur = int(l[1])
        a, b = o.detectIntraLoopClosureID(cur), o.detectInterLoopClosureID(cur)
        assert (int(l[2]), np.float32(l[3])) == (a[0], np.float32(a[1])), cur
        assert (int(l[4]), np.float32(l[5])) == (b[0], np.float32(b[1])), cur
        hits += a[0] >= 0
    assert hits > 10
    assert out[-1] == "index 1 0 -1 -1"
