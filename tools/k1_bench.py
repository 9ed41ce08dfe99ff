"""K1 (+K2 epilogue) alone: the descriptor arm of bench.py without the rest of the bench.
usage: python tools/k1_bench.py [batch] [iters]   ->  one JSON line (descriptors/s, ms per launch, roofline fraction)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from scl_slam_b200 import build, engine  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    build.build()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    res = bench.descriptor_bench(engine, dev, bench.peaks(), batch=batch, iters=iters, cpu=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
