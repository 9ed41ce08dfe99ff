"""Generates tests/golden/rowkey_golden.npz from the reference's OWN text of the row-key candidate stage of
lidar_iris_descriptor (oracle/_ref/libiris_ref.so: descriptor.h:1047-1063, 1087-1267 cut at build time; oracle/Makefile).
Run in the container that has /root/reference:  python tests/golden/make_golden_rowkey.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rowkey_scenario as sc                      # noqa: E402
from oracle_lib import IrisOracle, build_oracle, have_iris_ref   # noqa: E402

PARAMS = dict(rows=80, num_exclude_recent=30, num_candidates=10, dist_thres=0.32)

if __name__ == "__main__":
    build_oracle()
    assert have_iris_ref(), "needs /root/reference (oracle/_ref/libiris_ref.so)"
    out = {}
    for tag, (seed, this_id) in {"a": (11, 0), "b": (12, 1)}.items():
        saves = sc.make(seed)
        ref = sc.run(lambda **kw: IrisOracle(kind="ref", **kw), saves, 3, this_id, **PARAMS)
        out[tag + "_meta"] = np.array([seed, this_id, 3], np.int32)
        out[tag + "_keys"] = np.stack([s[0] for s in saves])
        out[tag + "_robot"] = np.array([s[1] for s in saves], np.int32)
        out[tag + "_idx"] = np.array([s[2] for s in saves], np.int32)
        out[tag + "_feat"] = np.array([s[3] for s in saves], np.float32)
        for part in ("intra", "inter"):
            for k, v in ref[part].items():
                out[f"{tag}_{part}_{k}"] = v
        out[tag + "_index"] = ref["index"]
        out[tag + "_sizes"] = ref["sizes"]
    path = os.path.join(ROOT, "tests", "golden", "rowkey_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: int((out[k] >= 0).sum()) for k in out if k.endswith("_id")})
