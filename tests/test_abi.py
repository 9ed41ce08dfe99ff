"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/scl_engine.h, include/scl_wire.h and include/scl_rowkey.h declare; without a GPU the engine refuses to come up
(no CPU fallback). No compute calls here."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from scl_slam_b200 import build, engine
    build.build()
    return engine.load_library()


def test_header_symbols_all_exported(lib):
    from scl_slam_b200 import engine
    from scl_slam_b200 import rowkey
    hdr = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("scl_engine.h", "scl_wire.h", "scl_rowkey.h"))
    declared = set(re.findall(r"\b(scl_[a-z_0-9]+)\s*\(", hdr)) - {"scl_rowkey_compare_fn"}
    assert declared == set(engine.EXPORTS) | set(rowkey.EXPORTS), declared ^ (set(engine.EXPORTS) | set(rowkey.EXPORTS))
    for name in declared:
        assert hasattr(lib, name), name


def test_every_entry_point_is_in_the_integration_map():
    """INTEGRATION.md §2 maps each C-ABI entry point to the reference member it stands in for."""
    hdr = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("scl_engine.h", "scl_wire.h", "scl_rowkey.h"))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = sorted(n for n in set(re.findall(r"\b(scl_[a-z_0-9]+)\s*\(", hdr)) if n not in doc)
    assert not missing, missing


def test_default_params_match_reference_ctor(lib):
    """descriptor.h:1307-1316"""
    from scl_slam_b200.engine import SclParams
    p = SclParams()
    lib.scl_default_params(C.byref(p))
    assert (p.num_ring, p.num_sector, p.num_candidates, p.num_exclude_recent, p.tree_making_period) == (20, 60, 3, 100, 10)
    assert (p.dist_thres, p.lidar_height, p.max_radius, p.search_ratio) == (0.14, 1.65, 80.0, 0.1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    from scl_slam_b200.engine import ScanContextB200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ScanContextB200()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under scl_slam_b200/ may reference it."""
    pkg = os.path.join(ROOT, "scl_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in src and "liboracle" not in src and "sco_" not in src, f


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: the three public headers compile as C99 on their own (no C++ types, no torch types)."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "scl_engine.h"\n#include "scl_wire.h"\n#include "scl_rowkey.h"\n'
                   "int main(void) { scl_params p; scl_rowkey_params r; (void)p; (void)r; return (int)sizeof(scl_loop_info) == 0; }\n")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], check=True)
