"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of oracle/_ref/libscl_ref.so.

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs only. The product package (scl_slam_b200/) never imports this module.

Both libraries export the same sco_* entry points (oracle/sc_oracle.h):
  kind="port": our restatement of /root/reference/include/descriptor.h:1304-1801
  kind="ref" : the reference's own class text + vendored nanoflann, built by oracle/Makefile
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libscl_ref.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build_oracle(quiet=True):
    """make -C oracle (liboracle.so always; _ref only where /root/reference exists)."""
    subprocess.run(["make", "-C", ORACLE_DIR], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have_ref():
    return os.path.exists(REF_SO)


def _load(path):
    lib = C.CDLL(path)
    lib.sco_create.restype = C.c_void_p
    lib.sco_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                               C.c_int, C.c_int, C.c_double]
    lib.sco_destroy.argtypes = [C.c_void_p]
    lib.sco_make_scancontext.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
    lib.sco_make_and_save.restype = C.c_int
    lib.sco_make_and_save.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int8, C.c_int, C.c_void_p]
    lib.sco_save.restype = C.c_int
    lib.sco_save.argtypes = [C.c_void_p, _f32p, C.c_int8, C.c_int]
    lib.sco_bulk_load.restype = C.c_int
    lib.sco_bulk_load.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.sco_size.restype = C.c_int
    lib.sco_size.argtypes = [C.c_void_p]
    lib.sco_get_index.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.sco_get_desc.argtypes = [C.c_void_p, C.c_int, _f32p]
    lib.sco_ring_key.argtypes = [C.c_void_p, C.c_int, _f32p]
    lib.sco_sector_key.argtypes = [C.c_void_p, C.c_int, _f64p]
    lib.sco_distance.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.sco_distance_raw.argtypes = [C.c_void_p, _f32p, _f32p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.sco_fast_align.restype = C.c_int
    lib.sco_fast_align.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.sco_dist_direct.restype = C.c_double
    lib.sco_dist_direct.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.sco_detect_intra.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    lib.sco_detect_inter.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    lib.sco_knn.restype = C.c_int
    lib.sco_knn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _f32p]
    lib.sco_query_batch.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    _i32p, _f32p, _f64p, _i32p, _i32p, _f64p, _i32p]
    lib.sco_atanf_libm.restype = C.c_float
    lib.sco_atanf_libm.argtypes = [C.c_float]
    lib.sco_atanf_port.restype = C.c_float
    lib.sco_atanf_port.argtypes = [C.c_float]
    if hasattr(lib, "sco_icp"):
        lib.sco_icp.restype = C.c_int
        lib.sco_icp.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, C.c_double, C.c_int,
                                C.c_double, C.c_double, _f32p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        lib.sco_verify_ransac.restype = C.c_int
        lib.sco_verify_ransac.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint, _f32p,
                                          C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.sco_nn.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, _i32p, _f32p]
        lib.sco_voxel_grid.restype = C.c_int
        lib.sco_voxel_grid.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _f32p]
    if hasattr(lib, "sco_voxel_grid_pcl"):
        lib.sco_voxel_grid_pcl.restype = C.c_int
        lib.sco_voxel_grid_pcl.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _f32p]
        lib.sco_assemble_submap.restype = C.c_int
        lib.sco_assemble_submap.argtypes = [_f32p, _i32p, C.c_int, C.c_int, _f32p, C.c_float, _f32p]
    return lib


_LIBS = {}


def get_lib(kind="port"):
    if kind not in _LIBS:
        path = PORT_SO if kind == "port" else REF_SO
        if not os.path.exists(path):
            build_oracle()
        _LIBS[kind] = _load(path)
    return _LIBS[kind]


def _cloud(pts):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    assert pts.ndim == 2 and pts.shape[1] >= 3
    return pts, pts.shape[0], pts.shape[1]


class Oracle:
    """Mirror of scan_context_descriptor's public interface (descriptor.h:1304-1801)."""

    def __init__(self, num_ring=20, num_sector=60, num_candidates=3, dist_thres=0.14,
                 lidar_height=1.65, max_radius=80.0, num_exclude_recent=100,
                 tree_making_period=10, search_ratio=0.1, kind="port"):
        self.lib = get_lib(kind)
        self.kind = kind
        self.R, self.S, self.K = num_ring, num_sector, num_candidates
        self.h = self.lib.sco_create(num_ring, num_sector, num_candidates, dist_thres, lidar_height,
                                     max_radius, num_exclude_recent, tree_making_period, search_ratio)
        self._keep = []

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.sco_destroy(self.h)
            self.h = None

    # -- descriptor build -------------------------------------------------------------------
    def make_scancontext(self, pts, want_bins=False):
        pts, n, stride = _cloud(pts)
        desc = np.empty(self.R * self.S, np.float32)
        ring = np.zeros(n, np.int32) if want_bins else None
        sector = np.zeros(n, np.int32) if want_bins else None
        self.lib.sco_make_scancontext(self.h, pts.ctypes.data, n, stride, desc.ctypes.data,
                                      ring.ctypes.data if want_bins else None,
                                      sector.ctypes.data if want_bins else None)
        desc = desc.reshape(self.R, self.S)
        return (desc, ring, sector) if want_bins else desc

    def makeAndSaveDescriptorAndKey(self, pts, robot, index):
        pts, n, stride = _cloud(pts)
        desc = np.empty(self.R * self.S, np.float32)
        self.lib.sco_make_and_save(self.h, pts.ctypes.data, n, stride, robot, index, desc.ctypes.data)
        return desc

    def saveDescriptorAndKey(self, desc, robot, index):
        desc = np.ascontiguousarray(desc, np.float32).reshape(-1)
        assert desc.size == self.R * self.S
        return self.lib.sco_save(self.h, desc, robot, index)

    def bulk_load(self, wires, borrow=True):
        wires = np.ascontiguousarray(wires, np.float32).reshape(-1, self.R * self.S)
        if borrow:
            self._keep.append(wires)
        return self.lib.sco_bulk_load(self.h, wires.ctypes.data, wires.shape[0], 1 if borrow else 0)

    # -- accessors ---------------------------------------------------------------------------
    def getSize(self, id_in=-1):
        return self.lib.sco_size(self.h)

    def getIndex(self, key):
        r, i = C.c_int(), C.c_int()
        self.lib.sco_get_index(self.h, key, C.byref(r), C.byref(i))
        return r.value, i.value

    def desc(self, key):
        out = np.empty(self.R * self.S, np.float32)
        self.lib.sco_get_desc(self.h, key, out)
        return out.reshape(self.R, self.S)

    def ring_key(self, key):
        out = np.empty(self.R, np.float32)
        self.lib.sco_ring_key(self.h, key, out)
        return out

    def sector_key(self, key):
        out = np.empty(self.S, np.float64)
        self.lib.sco_sector_key(self.h, key, out)
        return out

    # -- distance ----------------------------------------------------------------------------
    def distance(self, k1, k2):
        d, s = C.c_double(), C.c_int()
        self.lib.sco_distance(self.h, k1, k2, C.byref(d), C.byref(s))
        return d.value, s.value

    def distance_raw(self, d1, d2):
        d, s = C.c_double(), C.c_int()
        self.lib.sco_distance_raw(self.h, np.ascontiguousarray(d1, np.float32).reshape(-1),
                                  np.ascontiguousarray(d2, np.float32).reshape(-1), C.byref(d), C.byref(s))
        return d.value, s.value

    def fast_align(self, k1, k2):
        return self.lib.sco_fast_align(self.h, k1, k2)

    def dist_direct(self, k1, k2, shift):
        return self.lib.sco_dist_direct(self.h, k1, k2, shift)

    # -- queries -----------------------------------------------------------------------------
    def detectIntraLoopClosureID(self, cur):
        i, f = C.c_int(), C.c_float()
        self.lib.sco_detect_intra(self.h, cur, C.byref(i), C.byref(f))
        return i.value, f.value

    def detectInterLoopClosureID(self, cur):
        i, f = C.c_int(), C.c_float()
        self.lib.sco_detect_inter(self.h, cur, C.byref(i), C.byref(f))
        return i.value, f.value

    def knn(self, cur, n_db, k, metric=0):
        ids = np.empty(k, np.int32)
        d2 = np.empty(k, np.float32)
        found = self.lib.sco_knn(self.h, cur, n_db, k, metric, ids, d2)
        return found, ids, d2

    def query_batch(self, queries, n_db, k, metric=0, nthreads=1):
        q = np.ascontiguousarray(queries, np.int32)
        nq = q.size
        out = dict(cand_ids=np.empty((nq, k), np.int32), cand_d2=np.empty((nq, k), np.float32),
                   cand_dist=np.empty((nq, k), np.float64), cand_shift=np.empty((nq, k), np.int32),
                   best_id=np.empty(nq, np.int32), best_dist=np.empty(nq, np.float64),
                   best_shift=np.empty(nq, np.int32))
        self.lib.sco_query_batch(self.h, q, nq, n_db, k, metric, nthreads, out["cand_ids"], out["cand_d2"],
                                 out["cand_dist"], out["cand_shift"], out["best_id"], out["best_dist"],
                                 out["best_shift"])
        return out


def icp(src, tgt, max_corr_dist=100.0, max_iter=50, trans_eps=1e-6, fit_eps=1e-6):
    """PCL-default ICP restatement (distributedMapping.h:1108-1132). Returns (T 4x4, fitness, converged, iters)."""
    lib = get_lib("port")
    src, ns, st = _cloud(src)
    tgt, nt, st2 = _cloud(tgt)
    assert st == st2
    T = np.empty(16, np.float32)
    fit, conv = C.c_float(), C.c_int()
    it = lib.sco_icp(src.reshape(-1), ns, tgt.reshape(-1), nt, st, max_corr_dist, max_iter, trans_eps, fit_eps,
                     T, C.byref(fit), C.byref(conv))
    return T.reshape(4, 4), fit.value, bool(conv.value), it


def verify_ransac(src, tgt, max_iter=1000, inlier_thr=0.25, min_inlier_ratio=0.45, seed=1):
    """RANSAC + SVD restatement (distributedMapping.h:1211-1243). Returns (T, n_corr, n_inliers, success)."""
    lib = get_lib("port")
    src, ns, st = _cloud(src)
    tgt, nt, st2 = _cloud(tgt)
    assert st == st2
    T = np.empty(16, np.float32)
    nc, ni, ok = C.c_int(), C.c_int(), C.c_int()
    lib.sco_verify_ransac(src.reshape(-1), ns, tgt.reshape(-1), nt, st, max_iter, inlier_thr, min_inlier_ratio, seed, T,
                          C.byref(nc), C.byref(ni), C.byref(ok))
    return T.reshape(4, 4), nc.value, ni.value, bool(ok.value)


def nn_bruteforce(src, tgt):
    lib = get_lib("port")
    src, ns, st = _cloud(src)
    tgt, nt, st2 = _cloud(tgt)
    idx = np.empty(ns, np.int32)
    d2 = np.empty(ns, np.float32)
    lib.sco_nn(src.reshape(-1), ns, tgt.reshape(-1), nt, st, idx, d2)
    return idx, d2


def voxel_grid(pts, leaf):
    lib = get_lib("port")
    pts, n, st = _cloud(pts)
    out = np.empty((max(n, 1), 3), np.float32)
    m = lib.sco_voxel_grid(pts.reshape(-1), n, st, leaf, out.reshape(-1))
    return out[:m].copy()


def voxel_grid_pcl(pts, leaf):
    """pcl::VoxelGrid<PointXYZI> restatement: (m, 4) float32 centroids (x, y, z, intensity) in leaf-index order."""
    lib = get_lib("port")
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    assert pts.ndim == 2 and pts.shape[1] >= 4
    n, st = pts.shape
    out = np.empty((max(n, 1), 4), np.float32)
    m = lib.sco_voxel_grid_pcl(pts.reshape(-1), n, st, leaf, out.reshape(-1))
    return out[:m].copy()


def assemble_submap(clouds, poses6, leaf):
    """loopFindNearKeyframes restatement: clouds = list of (n_i, >=4) arrays sharing a stride, poses6 = (len, 6)."""
    lib = get_lib("port")
    st = clouds[0].shape[1]
    pts = np.ascontiguousarray(np.concatenate(clouds), dtype=np.float32)
    offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    poses = np.ascontiguousarray(poses6, dtype=np.float32)
    out = np.empty((max(pts.shape[0], 1), 4), np.float32)
    m = lib.sco_assemble_submap(pts.reshape(-1), offs, len(clouds), st, poses.reshape(-1), leaf, out.reshape(-1))
    return out[:m].copy()


# ---- row-key candidate stage of lidar_iris_descriptor (oracle/rowkey_oracle.h) ------------------------------------------
IRIS_REF_SO = os.path.join(ORACLE_DIR, "_ref", "libiris_ref.so")
_iris_libs = {}


def have_iris_ref():
    return os.path.exists(IRIS_REF_SO)


def _load_iris(kind):
    if kind not in _iris_libs:
        lib = C.CDLL(PORT_SO if kind == "port" else IRIS_REF_SO)
        lib.sco_iris_create.restype = C.c_void_p
        lib.sco_iris_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
        lib.sco_iris_destroy.argtypes = [C.c_void_p]
        lib.sco_iris_save.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_float]
        for f in (lib.sco_iris_detect_intra, lib.sco_iris_detect_inter):
            f.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_int), _i32p, _f32p]
        lib.sco_iris_get_index.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.sco_iris_size.restype = C.c_int
        lib.sco_iris_size.argtypes = [C.c_void_p, C.c_int]
        lib.sco_iris_knn_batch.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _f32p]
        _iris_libs[kind] = lib
    return _iris_libs[kind]


def iris_compare(feat_a, tag_a, feat_b, tag_b):
    """The stand-in compare() both oracle libraries use (rowkey_oracle.h): float |fa - fb|, bias (7 ta + 13 tb) % 360."""
    return float(np.abs(np.float32(feat_a) - np.float32(feat_b))), (7 * int(tag_a) + 13 * int(tag_b)) % 360


class IrisOracle:
    """kind="port": oracle/rowkey_oracle.cpp; kind="ref": the reference's own text (oracle/_ref/libiris_ref.so)."""

    def __init__(self, rows=80, num_exclude_recent=30, num_candidates=10, dist_thres=0.32, robot_num=1, this_id=0, kind="port"):
        self.lib = _load_iris(kind)
        self.K, self.rows = num_candidates, rows
        self.h = self.lib.sco_iris_create(rows, num_exclude_recent, num_candidates, dist_thres, robot_num, this_id)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.sco_iris_destroy(self.h)
            self.h = None

    def save(self, key, robot, index, feature):
        self.lib.sco_iris_save(self.h, np.ascontiguousarray(key, np.float32), robot, index, float(feature))

    def _detect(self, fn, cur):
        i, b, n = C.c_int(), C.c_float(), C.c_int()
        cand = np.full(self.K, -1, np.int32); d2 = np.full(self.K, np.inf, np.float32)
        fn(self.h, cur, C.byref(i), C.byref(b), C.byref(n), cand, d2)
        return i.value, b.value, n.value, cand, d2

    def detect_intra(self, cur):
        return self._detect(self.lib.sco_iris_detect_intra, cur)

    def detect_inter(self, cur):
        return self._detect(self.lib.sco_iris_detect_inter, cur)

    def get_index(self, key):
        r, i = C.c_int(), C.c_int()
        self.lib.sco_iris_get_index(self.h, key, C.byref(r), C.byref(i))
        return r.value, i.value

    def size(self, id_in=-1):
        return self.lib.sco_iris_size(self.h, id_in)

    def knn_batch(self, q_keys, robot, n, K, threads=1):
        q = np.ascontiguousarray(q_keys, np.float32).reshape(-1, self.rows)
        idx = np.empty((q.shape[0], K), np.int32); d2 = np.empty((q.shape[0], K), np.float32)
        self.lib.sco_iris_knn_batch(self.h, q, q.shape[0], robot, n, K, threads, idx, d2)
        return idx, d2
