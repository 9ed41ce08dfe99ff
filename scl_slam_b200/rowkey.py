"""Python mirror of the row-key candidate search (include/scl_rowkey.h): the kNN stage of the reference's
class lidar_iris_descriptor (/root/reference/include/descriptor.h:1047-1059, 1087-1267) on the K3 kernels.
Method names and argument meaning follow that class; compare() stays with the caller (OpenCV features, out of scope).
There is no CPU fallback: the constructor raises without the CUDA library or a CUDA device."""
import ctypes as C

import numpy as np

from . import engine

EXPORTS = [
    "scl_rowkey_default_params", "scl_rowkey_create", "scl_rowkey_destroy", "scl_rowkey_last_error", "scl_rowkey_save", "scl_rowkey_save_batch",
    "scl_rowkey_save_wire", "scl_rowkey_intra_candidates", "scl_rowkey_inter_candidates", "scl_rowkey_detect_intra", "scl_rowkey_detect_inter",
    "scl_rowkey_get_index", "scl_rowkey_size", "scl_rowkey_knn_batch", "scl_rowkey_knn_batch_dev", "scl_rowkey_set_stream", "scl_rowkey_knn_stats",
]

COMPARE_FN = C.CFUNCTYPE(C.c_float, C.c_void_p, C.c_int8, C.c_int, C.c_int8, C.c_int, C.POINTER(C.c_int))


class SclRowkeyParams(C.Structure):
    _fields_ = [("rows", C.c_int), ("num_exclude_recent", C.c_int), ("num_candidates", C.c_int), ("dist_thres", C.c_double),
                ("robot_num", C.c_int), ("this_id", C.c_int)]


_bound = False


def _lib():
    global _bound
    lib = engine.load_library()
    if not _bound:
        lib.scl_rowkey_last_error.restype = C.c_char_p
        lib.scl_rowkey_last_error.argtypes = [C.c_void_p]
        lib.scl_rowkey_default_params.argtypes = [C.POINTER(SclRowkeyParams)]
        lib.scl_rowkey_create.argtypes = [C.POINTER(SclRowkeyParams), C.c_int, C.POINTER(C.c_void_p)]
        lib.scl_rowkey_destroy.argtypes = [C.c_void_p]
        lib.scl_rowkey_save.argtypes = [C.c_void_p, C.c_void_p, C.c_int8, C.c_int, C.POINTER(C.c_int)]
        lib.scl_rowkey_save_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int8, C.c_void_p]
        lib.scl_rowkey_save_wire.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int8, C.c_int, C.POINTER(C.c_int)]
        lib.scl_rowkey_intra_candidates.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        lib.scl_rowkey_inter_candidates.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p]
        lib.scl_rowkey_detect_intra.argtypes = [C.c_void_p, C.c_int, COMPARE_FN, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.scl_rowkey_detect_inter.argtypes = lib.scl_rowkey_detect_intra.argtypes
        lib.scl_rowkey_get_index.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int8), C.POINTER(C.c_int)]
        lib.scl_rowkey_size.argtypes = [C.c_void_p, C.c_int]
        lib.scl_rowkey_knn_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.scl_rowkey_knn_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.scl_rowkey_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.scl_rowkey_knn_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        _bound = True
    return lib


class LidarIrisRowKeysB200:
    """Constructor arguments as lidar_iris_descriptor's (descriptor.h:472-485) where they reach the candidate stage."""

    def __init__(self, rows=80, numExcludeRecent=30, numCandidates=10, distThres=0.32, robotNum=1, thisID=0, device=0):
        self.lib = _lib()
        p = SclRowkeyParams(rows, numExcludeRecent, numCandidates, distThres, robotNum, thisID)
        h = C.c_void_p()
        rc = self.lib.scl_rowkey_create(C.byref(p), device, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"scl_rowkey_create failed with status {rc}: no CUDA device or unsupported parameters (there is no CPU fallback)")
        self.h, self.p = h, p

    def close(self):
        if getattr(self, "h", None):
            self.lib.scl_rowkey_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f"scl_rowkey status {rc}: {self.lib.scl_rowkey_last_error(self.h).decode()}")

    # ---- the reference's members ----------------------------------------------------------
    def save(self, rowKey, robot, index):
        """save(), descriptor.h:1047-1059 (the key part); returns the global key."""
        k = np.ascontiguousarray(rowKey, np.float32).reshape(self.p.rows)
        g = C.c_int()
        self._ck(self.lib.scl_rowkey_save(self.h, k.ctypes.data, robot, index, C.byref(g)))
        return g.value

    def save_batch(self, rowKeys, robot, index=None):
        k = np.ascontiguousarray(rowKeys, np.float32).reshape(-1, self.p.rows)
        idx = None if index is None else np.ascontiguousarray(index, np.int32)
        self._ck(self.lib.scl_rowkey_save_batch(self.h, k.ctypes.data, k.shape[0], robot, None if idx is None else idx.ctypes.data))

    def saveDescriptorAndKey(self, iris, cols, robot, index):
        """descriptor.h:1025-1045: the wire vector; only its key part is kept."""
        w = np.ascontiguousarray(iris, np.float32).reshape(-1)
        g = C.c_int()
        self._ck(self.lib.scl_rowkey_save_wire(self.h, w.ctypes.data, cols, robot, index, C.byref(g)))
        return g.value

    def _wrap(self, compare):
        def thunk(user, ra, la, rb, lb, bias):
            d, b = compare(int(ra), int(la), int(rb), int(lb))
            bias[0] = int(b)
            return float(d)
        return COMPARE_FN(thunk)

    def detectIntraLoopClosureID(self, curPtr, compare):
        """compare(robot_a, local_a, robot_b, local_b) -> (distance, bias); returns (id, bias) like the reference's pair."""
        cb = self._wrap(compare)
        i, b, m = C.c_int(), C.c_float(), C.c_float()
        self._ck(self.lib.scl_rowkey_detect_intra(self.h, curPtr, cb, None, C.byref(i), C.byref(b), C.byref(m)))
        return i.value, b.value

    def detectInterLoopClosureID(self, curPtr, compare):
        cb = self._wrap(compare)
        i, b, m = C.c_int(), C.c_float(), C.c_float()
        self._ck(self.lib.scl_rowkey_detect_inter(self.h, curPtr, cb, None, C.byref(i), C.byref(b), C.byref(m)))
        return i.value, b.value

    def getIndex(self, key):
        r, i = C.c_int8(), C.c_int()
        self._ck(self.lib.scl_rowkey_get_index(self.h, key, C.byref(r), C.byref(i)))
        return r.value, i.value

    def getSize(self, idIn=-1):
        return self.lib.scl_rowkey_size(self.h, idIn)

    # ---- the kNN lists alone ---------------------------------------------------------------
    def intra_candidates(self, curPtr):
        K = self.p.num_candidates
        n = C.c_int()
        idx = np.full(K, -1, np.int32); d2 = np.full(K, np.inf, np.float32)
        self._ck(self.lib.scl_rowkey_intra_candidates(self.h, curPtr, C.byref(n), idx.ctypes.data, d2.ctypes.data))
        return n.value, idx, d2

    def inter_candidates(self, curPtr):
        K = self.p.num_candidates
        n = C.c_int()
        idx = np.full(K, -1, np.int32); gk = np.full(K, -1, np.int32); d2 = np.full(K, np.inf, np.float32)
        self._ck(self.lib.scl_rowkey_inter_candidates(self.h, curPtr, C.byref(n), idx.ctypes.data, gk.ctypes.data, d2.ctypes.data))
        return n.value, idx, gk, d2

    def knn_batch(self, q_keys, from_robot, n_limit=0, K=None, knn_mode=0):
        K = K or self.p.num_candidates
        q = np.ascontiguousarray(q_keys, np.float32).reshape(-1, self.p.rows)
        Q = q.shape[0]
        idx = np.empty((Q, K), np.int32); gk = np.empty((Q, K), np.int32); d2 = np.empty((Q, K), np.float32)
        self._ck(self.lib.scl_rowkey_knn_batch(self.h, q.ctypes.data, Q, from_robot, n_limit, K, knn_mode, idx.ctypes.data, gk.ctypes.data, d2.ctypes.data))
        return idx, gk, d2

    def knn_batch_dev(self, q_dev, Q, from_robot, n_limit, K, knn_mode, idx_dev, d2_dev):
        """Asynchronous on the object's stream. An object that still runs on its own (non-blocking) stream is not ordered against
        torch's: the call is then made synchronous on both sides (see engine._ordered_dev_call); bind it with set_stream to avoid that."""
        bound = getattr(self, "_bound", False)
        if not bound:
            import torch
            torch.cuda.current_stream().synchronize()
        self._ck(self.lib.scl_rowkey_knn_batch_dev(self.h, q_dev.data_ptr(), Q, from_robot, n_limit, K, knn_mode, idx_dev.data_ptr(), d2_dev.data_ptr()))
        if not bound:
            torch.cuda.synchronize()

    def set_stream(self, cuda_stream):
        self._ck(self.lib.scl_rowkey_set_stream(self.h, cuda_stream))
        self._bound = True

    def knn_stats(self):
        a, b = C.c_longlong(), C.c_longlong()
        self._ck(self.lib.scl_rowkey_knn_stats(self.h, C.byref(a), C.byref(b)))
        return {"tc_queries": a.value, "fallback_queries": b.value}
