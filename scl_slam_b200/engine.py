"""Python mirror of the reference's descriptor interface over the C-ABI engine.

`ScanContextB200` exposes the six `scan_descriptor` virtuals of
/root/reference/include/descriptor.h:21-36 with the reference's own names and argument meaning
(`makeAndSaveDescriptorAndKey`, `saveDescriptorAndKey`, `detectIntraLoopClosureID`,
`detectInterLoopClosureID`, `getIndex`, `getSize`) plus the batched throughput calls. Every
method is one call into scl_slam_b200/libscl_b200.so (include/scl_engine.h); torch is used only
to own device buffers and streams. There is NO CPU fallback: importing this module without the
built library, or constructing an engine without a CUDA device, raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("SCL_B200_LIB") or os.path.join(_HERE, "libscl_b200.so")   # the override is for developer builds (tools/)

SCL_OK = 0
_STATUS = {1: "SCL_ERR_INVALID", 2: "SCL_ERR_CUDA", 3: "SCL_ERR_UNSUPPORTED", 4: "SCL_ERR_RANGE", 5: "SCL_ERR_NOMEM"}

# every symbol include/scl_engine.h declares (tests check the library exports all of them)
EXPORTS = [
    "scl_default_params", "scl_default_icp_params", "scl_create", "scl_destroy", "scl_last_error", "scl_set_stream",
    "scl_reserve", "scl_set_shard", "scl_build_insert", "scl_make_scancontext", "scl_polar_tables", "scl_build_batch", "scl_build_batch_dev",
    "scl_insert", "scl_insert_batch", "scl_insert_batch_dev", "scl_get_index", "scl_size", "scl_get_descriptor",
    "scl_get_ring_key", "scl_query_intra", "scl_query_inter", "scl_query_batch", "scl_query_batch_dev",
    "scl_merge_shards_dev", "scl_icp", "scl_set_profiling", "scl_stage_time", "scl_set_knn_mode", "scl_set_scdist_mode", "scl_set_scdist_tiles", "scl_set_tc_stages", "scl_export_keys_dev", "scl_set_replicated_keys_dev", "scl_knn_stats",
    "scl_default_ransac_params", "scl_verify_ransac", "scl_knn_batch_dev", "scl_merge_topk_dev", "scl_scdist_owned_dev", "scl_combine_owned_dev",
    "scl_query_batch_submit", "scl_query_batch_wait", "scl_voxel_grid", "scl_assemble_submap", "scl_build_insert_filtered",
    "scl_xchg_create", "scl_xchg_open", "scl_xchg_merge_topk_dev", "scl_xchg_combine_dev", "scl_xchg_close",
    "scl_xchg_open_local", "scl_xchg_buffer", "scl_xchg_bytes", "scl_num_lanes", "scl_query_batch_dev_lane", "scl_lanes_fork", "scl_lanes_join",
    "scl_lane_sync", "scl_shard_query_dev", "scl_shard_query_submit", "scl_store_keyframe_cloud", "scl_keyframe_clouds", "scl_verify_intra",
    "scl_create_sharded", "scl_sharded_destroy", "scl_sharded_last_error", "scl_sharded_world", "scl_sharded_engine", "scl_sharded_size",
    "scl_sharded_get_index", "scl_sharded_get_descriptor", "scl_sharded_insert_batch", "scl_sharded_build_insert", "scl_sharded_query_batch",
    "scl_sharded_query_submit", "scl_sharded_query_wait", "scl_sharded_query_intra", "scl_sharded_query_inter",
    "scl_wire_pose6_to_transform", "scl_wire_loop_between", "scl_wire_make_loop_info", "scl_wire_make_global_descriptor",
    "scl_wire_encode_global_descriptor", "scl_wire_decode_global_descriptor", "scl_wire_encode_loop_info", "scl_wire_decode_loop_info",
]


class SclTransform(C.Structure):
    """geometry_msgs/Transform (include/scl_wire.h)"""
    _fields_ = [("tx", C.c_double), ("ty", C.c_double), ("tz", C.c_double), ("qx", C.c_double), ("qy", C.c_double), ("qz", C.c_double), ("qw", C.c_double)]

    def as_tuple(self):
        return (self.tx, self.ty, self.tz, self.qx, self.qy, self.qz, self.qw)


class SclGlobalDescriptor(C.Structure):
    """global_descriptor.msg"""
    _fields_ = [("index", C.c_int32), ("pre_pose", SclTransform), ("cur_pose", SclTransform), ("values", C.POINTER(C.c_float)), ("n_values", C.c_int32)]


class SclLoopInfo(C.Structure):
    """loop_info.msg"""
    _fields_ = [("robot0", C.c_int32), ("robot1", C.c_int32), ("index0", C.c_int32), ("index1", C.c_int32), ("noise", C.c_float), ("bet_pose", SclTransform)]


def wire_loop_between(T_align, pose_cur6, pose_pre6, quat_from_rpy=False):
    """distributedMapping.h:1129-1141 / 1249-1256: the pose between two keyframes of a verified loop (tx, ty, tz, qx, qy, qz, qw)."""
    lib = load_library()
    T = np.ascontiguousarray(T_align, np.float32).reshape(16)
    a = np.ascontiguousarray(pose_cur6, np.float32).reshape(6)
    b = np.ascontiguousarray(pose_pre6, np.float32).reshape(6)
    out = SclTransform()
    lib.scl_wire_loop_between(T.ctypes.data, a.ctypes.data, b.ctypes.data, int(bool(quat_from_rpy)), C.byref(out))
    return out.as_tuple()


def wire_pose6_to_transform(pose6):
    lib = load_library()
    a = np.ascontiguousarray(pose6, np.float32).reshape(6)
    out = SclTransform()
    lib.scl_wire_pose6_to_transform(a.ctypes.data, C.byref(out))
    return out.as_tuple()


class SclParams(C.Structure):
    _fields_ = [("num_ring", C.c_int), ("num_sector", C.c_int), ("num_candidates", C.c_int),
                ("dist_thres", C.c_double), ("lidar_height", C.c_double), ("max_radius", C.c_double),
                ("num_exclude_recent", C.c_int), ("tree_making_period", C.c_int), ("search_ratio", C.c_double)]


class SclBatchQuery(C.Structure):
    _fields_ = [("q_desc", C.c_void_p), ("q_ids", C.c_void_p), ("Q", C.c_int), ("K", C.c_int),
                ("n_db", C.c_int), ("metric", C.c_int)]


class SclBatchResult(C.Structure):
    _fields_ = [("cand_ids", C.c_void_p), ("cand_d2", C.c_void_p), ("cand_dist", C.c_void_p),
                ("cand_shift", C.c_void_p), ("best_id", C.c_void_p), ("best_dist", C.c_void_p),
                ("best_shift", C.c_void_p)]


class SclIcpParams(C.Structure):
    _fields_ = [("max_corr_dist", C.c_double), ("max_iterations", C.c_int), ("trans_eps", C.c_double),
                ("fitness_eps", C.c_double)]


class SclIntraResult(C.Structure):
    _fields_ = [("accepted", C.c_int), ("converged", C.c_int), ("iterations", C.c_int), ("fitness", C.c_float), ("T", C.c_float * 16),
                ("n_src", C.c_int), ("n_tgt", C.c_int)]


class SclRansacParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("inlier_threshold", C.c_double), ("min_inlier_ratio", C.c_double), ("seed", C.c_uint)]


_lib = None


def load_library():
    """dlopen the engine. Raises if it has not been built (python -m scl_slam_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise RuntimeError(f"{_SO} is missing: build it with `python scl_slam_b200/build.py` "
                           "(there is no CPU fallback for the Scan Context engine)")
    lib = C.CDLL(_SO)
    lib.scl_last_error.restype = C.c_char_p
    lib.scl_last_error.argtypes = [C.c_void_p]
    lib.scl_create.argtypes = [C.POINTER(SclParams), C.c_int, C.POINTER(C.c_void_p)]
    lib.scl_destroy.argtypes = [C.c_void_p]
    lib.scl_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.scl_reserve.argtypes = [C.c_void_p, C.c_int]
    lib.scl_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.scl_build_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int8, C.c_int, C.c_void_p]
    lib.scl_make_scancontext.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scl_polar_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    lib.scl_build_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scl_build_batch_dev.argtypes = lib.scl_build_batch.argtypes
    lib.scl_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_int8, C.c_int]
    lib.scl_insert_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.scl_insert_batch_dev.argtypes = lib.scl_insert_batch.argtypes
    lib.scl_get_index.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int8), C.POINTER(C.c_int)]
    lib.scl_size.argtypes = [C.c_void_p]
    lib.scl_get_descriptor.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.scl_get_ring_key.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.scl_query_intra.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    lib.scl_query_inter.argtypes = lib.scl_query_intra.argtypes
    lib.scl_query_batch.argtypes = [C.c_void_p, C.POINTER(SclBatchQuery), C.POINTER(SclBatchResult)]
    lib.scl_query_batch_dev.argtypes = lib.scl_query_batch.argtypes
    lib.scl_query_batch_submit.argtypes = [C.c_void_p, C.POINTER(SclBatchQuery), C.POINTER(SclBatchResult), C.POINTER(C.c_int)]
    lib.scl_query_batch_wait.argtypes = [C.c_void_p, C.c_int]
    lib.scl_merge_shards_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.POINTER(SclBatchResult)]
    lib.scl_knn_batch_dev.argtypes = [C.c_void_p, C.POINTER(SclBatchQuery), C.c_void_p, C.c_void_p]
    lib.scl_voxel_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.POINTER(C.c_int)]
    lib.scl_assemble_submap.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.POINTER(C.c_int)]
    lib.scl_build_insert_filtered.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int8, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.scl_xchg_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.scl_xchg_open_local.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.scl_xchg_buffer.argtypes = [C.c_void_p]
    lib.scl_xchg_buffer.restype = C.c_void_p
    lib.scl_xchg_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.scl_query_batch_dev_lane.argtypes = [C.c_void_p, C.c_int, C.POINTER(SclBatchQuery), C.POINTER(SclBatchResult)]
    lib.scl_lanes_fork.argtypes = [C.c_void_p, C.c_void_p]
    lib.scl_lanes_join.argtypes = [C.c_void_p, C.c_void_p]
    lib.scl_lane_sync.argtypes = [C.c_void_p, C.c_int]
    lib.scl_shard_query_dev.argtypes = [C.c_void_p, C.c_int, C.POINTER(SclBatchQuery), C.POINTER(SclBatchResult)]
    lib.scl_shard_query_submit.argtypes = [C.c_void_p, C.POINTER(SclBatchQuery), C.POINTER(SclBatchResult), C.POINTER(C.c_int)]
    lib.scl_xchg_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.scl_xchg_close.argtypes = [C.c_void_p]
    lib.scl_xchg_merge_topk_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scl_xchg_combine_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SclBatchResult)]
    lib.scl_wire_pose6_to_transform.argtypes = [C.c_void_p, C.POINTER(SclTransform)]
    lib.scl_wire_loop_between.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(SclTransform)]
    lib.scl_wire_make_loop_info.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SclLoopInfo)]
    lib.scl_wire_make_global_descriptor.argtypes = [C.c_int, C.c_void_p, C.c_int, C.POINTER(SclTransform), C.c_int, C.c_void_p, C.POINTER(SclGlobalDescriptor)]
    lib.scl_wire_encode_global_descriptor.argtypes = [C.POINTER(SclGlobalDescriptor), C.c_void_p, C.c_int]
    lib.scl_wire_decode_global_descriptor.argtypes = [C.c_void_p, C.c_int, C.POINTER(SclGlobalDescriptor)]
    lib.scl_wire_encode_loop_info.argtypes = [C.POINTER(SclLoopInfo), C.c_void_p, C.c_int]
    lib.scl_wire_decode_loop_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(SclLoopInfo)]
    lib.scl_merge_topk_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.scl_scdist_owned_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scl_combine_owned_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_uint64, C.POINTER(SclBatchResult)]
    lib.scl_verify_ransac.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(SclRansacParams),
                                      C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.scl_create_sharded.argtypes = [C.POINTER(SclParams), C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.scl_sharded_destroy.argtypes = [C.c_void_p]
    lib.scl_sharded_last_error.argtypes = [C.c_void_p]
    lib.scl_sharded_last_error.restype = C.c_char_p
    lib.scl_sharded_world.argtypes = [C.c_void_p]
    lib.scl_sharded_engine.argtypes = [C.c_void_p, C.c_int]
    lib.scl_sharded_engine.restype = C.c_void_p
    lib.scl_sharded_size.argtypes = [C.c_void_p]
    lib.scl_sharded_get_index.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int8), C.POINTER(C.c_int)]
    lib.scl_sharded_get_descriptor.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.scl_sharded_insert_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.scl_sharded_build_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int8, C.c_int, C.c_void_p]
    lib.scl_sharded_query_batch.argtypes = [C.c_void_p, C.POINTER(SclBatchQuery), C.POINTER(SclBatchResult)]
    lib.scl_sharded_query_submit.argtypes = [C.c_void_p, C.POINTER(SclBatchQuery), C.POINTER(SclBatchResult), C.POINTER(C.c_int)]
    lib.scl_sharded_query_wait.argtypes = [C.c_void_p, C.c_int]
    lib.scl_sharded_query_intra.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    lib.scl_sharded_query_inter.argtypes = lib.scl_sharded_query_intra.argtypes
    lib.scl_store_keyframe_cloud.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    lib.scl_keyframe_clouds.argtypes = [C.c_void_p]
    lib.scl_verify_intra.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.POINTER(SclIcpParams), C.c_float,
                                     C.POINTER(SclIntraResult)]
    lib.scl_set_knn_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.scl_set_scdist_mode.argtypes = [C.c_void_p, C.c_int]
    lib.scl_set_tc_stages.argtypes = [C.c_void_p, C.c_int]
    lib.scl_set_scdist_tiles.argtypes = [C.c_void_p, C.c_int]
    lib.scl_export_keys_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.scl_set_replicated_keys_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.scl_knn_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.scl_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.scl_stage_time.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.scl_icp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(SclIcpParams),
                            C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


def _cloud(points):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    if pts.ndim != 2 or pts.shape[1] < 3:
        raise ValueError("points must be [P, >=3] float32 (x, y, z first)")
    return pts, pts.shape[0], pts.shape[1] * 4


def polar_tables(numRing=20, numSector=60, maxRadius=80.0):
    """The bin tables of a geometry (scl_polar_tables; host code, no GPU needed): dict with ring_thr, s_max, and per quadrant
    sec_thr[q], sec_base[q], sec_dir[q]."""
    lib = load_library()
    p = SclParams(numRing, numSector, 3, 0.14, 1.65, maxRadius, 100, 10, 0.1)
    ring = np.zeros(63, np.float32); sec = np.zeros((4, 31), np.float32)
    n_ring = C.c_int32(); s_max = C.c_float()
    n_sec = np.zeros(4, np.int32); base = np.zeros(4, np.int32); sdir = np.zeros(4, np.int32)
    rc = lib.scl_polar_tables(C.byref(p), ring.ctypes.data, C.addressof(n_ring), C.addressof(s_max), sec.ctypes.data,
                              n_sec.ctypes.data, base.ctypes.data, sdir.ctypes.data)
    if rc != SCL_OK:
        raise RuntimeError(f"scl_polar_tables: {_STATUS.get(rc, rc)}")
    return {"ring_thr": ring[:n_ring.value].copy(), "s_max": np.float32(s_max.value),
            "sec_thr": [sec[q, :n_sec[q]].copy() for q in range(4)], "sec_base": base.tolist(), "sec_dir": sdir.tolist()}


def _torch_stream_sync():
    """Block the host until everything queued on torch's current CUDA stream has finished (no-op without torch / CUDA)."""
    try:
        import torch
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
    except ImportError:
        pass


def _ordered_dev_call(fn):
    """Device-pointer calls of an engine that runs on its OWN (non-blocking) stream are not ordered against the caller's stream:
    a tensor torch is still filling may be read too early, and a temporary handed to the call may be freed and reused by torch's
    allocator while the engine still reads it. The Python mirror therefore makes such a call synchronous on both sides; an
    engine bound to the caller's stream (set_stream) is ordered by the stream itself and pays nothing."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        if getattr(self, "_bound", False):
            return fn(self, *a, **kw)
        _torch_stream_sync()
        try:
            return fn(self, *a, **kw)
        finally:
            if getattr(self, "h", None):
                self.lib.scl_lane_sync(self.h, 0)
    return wrapper


class ScanContextB200:
    """Drop-in for scan_context_descriptor (descriptor.h:1304-1801); constructor arguments and
    defaults are those of descriptor.h:1307-1316."""

    def __init__(self, numRing=20, numSector=60, numCandidates=3, distThres=0.14, lidarHeight=1.65,
                 maxRadius=80.0, numExcludeRecent=100, treeMakingPeriod=10, searchRatio=0.1, device=0):
        self.lib = load_library()
        self.params = SclParams(numRing, numSector, numCandidates, distThres, lidarHeight, maxRadius,
                                numExcludeRecent, treeMakingPeriod, searchRatio)
        self.R, self.S, self.K = numRing, numSector, numCandidates
        self.h = C.c_void_p()
        rc = self.lib.scl_create(C.byref(self.params), device, C.byref(self.h))
        if rc != SCL_OK:
            self.h = None
            raise RuntimeError(f"scl_create failed: {_STATUS.get(rc, rc)} (a CUDA device is required; no CPU fallback)")
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.scl_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != SCL_OK:
            raise RuntimeError(f"{_STATUS.get(rc, rc)}: {self.lib.scl_last_error(self.h).decode()}")

    # ---- the six scan_descriptor virtuals (descriptor.h:25-35) ----------------------------
    def makeAndSaveDescriptorAndKey(self, scan, robot, index):
        pts, n, stride = _cloud(scan)
        out = np.empty(self.R * self.S, np.float32)
        self._ck(self.lib.scl_build_insert(self.h, pts.ctypes.data, n, stride, robot, index, out.ctypes.data))
        return out

    def saveDescriptorAndKey(self, descriptorMat, robot, index):
        d = np.ascontiguousarray(descriptorMat, np.float32).reshape(-1)
        if d.size != self.R * self.S:
            raise ValueError("descriptor must have R*S floats")
        self._ck(self.lib.scl_insert(self.h, d.ctypes.data, robot, index))

    def detectIntraLoopClosureID(self, currentPtr):
        i, f = C.c_int(), C.c_float()
        self._ck(self.lib.scl_query_intra(self.h, currentPtr, C.byref(i), C.byref(f)))
        return i.value, f.value

    def detectInterLoopClosureID(self, currentPtr):
        i, f = C.c_int(), C.c_float()
        self._ck(self.lib.scl_query_inter(self.h, currentPtr, C.byref(i), C.byref(f)))
        return i.value, f.value

    def getIndex(self, key):
        r, i = C.c_int8(), C.c_int()
        self._ck(self.lib.scl_get_index(self.h, key, C.byref(r), C.byref(i)))
        return r.value, i.value

    def getSize(self, idIn=-1):
        return self.lib.scl_size(self.h)

    # ---- pieces of the path, for parity checks --------------------------------------------
    def make_scancontext(self, scan, want_bins=False):
        pts, n, stride = _cloud(scan)
        out = np.empty(self.R * self.S, np.float32)
        ring = np.zeros(n, np.int32) if want_bins else None
        sector = np.zeros(n, np.int32) if want_bins else None
        self._ck(self.lib.scl_make_scancontext(self.h, pts.ctypes.data, n, stride, out.ctypes.data, _ptr(ring), _ptr(sector)))
        out = out.reshape(self.R, self.S)
        return (out, ring, sector) if want_bins else out

    def desc(self, key):
        out = np.empty(self.R * self.S, np.float32)
        self._ck(self.lib.scl_get_descriptor(self.h, key, out.ctypes.data))
        return out.reshape(self.R, self.S)

    def ring_key(self, key):
        out = np.empty(self.R, np.float32)
        self._ck(self.lib.scl_get_ring_key(self.h, key, out.ctypes.data))
        return out

    # ---- batched forms ----------------------------------------------------------------------
    def reserve(self, capacity):
        self._ck(self.lib.scl_reserve(self.h, capacity))

    def set_shard(self, rank, world):
        self._ck(self.lib.scl_set_shard(self.h, rank, world))

    def set_stream(self, cuda_stream_handle):
        self._ck(self.lib.scl_set_stream(self.h, C.c_void_p(cuda_stream_handle)))
        self._bound = True                   # the caller's stream now orders the device-pointer calls

    def build_batch(self, clouds, insert=True, robots=None, indices=None):
        """clouds: list of [P_i, C] float32 arrays with one common C. Returns [n, R, S] descriptors."""
        arrs = [np.ascontiguousarray(c, np.float32) for c in clouds]
        stride = arrs[0].shape[1] * 4
        offs = np.zeros(len(arrs) + 1, np.int32)
        offs[1:] = np.cumsum([a.shape[0] for a in arrs])
        pts = np.concatenate(arrs) if arrs else np.zeros((0, 4), np.float32)
        out = np.empty((len(arrs), self.R, self.S), np.float32)
        rb = None if robots is None else np.ascontiguousarray(robots, np.int8)
        ix = None if indices is None else np.ascontiguousarray(indices, np.int32)
        self._ck(self.lib.scl_build_batch(self.h, pts.ctypes.data, offs.ctypes.data, len(arrs), stride, int(insert),
                                          _ptr(rb), _ptr(ix), out.ctypes.data))
        return out

    @_ordered_dev_call
    def build_batch_dev(self, pts_dev, offsets, stride_bytes, insert=True, out_dev=None):
        offs = np.ascontiguousarray(offsets, np.int32)
        self._ck(self.lib.scl_build_batch_dev(self.h, _ptr(pts_dev), offs.ctypes.data, offs.size - 1, stride_bytes,
                                              int(insert), None, None, _ptr(out_dev)))

    def insert_batch(self, descs, robots=None, indices=None):
        d = np.ascontiguousarray(descs, np.float32).reshape(-1, self.R * self.S)
        rb = None if robots is None else np.ascontiguousarray(robots, np.int8)
        ix = None if indices is None else np.ascontiguousarray(indices, np.int32)
        self._ck(self.lib.scl_insert_batch(self.h, d.ctypes.data, d.shape[0], _ptr(rb), _ptr(ix)))

    @_ordered_dev_call
    def insert_batch_dev(self, descs_dev):
        """descs_dev: contiguous float32 CUDA tensor [n, R, S] (or [n, R*S])."""
        n = descs_dev.shape[0]
        self._ck(self.lib.scl_insert_batch_dev(self.h, descs_dev.data_ptr(), n, None, None))

    def query_batch(self, q_desc=None, q_ids=None, K=None, n_db=None, metric=0):
        """Host-buffer batch query. Returns a dict of numpy arrays (see scl_batch_result)."""
        K = K or self.K
        qd = None if q_desc is None else np.ascontiguousarray(q_desc, np.float32).reshape(-1, self.R * self.S)
        qi = None if q_ids is None else np.ascontiguousarray(q_ids, np.int32)
        Q = qd.shape[0] if qd is not None else qi.size
        n_db = self.getSize() if n_db is None else n_db
        out = dict(cand_ids=np.empty((Q, K), np.int32), cand_d2=np.empty((Q, K), np.float32),
                   cand_dist=np.empty((Q, K), np.float64), cand_shift=np.empty((Q, K), np.int32),
                   best_id=np.empty(Q, np.int32), best_dist=np.empty(Q, np.float64), best_shift=np.empty(Q, np.int32))
        q = SclBatchQuery(_ptr(qd), _ptr(qi), Q, K, n_db, metric)
        r = SclBatchResult(*[out[k].ctypes.data for k in
                             ("cand_ids", "cand_d2", "cand_dist", "cand_shift", "best_id", "best_dist", "best_shift")])
        self._ck(self.lib.scl_query_batch(self.h, C.byref(q), C.byref(r)))
        return out

    def query_batch_submit(self, q_desc, out, K=None, n_db=None, metric=0, q_ids=None):
        """Pipelined host-buffer query (scl_query_batch_submit): q_desc (fresh descriptors) or q_ids (keys of stored entries,
        what detectIntra/InterLoopClosureID take) and the arrays of `out` (scl_batch_result field names -> numpy arrays,
        ideally page-locked) must stay alive until query_batch_wait(ticket) returns."""
        K = K or self.K
        Q = q_desc.shape[0] if q_desc is not None else q_ids.shape[0]
        n_db = self.getSize() if n_db is None else n_db
        q = SclBatchQuery(q_desc.ctypes.data if q_desc is not None else None, q_ids.ctypes.data if q_ids is not None else None, Q, K, n_db, metric)
        r = SclBatchResult(*[out[k].ctypes.data if out.get(k) is not None else None for k in
                             ("cand_ids", "cand_d2", "cand_dist", "cand_shift", "best_id", "best_dist", "best_shift")])
        t = C.c_int()
        self._ck(self.lib.scl_query_batch_submit(self.h, C.byref(q), C.byref(r), C.byref(t)))
        return t.value

    def query_batch_wait(self, ticket):
        self._ck(self.lib.scl_query_batch_wait(self.h, ticket))

    @_ordered_dev_call
    def query_batch_dev(self, q_desc_dev, q_ids_dev, Q, K, n_db, metric, out):
        """Device-pointer batch query, asynchronous on the engine's stream. `out` maps the
        scl_batch_result field names to CUDA tensors (missing names are not produced)."""
        q = SclBatchQuery(_ptr(q_desc_dev), _ptr(q_ids_dev), Q, K, n_db, metric)
        r = SclBatchResult(*[_ptr(out.get(k)) for k in
                             ("cand_ids", "cand_d2", "cand_dist", "cand_shift", "best_id", "best_dist", "best_shift")])
        self._ck(self.lib.scl_query_batch_dev(self.h, C.byref(q), C.byref(r)))

    @_ordered_dev_call
    def merge_shards_dev(self, world, Q, K, q_ids_dev, all_ids, all_d2, all_dist, all_shift, out):
        r = SclBatchResult(*[_ptr(out.get(k)) for k in
                             ("cand_ids", "cand_d2", "cand_dist", "cand_shift", "best_id", "best_dist", "best_shift")])
        self._ck(self.lib.scl_merge_shards_dev(self.h, world, Q, K, _ptr(q_ids_dev), _ptr(all_ids), _ptr(all_d2),
                                               _ptr(all_dist), _ptr(all_shift), C.byref(r)))

    def set_knn_mode(self, mode, count_fallbacks=False):
        """0 auto, 1 exact CUDA-core kNN, 2 tensor-core prefilter + exact re-rank (same results)."""
        self._ck(self.lib.scl_set_knn_mode(self.h, mode, int(count_fallbacks)))

    def set_scdist_mode(self, mode):
        """0 FP32 prefilter + exact evaluation of the shifts that can win, 1 every shift exactly (same results)."""
        self._ck(self.lib.scl_set_scdist_mode(self.h, mode))

    def set_tc_stages(self, stages):
        self._ck(self.lib.scl_set_tc_stages(self.h, stages))

    @_ordered_dev_call
    def export_keys_dev(self, keys_out_dev, n):
        """This engine's first n ring keys [n][R] (device to device)."""
        self._ck(self.lib.scl_export_keys_dev(self.h, _ptr(keys_out_dev), n))

    @_ordered_dev_call
    def set_replicated_keys_dev(self, keys_dev, n_total):
        """Hybrid sharding: every ring key, in global key order, on this rank (n_total = 0: back to plain sharding)."""
        self._ck(self.lib.scl_set_replicated_keys_dev(self.h, _ptr(keys_dev), n_total))

    def set_scdist_tiles(self, tiles):
        """Candidate tiles per K4 CTA (0 = one per candidate; 4 = small CTAs that fit beside the tensor-core kNN pass)."""
        self._ck(self.lib.scl_set_scdist_tiles(self.h, tiles))

    def knn_stats(self):
        a, b = C.c_longlong(), C.c_longlong()
        self._ck(self.lib.scl_knn_stats(self.h, C.byref(a), C.byref(b)))
        return {"tc_queries": a.value, "fallback_queries": b.value}

    def set_profiling(self, on):
        self._ck(self.lib.scl_set_profiling(self.h, int(on)))

    def stage_time(self, stage):
        """(total ms, launches) of stage 0=K2 query keys, 1=K3 kNN, 2=K4 SC distance, 3=K1 binning."""
        ms, n = C.c_double(), C.c_int()
        self._ck(self.lib.scl_stage_time(self.h, stage, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- two-phase multi-GPU exchange (include/scl_engine.h) -------------------------------
    @_ordered_dev_call
    def knn_batch_dev(self, q_desc_dev, Q, K, n_db, metric, ids_dev, d2_dev):
        q = SclBatchQuery(_ptr(q_desc_dev), None, Q, K, n_db, metric)
        self._ck(self.lib.scl_knn_batch_dev(self.h, C.byref(q), _ptr(ids_dev), _ptr(d2_dev)))

    @_ordered_dev_call
    def merge_topk_dev(self, world, Q, K, ids_base, d2_base, rank_stride_bytes, out_ids, out_d2):
        self._ck(self.lib.scl_merge_topk_dev(self.h, world, Q, K, _ptr(ids_base), _ptr(d2_base), rank_stride_bytes,
                                             _ptr(out_ids), _ptr(out_d2)))

    @_ordered_dev_call
    def scdist_owned_dev(self, q_desc_dev, Q, K, cand_ids_dev, dist_dev, shift_dev):
        self._ck(self.lib.scl_scdist_owned_dev(self.h, _ptr(q_desc_dev), None, Q, K, _ptr(cand_ids_dev), _ptr(dist_dev), _ptr(shift_dev)))

    @_ordered_dev_call
    def combine_owned_dev(self, world, Q, K, cand_ids, dist_base, shift_base, rank_stride_bytes, out):
        r = SclBatchResult(*[_ptr(out.get(k)) for k in
                             ("cand_ids", "cand_d2", "cand_dist", "cand_shift", "best_id", "best_dist", "best_shift")])
        self._ck(self.lib.scl_combine_owned_dev(self.h, world, Q, K, None, _ptr(cand_ids), _ptr(dist_base), _ptr(shift_base),
                                                rank_stride_bytes, C.byref(r)))

    # ---- peer-memory exchange (csrc/k7_exchange.cu) ----------------------------------------
    def xchg_create(self, world, max_q, max_k):
        h = (C.c_ubyte * 64)()
        self._ck(self.lib.scl_xchg_create(self.h, world, max_q, max_k, h))
        return bytes(h)

    # ---- lanes: several batches in flight on one engine (include/scl_engine.h) -------------------
    def num_lanes(self):
        return self.lib.scl_num_lanes()

    def _result(self, out):
        return SclBatchResult(*[_ptr(out.get(k)) for k in
                                ("cand_ids", "cand_d2", "cand_dist", "cand_shift", "best_id", "best_dist", "best_shift")])

    def query_batch_dev_lane(self, lane, q_desc_dev, q_ids_dev, Q, K, n_db, metric, out):
        q = SclBatchQuery(_ptr(q_desc_dev), _ptr(q_ids_dev), Q, K, n_db, metric)
        r = self._result(out)
        self._ck(self.lib.scl_query_batch_dev_lane(self.h, lane, C.byref(q), C.byref(r)))

    def lanes_fork(self, cuda_stream_handle):
        self._ck(self.lib.scl_lanes_fork(self.h, C.c_void_p(cuda_stream_handle)))

    def lanes_join(self, cuda_stream_handle):
        self._ck(self.lib.scl_lanes_join(self.h, C.c_void_p(cuda_stream_handle)))

    def lane_sync(self, lane):
        self._ck(self.lib.scl_lane_sync(self.h, lane))

    def shard_query_dev(self, lane, q_desc_dev, Q, K, n_db, metric, out):
        """One sharded query step on lane `lane` (every rank calls it with the same arguments)."""
        q = SclBatchQuery(_ptr(q_desc_dev), None, Q, K, n_db, metric)
        r = self._result(out)
        self._ck(self.lib.scl_shard_query_dev(self.h, lane, C.byref(q), C.byref(r)))

    def shard_query_submit(self, q_desc_host, out, K=None, n_db=None, metric=0):
        K = K or self.K
        n_db = self.getSize() if n_db is None else n_db
        q = SclBatchQuery(q_desc_host.ctypes.data, None, q_desc_host.shape[0], K, n_db, metric)
        r = SclBatchResult(*[out[k].ctypes.data if out.get(k) is not None else None for k in
                             ("cand_ids", "cand_d2", "cand_dist", "cand_shift", "best_id", "best_dist", "best_shift")])
        t = C.c_int()
        self._ck(self.lib.scl_shard_query_submit(self.h, C.byref(q), C.byref(r), C.byref(t)))
        return t.value

    def xchg_open(self, world, rank, handles):
        buf = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
        self._ck(self.lib.scl_xchg_open(self.h, world, rank, buf))

    def xchg_close(self):
        self._ck(self.lib.scl_xchg_close(self.h))

    @_ordered_dev_call
    def xchg_merge_topk_dev(self, seq, Q, K, my_block, out_ids, out_d2):
        self._ck(self.lib.scl_xchg_merge_topk_dev(self.h, seq, Q, K, _ptr(my_block), _ptr(out_ids), _ptr(out_d2)))

    @_ordered_dev_call
    def xchg_combine_dev(self, seq, Q, K, my_block, cand_ids, out):
        r = SclBatchResult(*[_ptr(out.get(k)) for k in
                             ("cand_ids", "cand_d2", "cand_dist", "cand_shift", "best_id", "best_dist", "best_shift")])
        self._ck(self.lib.scl_xchg_combine_dev(self.h, seq, Q, K, _ptr(my_block), None, _ptr(cand_ids), C.byref(r)))

    # ---- geometric verification (distributedMapping.h:1108-1132) ---------------------------
    def icp(self, src, tgt, max_corr_dist=100.0, max_iterations=50, trans_eps=1e-6, fitness_eps=1e-6):
        s, ns, stride = _cloud(src)
        t, nt, stride_t = _cloud(tgt)
        if stride != stride_t:
            raise ValueError("src and tgt must share a point stride")
        p = SclIcpParams(max_corr_dist, max_iterations, trans_eps, fitness_eps)
        T = np.empty(16, np.float32)
        fit, conv, it = C.c_float(), C.c_int(), C.c_int()
        self._ck(self.lib.scl_icp(self.h, s.ctypes.data, ns, t.ctypes.data, nt, stride, C.byref(p), T.ctypes.data,
                                  C.byref(fit), C.byref(conv), C.byref(it)))
        return T.reshape(4, 4), fit.value, bool(conv.value), it.value

    def verify_ransac(self, src, tgt, max_iterations=1000, inlier_threshold=0.25, min_inlier_ratio=0.45, seed=1):
        """RANSAC + SVD verification (geometricVerificationService, distributedMapping.h:1211-1243).
        Returns (T 4x4, n_correspondences, n_inliers, success)."""
        s, ns, stride = _cloud(src)
        t, nt, stride_t = _cloud(tgt)
        if stride != stride_t:
            raise ValueError("src and tgt must share a point stride")
        p = SclRansacParams(max_iterations, inlier_threshold, min_inlier_ratio, seed)
        T = np.empty(16, np.float32)
        nc, ni, ok = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.scl_verify_ransac(self.h, s.ctypes.data, ns, t.ctypes.data, nt, stride, C.byref(p), T.ctypes.data,
                                            C.byref(nc), C.byref(ni), C.byref(ok)))
        return T.reshape(4, 4), nc.value, ni.value, bool(ok.value)

    def store_keyframe_cloud(self, key, pts):
        """robots[id].keyFrameArray.push_back(cloud): the cloud stays on the device for verify_intra."""
        p, n, stride = _cloud(pts)
        self._ck(self.lib.scl_store_keyframe_cloud(self.h, key, p.ctypes.data, n, stride))

    def verify_intra(self, key_cur, key_pre, search_num, poses6, leaf, fitness_threshold=0.3, max_corr_dist=100.0, max_iterations=50,
                     trans_eps=1e-6, fitness_eps=1e-6):
        """performIntraLoopClosure after the descriptor stage (distributedMapping.h:1096-1143) as one call. Returns a dict."""
        poses = np.ascontiguousarray(poses6, np.float32).reshape(-1, 6)
        prm = SclIcpParams(max_corr_dist, max_iterations, trans_eps, fitness_eps)
        r = SclIntraResult()
        self._ck(self.lib.scl_verify_intra(self.h, key_cur, key_pre, search_num, poses.ctypes.data, poses.shape[0], leaf, C.byref(prm),
                                           fitness_threshold, C.byref(r)))
        return dict(accepted=bool(r.accepted), converged=bool(r.converged), iterations=r.iterations, fitness=r.fitness,
                    T=np.array(list(r.T), np.float32).reshape(4, 4), n_src=r.n_src, n_tgt=r.n_tgt)

    def voxel_grid(self, pts, leaf):
        """pcl::VoxelGrid<PointXYZI> (distributedMapping.h:996-998, 1181-1185): (m, 4) float32 centroids x, y, z, intensity."""
        p, n, stride = _cloud(pts)
        out = np.empty((max(n, 1), 4), np.float32)
        m = C.c_int()
        self._ck(self.lib.scl_voxel_grid(self.h, p.ctypes.data, n, stride, leaf, out.ctypes.data, C.byref(m)))
        return out[:m.value].copy()

    def assemble_submap(self, clouds, poses6, leaf):
        """loopFindNearKeyframes (distributedMapping.h:1163-1186): clouds moved by their (x, y, z, roll, pitch, yaw) poses,
        concatenated and voxel-filtered (leaf <= 0: not filtered). Returns (m, 4) float32."""
        parts = [_cloud(c) for c in clouds]
        stride = parts[0][2]
        if any(s != stride for _, _, s in parts):
            raise ValueError("clouds must share a point stride")
        pts = np.ascontiguousarray(np.concatenate([p for p, _, _ in parts]))
        offs = np.concatenate([[0], np.cumsum([n for _, n, _ in parts])]).astype(np.int32)
        poses = np.ascontiguousarray(poses6, dtype=np.float32).reshape(len(clouds), 6)
        out = np.empty((max(pts.shape[0], 1), 4), np.float32)
        m = C.c_int()
        self._ck(self.lib.scl_assemble_submap(self.h, pts.ctypes.data, offs.ctypes.data, len(clouds), stride, poses.ctypes.data,
                                              leaf, out.ctypes.data, C.byref(m)))
        return out[:m.value].copy()

    def makeAndSaveDescriptorAndKeyFiltered(self, scan, leaf, robot, index):
        """distributedMapping.h:996-1003 in one call: VoxelGrid(leaf) then makeAndSaveDescriptorAndKey on the device-resident
        filtered cloud. Returns (wire vector of R*S floats, number of filtered points)."""
        p, n, stride = _cloud(scan)
        out = np.empty(self.params.num_ring * self.params.num_sector, np.float32)
        m = C.c_int()
        self._ck(self.lib.scl_build_insert_filtered(self.h, p.ctypes.data, n, stride, leaf, robot, index,
                                                    out.ctypes.data, C.byref(m)))
        return out, m.value


class ShardedScanContextB200:
    """The same descriptor object with its keyframe database sharded by keyframe index over several GPUs of one box, in
    one process (scl_create_sharded, include/scl_engine.h): the six scan_descriptor virtuals + the batched query."""

    def __init__(self, devices, numRing=20, numSector=60, numCandidates=3, distThres=0.14, lidarHeight=1.65, maxRadius=80.0,
                 numExcludeRecent=100, treeMakingPeriod=10, searchRatio=0.1, max_q=1024, max_k=10):
        self.lib = load_library()
        self.params = SclParams(numRing, numSector, numCandidates, distThres, lidarHeight, maxRadius, numExcludeRecent, treeMakingPeriod, searchRatio)
        self.R, self.S, self.K = numRing, numSector, numCandidates
        devs = np.ascontiguousarray(devices, np.int32)
        self.h = C.c_void_p()
        rc = self.lib.scl_create_sharded(C.byref(self.params), devs.size, devs.ctypes.data, max_q, max(max_k, numCandidates), C.byref(self.h))
        if rc != SCL_OK:
            self.h = None
            raise RuntimeError(f"scl_create_sharded failed: {_STATUS.get(rc, rc)} (CUDA devices with peer access are required; no CPU fallback)")

    def close(self):
        if getattr(self, "h", None):
            self.lib.scl_sharded_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != SCL_OK:
            raise RuntimeError(f"{_STATUS.get(rc, rc)}: {self.lib.scl_sharded_last_error(self.h).decode()}")

    def makeAndSaveDescriptorAndKey(self, scan, robot, index):
        pts, n, stride = _cloud(scan)
        out = np.empty(self.R * self.S, np.float32)
        self._ck(self.lib.scl_sharded_build_insert(self.h, pts.ctypes.data, n, stride, robot, index, out.ctypes.data))
        return out

    def saveDescriptorAndKey(self, descriptorMat, robot, index):
        d = np.ascontiguousarray(descriptorMat, np.float32).reshape(-1)
        rb, ix = np.array([robot], np.int8), np.array([index], np.int32)
        self._ck(self.lib.scl_sharded_insert_batch(self.h, d.ctypes.data, 1, rb.ctypes.data, ix.ctypes.data))

    def insert_batch(self, descs, robots=None, indices=None):
        d = np.ascontiguousarray(descs, np.float32).reshape(-1, self.R * self.S)
        rb = None if robots is None else np.ascontiguousarray(robots, np.int8)
        ix = None if indices is None else np.ascontiguousarray(indices, np.int32)
        self._ck(self.lib.scl_sharded_insert_batch(self.h, d.ctypes.data, d.shape[0], _ptr(rb), _ptr(ix)))

    def detectIntraLoopClosureID(self, currentPtr):
        i, f = C.c_int(), C.c_float()
        self._ck(self.lib.scl_sharded_query_intra(self.h, currentPtr, C.byref(i), C.byref(f)))
        return i.value, f.value

    def detectInterLoopClosureID(self, currentPtr):
        i, f = C.c_int(), C.c_float()
        self._ck(self.lib.scl_sharded_query_inter(self.h, currentPtr, C.byref(i), C.byref(f)))
        return i.value, f.value

    def getIndex(self, key):
        r, i = C.c_int8(), C.c_int()
        self._ck(self.lib.scl_sharded_get_index(self.h, key, C.byref(r), C.byref(i)))
        return r.value, i.value

    def getSize(self, idIn=-1):
        return self.lib.scl_sharded_size(self.h)

    def desc(self, key):
        out = np.empty(self.R * self.S, np.float32)
        self._ck(self.lib.scl_sharded_get_descriptor(self.h, key, out.ctypes.data))
        return out.reshape(self.R, self.S)

    def query_batch(self, q_desc, K=None, n_db=None, metric=0, q_ids=None):
        K = K or self.K
        qd = np.ascontiguousarray(q_desc, np.float32).reshape(-1, self.R * self.S)
        qi = None if q_ids is None else np.ascontiguousarray(q_ids, np.int32)
        Q = qd.shape[0]
        n_db = self.getSize() if n_db is None else n_db
        out = dict(cand_ids=np.empty((Q, K), np.int32), cand_d2=np.empty((Q, K), np.float32), cand_dist=np.empty((Q, K), np.float64),
                   cand_shift=np.empty((Q, K), np.int32), best_id=np.empty(Q, np.int32), best_dist=np.empty(Q, np.float64), best_shift=np.empty(Q, np.int32))
        q = SclBatchQuery(qd.ctypes.data, _ptr(qi), Q, K, n_db, metric)
        r = SclBatchResult(*[out[k].ctypes.data for k in ("cand_ids", "cand_d2", "cand_dist", "cand_shift", "best_id", "best_dist", "best_shift")])
        self._ck(self.lib.scl_sharded_query_batch(self.h, C.byref(q), C.byref(r)))
        return out
