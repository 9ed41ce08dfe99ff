/*
 * scl_engine.h — C-ABI of the B200-native Scan Context loop-closure engine.
 *
 * This is the thin boundary the reference's host code binds to. The reference has no C ABI of
 * its own; the interface being replaced is the C++ class boundary
 *   class scan_descriptor            /root/reference/include/descriptor.h:21-36
 *   class scan_context_descriptor    /root/reference/include/descriptor.h:1304-1801
 * as called from distributed_mapping (/root/reference/include/distributedMapping.h:404, 627,
 * 1002, 1072, 1078, 1274, 1280-1284) plus the ICP block of performIntraLoopClosure
 * (distributedMapping.h:1108-1132). include/descriptor_b200.h adapts these entry points back to
 * the six scan_descriptor virtuals; INTEGRATION.md shows the two-line change in
 * distributedMapping.h.
 *
 * Conventions
 *  - every function returns an scl_status code, 0 = OK; nothing throws across the boundary;
 *  - plain pointers and sizes only; pointers are HOST pointers unless the name says _dev;
 *  - descriptors cross the boundary as R*S row-major float32 — the layout of
 *    global_descriptor.msg `values` (descriptor.h:1446-1455 writer, :1575-1582 reader);
 *  - keys are dense global insertion indices 0..N-1 as in the reference (descriptor.h:1593-1599);
 *  - an engine handle is internally serialised (one mutex; inserts on one CUDA stream, query batches on lanes), so the reference's
 *    {insert on the ROS/LIO threads || query on loopClosureThread} pattern
 *    (distributedMapping.h:625-628,1001-1003 vs :1078,1280) is safe;
 *  - there is no CPU fallback: without a CUDA device scl_create fails with SCL_ERR_CUDA.
 */
#ifndef SCL_ENGINE_H_
#define SCL_ENGINE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scl_engine scl_engine;

typedef enum {
    SCL_OK = 0,
    SCL_ERR_INVALID = 1,     /* bad argument */
    SCL_ERR_CUDA = 2,        /* CUDA runtime error (see scl_last_error) */
    SCL_ERR_UNSUPPORTED = 3, /* geometry outside what the kernels are built for */
    SCL_ERR_RANGE = 4,       /* key out of range */
    SCL_ERR_NOMEM = 5
} scl_status;

/* constructor arguments of scan_context_descriptor, descriptor.h:1307-1316 (same defaults) */
typedef struct {
    int num_ring;            /* 20  */
    int num_sector;          /* 60  */
    int num_candidates;      /* 3   */
    double dist_thres;       /* 0.14 */
    double lidar_height;     /* 1.65 */
    double max_radius;       /* 80.0 */
    int num_exclude_recent;  /* 100 */
    int tree_making_period;  /* 10  */
    double search_ratio;     /* 0.1 */
} scl_params;

void scl_default_params(scl_params* p);

/* replaces: new scan_context_descriptor(...) at distributedMapping.h:404 */
int scl_create(const scl_params* p, int device, scl_engine** out);
int scl_destroy(scl_engine* e);
const char* scl_last_error(scl_engine* e);
/* Run on the caller's CUDA stream (a cudaStream_t) instead of the engine's own.
 * Ordering contract of the device-pointer entry points (*_dev): they are ASYNCHRONOUS on the engine's stream, which is a
 * non-blocking stream of its own until this call replaces it. The engine does not order itself against the streams that
 * produce its inputs or consume its outputs: either bind it to that stream here (then plain stream order does it, and a
 * stream-ordered allocator may recycle a buffer right after the call), or synchronise around the call (inputs complete
 * before it, scl_lane_sync(e, 0) before the outputs are read or the inputs freed). Host-pointer entry points are synchronous. */
int scl_set_stream(scl_engine* e, void* cuda_stream);
/* pre-size the database arrays in HBM (they otherwise grow by doubling) */
int scl_reserve(scl_engine* e, int capacity);
/* Database sharding across ranks: local key l stands for global key
 * l*world + rank in every id this engine reports. Default rank 0, world 1. */
int scl_set_shard(scl_engine* e, int rank, int world);
/* Hybrid sharding: descriptors stay sharded by key mod world (what K4 gathers: 4.8 KB per candidate), the ring
 * keys (80 B per keyframe) are REPLICATED on every rank, in global key order. K3 then runs query-parallel: a rank searches every
 * key for its 1 / world of a batch and the other queries' lists stay empty, so the exchange's per-query merge returns the
 * owner's list; no rank repeats K3 start-up, the re-rank or the fallback for all Q queries. With replicated keys set, n_db of
 * scl_knn_batch_dev / scl_shard_query_dev / scl_shard_query_submit is the GLOBAL search bound. n_total = 0 returns to
 * plain sharding. scl_export_keys_dev copies this engine's first n ring keys ([n][num_ring] floats, device to device), which
 * the ranks all-gather and interleave (global key = local * world + rank) before handing them back. */
int scl_export_keys_dev(scl_engine* e, float* keys_out_dev, int n);
int scl_set_replicated_keys_dev(scl_engine* e, const float* keys_dev, int n_total);

/* ---- descriptor build --------------------------------------------------------------------
 * replaces makeAndSaveDescriptorAndKey, descriptor.h:1604-1611 (call site distributedMapping.h:1002).
 * pts: n points, stride_bytes apart (32 for pcl::PointXYZI), float x,y,z at byte 0,4,8.
 * out_desc (R*S floats, may be NULL) is the return vector of the reference function. */
int scl_build_insert(scl_engine* e, const void* pts, int n, int stride_bytes, int8_t robot, int index,
                     float* out_desc);
/* makeScancontext alone (descriptor.h:1404-1461): no insert. out_ring/out_sector (n ints each,
 * may be NULL) receive the 1-based bin of every point, 0 for dropped points (parity checks). */
int scl_make_scancontext(scl_engine* e, const void* pts, int n, int stride_bytes, float* out_desc,
                         int32_t* out_ring, int32_t* out_sector);
/* The bin tables of a geometry (host code only, no device needed; for inspection and for the CPU parity test). The kernel does
 * not evaluate atanf, sqrt or the double-precision index formulas of descriptor.h:1352-1374,1425-1435 per point: ring and
 * sector are monotone step functions of one float each, and these are the exact floats at which they step.
 *   ring   = 1 + #{ i < *n_ring : ring_thr[i] <= s },  s = fl(fl(x*x) + fl(y*y))   (the float the reference hands to sqrt);
 *            a point is dropped when s > *s_max  (double(sqrtf(s)) > max_radius, :1429)
 *   sector = sec_base[q] + sec_dir[q] * #{ k < n_sec[q] : sec_thr[31*q + k] <= t },  q = quadrant (0: x>=0,y>=0; 1: x<0,y>=0;
 *            2: x<0,y<0; 3: x>=0,y<0),  t = fl(|y| / |x|) as xy2theta hands it to atan; t NaN -> sector 1
 * ring_thr: room for 63 floats, sec_thr: 4 x 31 floats, n_sec / sec_base / sec_dir: 4 ints each. Any pointer may be NULL. */
int scl_polar_tables(const scl_params* p, float* ring_thr, int32_t* n_ring, float* s_max, float* sec_thr, int32_t* n_sec,
                     int32_t* sec_base, int32_t* sec_dir);
/* Batched build: n_scans clouds concatenated in pts; scan i is points [offsets[i], offsets[i+1]).
 * insert != 0 appends them (robots/indices give the metadata, may be NULL -> robot 0, index = key).
 * out_desc: n_scans*R*S floats or NULL. */
int scl_build_batch(scl_engine* e, const void* pts, const int32_t* offsets, int n_scans, int stride_bytes,
                    int insert, const int8_t* robots, const int32_t* indices, float* out_desc);
/* same with pts/out_desc in device memory (offsets stay on the host) */
int scl_build_batch_dev(scl_engine* e, const void* pts_dev, const int32_t* offsets, int n_scans, int stride_bytes,
                        int insert, const int8_t* robots, const int32_t* indices, float* out_desc_dev);

/* ---- database insert ---------------------------------------------------------------------
 * replaces saveDescriptorAndKey, descriptor.h:1572-1585 (call site distributedMapping.h:627) */
int scl_insert(scl_engine* e, const float* desc, int8_t robot, int index);
int scl_insert_batch(scl_engine* e, const float* descs, int n, const int8_t* robots, const int32_t* indices);
int scl_insert_batch_dev(scl_engine* e, const float* descs_dev, int n, const int8_t* robots, const int32_t* indices);

/* ---- accessors ---------------------------------------------------------------------------
 * getIndex / getSize, descriptor.h:1758-1766. Out-of-range keys (the reference's getIndex(-1) at
 * distributedMapping.h:1282) give robot = -1, index = -1 and SCL_OK. */
int scl_get_index(scl_engine* e, int key, int8_t* robot, int* index);
int scl_size(scl_engine* e);
int scl_get_descriptor(scl_engine* e, int key, float* out_desc); /* R*S floats */
int scl_get_ring_key(scl_engine* e, int key, float* out_key);    /* R floats, descriptor.h:1463-1475 */

/* ---- loop queries ------------------------------------------------------------------------
 * replaces detectIntraLoopClosureID (descriptor.h:1613-1674, call site distributedMapping.h:1078):
 *   id = -1 if no loop; second = best shift as a float.
 * and detectInterLoopClosureID (descriptor.h:1676-1756, call site distributedMapping.h:1280):
 *   id = -1 if no loop; second = yaw difference in radians; the searched key range is refreshed
 *   every tree_making_period calls exactly like the reference's KD-tree rebuild. */
int scl_query_intra(scl_engine* e, int cur, int* id, float* second);
int scl_query_inter(scl_engine* e, int cur, int* id, float* second);

/* Batched throughput form (ring-key top-K + shift-aligned SC distance for every candidate).
 *   q_desc : Q*R*S query descriptors, or NULL to take the queries from the database by q_ids
 *   q_ids  : Q keys; with q_desc they only drive the self-skip rule (descriptor.h:1731), may be NULL
 *   n_db   : keys [0, n_db) are searched
 *   metric : 0 = nanoflann accumulation order (nanoflann.hpp:383-408)
 *            1 = sequential order + libnabo's self-match rule (d2 <= FLT_EPSILON skipped)
 * Results, any of which may be NULL: cand_* are Q*K in kNN order (id -1 / d2 FLT_MAX / dist NaN for
 * missing); best_* apply the strict-< scan of descriptor.h:1721-1737 (best_id -1, dist 1e7 if none). */
typedef struct {
    const float* q_desc;
    const int32_t* q_ids;
    int Q, K, n_db, metric;
} scl_batch_query;
typedef struct {
    int32_t* cand_ids;
    float* cand_d2;
    double* cand_dist;
    int32_t* cand_shift;
    int32_t* best_id;
    double* best_dist;
    int32_t* best_shift;
} scl_batch_result;
int scl_query_batch(scl_engine* e, const scl_batch_query* q, scl_batch_result* r);
/* all pointers in q and r are device pointers; asynchronous on the engine's stream */
int scl_query_batch_dev(scl_engine* e, const scl_batch_query* q, scl_batch_result* r);
/* Pipelined host-buffer form, for a caller that streams batches (loopClosureThread draining a backlog,
 * distributedMapping.h:1450-1473): submit returns as soon as the batch is enqueued, so the next batch can be
* submitted at once and its host-to-device copy (every batch runs on its own query lane: stream + staging + scratch) overlaps
 * the kernels of this one. scl_query_batch_wait(ticket) blocks until that batch's results are in the host arrays of r.
 * At most scl_num_lanes() batches may be in flight (tickets are consecutive; wait for them in order); q_desc and the result arrays should be page-locked host memory (pageable
 * memory works but serialises the copies) and must stay valid until the wait returns. */
int scl_query_batch_submit(scl_engine* e, const scl_batch_query* q, scl_batch_result* r, int* ticket);
int scl_query_batch_wait(scl_engine* e, int ticket);

/* Query lanes. An engine keeps scl_num_lanes() independent sets of query scratch, each with its own CUDA stream, so that
 * several batches are in flight at once: the kernels of one batch leave SMs idle (start-up, the exact re-rank, exchange
 * waits) that the next batch fills. scl_query_batch_submit rotates over the lanes by itself; for device-resident queries
 * scl_query_batch_dev_lane enqueues a batch on lane `lane` (lane 0 is the engine's own stream, what scl_query_batch_dev
 * uses). scl_lanes_fork makes every lane wait for the work enqueued so far on `cuda_stream` (the stream the caller filled
 * the query buffers on), scl_lanes_join makes `cuda_stream` wait for everything enqueued so far on every lane, so a caller
 * can bracket a group of batches with events on its own stream; scl_lane_sync blocks the host until lane `lane` is idle.
 * Results of a lane are valid once its stream has passed them; a lane's scratch is reused by its next batch. */
int scl_num_lanes(void);
int scl_query_batch_dev_lane(scl_engine* e, int lane, const scl_batch_query* q, scl_batch_result* r);
int scl_lanes_fork(scl_engine* e, void* cuda_stream);
int scl_lanes_join(scl_engine* e, void* cuda_stream);
int scl_lane_sync(scl_engine* e, int lane);

/* Multi-GPU merge: `world` per-rank result blocks of Q*K records, gathered
 * rank-major in device memory, are reduced to the global top-K by (d2, id) and then to the winner
 * by the same strict-< scan as the unsharded query. */
int scl_merge_shards_dev(scl_engine* e, int world, int Q, int K, const int32_t* q_ids,
                         const int32_t* all_ids, const float* all_d2, const double* all_dist, const int32_t* all_shift,
                         scl_batch_result* merged);

/* Two-phase exchange (what bench.py uses for N > 1): the SC distance is computed once per GLOBAL
 * candidate, by the rank that owns it, instead of K times per rank.
 *   1. scl_knn_batch_dev      : K2 + K3 on the shard -> local (id, d2) lists (device pointers, asynchronous)
 *      all-gather of the (id, d2) blocks, rank blocks rank_stride_bytes apart
 *   2. scl_merge_topk_dev     : global top-K by (d2, id), identical on every rank
 *   3. scl_scdist_owned_dev   : K4 for the candidates with id mod world == rank (others: NaN / 0)
 *      all-gather of the (dist, shift) blocks
 *   4. scl_combine_owned_dev  : take each candidate's result from its owner's block + the winner scan */
int scl_knn_batch_dev(scl_engine* e, const scl_batch_query* q, int32_t* ids_dev, float* d2_dev);
int scl_merge_topk_dev(scl_engine* e, int world, int Q, int K, const void* ids_base, const void* d2_base, uint64_t rank_stride_bytes,
                       int32_t* out_ids, float* out_d2);
int scl_scdist_owned_dev(scl_engine* e, const float* q_desc_dev, const int32_t* q_ids_dev, int Q, int K, const int32_t* cand_ids_dev,
                         double* dist_dev, int32_t* shift_dev);
int scl_combine_owned_dev(scl_engine* e, int world, int Q, int K, const int32_t* q_ids, const int32_t* cand_ids, const void* dist_base,
                          const void* shift_base, uint64_t rank_stride_bytes, scl_batch_result* merged);

/* The same two exchange points over NVLink peer memory instead of NCCL (csrc/k7_exchange.cu): every rank creates an exchange
 * buffer (room for batches of max_q queries with up to max_k candidates, one region per query lane) and hands its 64-byte
 * IPC handle to the others (any side channel: torch.distributed, MPI, a file); after scl_xchg_open,
 * scl_xchg_merge_topk_dev replaces {all-gather, scl_merge_topk_dev} and scl_xchg_combine_dev replaces
 * {all-gather, scl_combine_owned_dev}: one kernel each stores this rank's block into every peer's buffer, raises a flag,
 * waits for all ranks' flags and merges. seq = 1, 2, 3, ... must advance by one per query step, identically on all ranks;
 * my_block_dev = [ids i32 Q*K | d2 f32 Q*K] resp. [dist f64 Q*K | shift i32 Q*K], 16-byte aligned; Q*K a multiple of 4.
 * (These two run on lane 0 with the caller's step counter; scl_shard_query_dev below keeps the counters itself.)
 * scl_xchg_open_local is scl_xchg_open for engines of ONE process (one per device, peer access enabled): peer_bufs[r] =
 * scl_xchg_buffer(engine of rank r). */
int scl_xchg_create(scl_engine* e, int world, int max_q, int max_k, unsigned char* handle64);
int scl_xchg_open(scl_engine* e, int world, int rank, const unsigned char* handles /* world x 64 bytes, rank-major */);
int scl_xchg_open_local(scl_engine* e, int world, int rank, void* const* peer_bufs);
void* scl_xchg_buffer(scl_engine* e);
int scl_xchg_bytes(scl_engine* e, int world, int max_q, int max_k, uint64_t* bytes);
int scl_xchg_merge_topk_dev(scl_engine* e, int seq, int Q, int K, const void* my_block_dev, int32_t* out_ids, float* out_d2);
int scl_xchg_combine_dev(scl_engine* e, int seq, int Q, int K, const void* my_block_dev, const int32_t* q_ids, const int32_t* cand_ids,
                         scl_batch_result* merged);
int scl_xchg_close(scl_engine* e);

/* The sharded query as ONE call per batch (every rank makes the same calls in the same order, with the same lane):
 *   scl_shard_query_dev   : q->q_desc = the whole batch in device memory (every rank holds it); K2 + K3 on the shard,
 *                           exchange + global top-K, K4 on the candidates this rank owns, exchange + winner scan, all on
 *                           lane `lane`; r holds device pointers (any may be NULL), identical on all ranks afterwards.
 *   scl_shard_query_submit: host buffers, pipelined like scl_query_batch_submit (wait with scl_query_batch_wait). Every
 *                           rank passes the same whole batch in q->q_desc but uploads only its 1/world of the rows; a
 *                           gather kernel stores them into every peer's query area over NVLink, so the host link carries each
 *                           descriptor once. q->q_ids (optional, host) are the queries' own keys for the self-skip rule. */
int scl_shard_query_dev(scl_engine* e, int lane, const scl_batch_query* q, scl_batch_result* r);
int scl_shard_query_submit(scl_engine* e, const scl_batch_query* q, scl_batch_result* r, int* ticket);

/* ---- ring-key kNN variant ---------------------------------------------
 * mode 0 = automatic (tensor-core prefilter for batches of more than 3 queries on >= 32768 keys, exact
 * CUDA-core kernel otherwise), 1 = always exact, 2 = always tensor core. Both produce identical
 * results: the prefilter's proposals are re-ranked exactly and certified, uncertified queries are
 * redone by the exact kernel. With count_fallbacks != 0 every batch reads back how many queries
 * needed that (a host sync; for tests and reports). */
int scl_set_knn_mode(scl_engine* e, int mode, int count_fallbacks);
/* ---- SC distance variant ------------------------------------------------
 * mode 0 (default): FP32 estimates of every alignment / window shift, then the shifts that can win (within a proven error
 * bound of the best estimate) in FP64 with the operation order of distanceBtnScanContext (descriptor.h:1538-1569);
 * mode 1: every shift in FP64. Both give bit-identical distances and shifts; mode 1 exists for the tests. */
int scl_set_scdist_mode(scl_engine* e, int mode);
/* Candidate tiles (= scoring warps) of one K4 CTA on an unsharded engine: 0 (default) = one per candidate, up to 10 at 20x60;
 * 2..16 = that many, every warp then scores several candidates in turn and the CTA keeps no double copy of the query.
 * Small CTAs (4 tiles: 128 threads, 48 KB) fit beside a knn_tc_kernel CTA of another query lane on the same SM - its
 * register file is divided by scheduler, 16 384 registers each, of which knn_tc holds 9 216 - so that the SC distances of
 * one batch run under the tensor-core pass of the next. Results do not depend on it. */
int scl_set_scdist_tiles(scl_engine* e, int tiles);
int scl_knn_stats(scl_engine* e, long long* tc_queries, long long* fallback_queries);
/* Shared memory of the tensor-core kNN kernel: `stages` key tiles (32 KB each at 20 rings) are in flight per SM, 2..5
 * (default 2: measured no slower than 5, the kernel is not TMA-bound). Fewer stages leave shared memory to the kernels of other
 * query lanes (re-rank, SC distance), which then run on the same SMs at the same time; results do not depend on it. */
int scl_set_tc_stages(scl_engine* e, int stages);

/* ---- per-stage device timing (for roofline reports) ----------------------------------------
 * With profiling on, every batched query records CUDA events on the engine's stream around each
 * stage. scl_stage_time synchronises, sums the elapsed time of the recorded launches of `stage`
 * (0 = query ring keys K2, 1 = ring-key kNN K3, 2 = SC distance K4, 3 = polar binning K1) into
 * *ms and their number into *launches, and clears the record. */
int scl_set_profiling(scl_engine* e, int on);
int scl_stage_time(scl_engine* e, int stage, double* ms, int* launches);

/* ---- geometric verification --------------------------------------------------------------
 * replaces the pcl::IterativeClosestPoint block of performIntraLoopClosure,
 * distributedMapping.h:1108-1132: point-to-point ICP of src onto tgt.
 * T_out: row-major 4x4 (getFinalTransformation), fitness = getFitnessScore(), converged = hasConverged(). */
typedef struct {
    double max_corr_dist;   /* 100.0, distributedMapping.h:1109 */
    int max_iterations;     /* 50,    :1110 */
    double trans_eps;       /* 1e-6,  :1111 */
    double fitness_eps;     /* 1e-6,  :1112 */
} scl_icp_params;
void scl_default_icp_params(scl_icp_params* p);
int scl_icp(scl_engine* e, const void* src, int n_src, const void* tgt, int n_tgt, int stride_bytes,
            const scl_icp_params* p, float* T_out, float* fitness, int* converged, int* iterations);

/* replaces the RANSAC + SVD block of geometricVerificationService, distributedMapping.h:1211-1243 (the inter-robot
 * verification): nearest-neighbour correspondences src -> tgt, max_iterations three-point hypotheses scored by the
 * inlier threshold, closed-form fit on the inliers of the best one, success iff inliers >= min_inlier_ratio * corr.
 * PCL's sampler is random (and stops early); here every hypothesis is evaluated and the seed makes the result
 * reproducible. T_out: row-major 4x4. */
typedef struct {
    int max_iterations;       /* ransacMaxIter 1000, distributedMapping.h:187 */
    double inlier_threshold;  /* ransacOutlierTreshold 0.25 m, :188 */
    double min_inlier_ratio;  /* inlierTreshold 0.45, :189 */
    unsigned seed;
} scl_ransac_params;
void scl_default_ransac_params(scl_ransac_params* p);
int scl_verify_ransac(scl_engine* e, const void* src, int n_src, const void* tgt, int n_tgt, int stride_bytes,
                      const scl_ransac_params* p, float* T_out, int* n_corr, int* n_inliers, int* success);

/* ---- cloud preparation --------------------------------------------------
 * Points are 16-byte aligned records stride_bytes apart with x, y, z in the first 12 bytes: stride 16 = packed
 * (x, y, z, intensity); stride >= 32 = pcl::PointXYZI (intensity at byte 16).
 * Outputs are packed: 4 floats (x, y, z, intensity) per point; the caller provides room for every input point.
 *
 * scl_voxel_grid replaces pcl::VoxelGrid<PointXYZI>::filter with setLeafSize(leaf, leaf, leaf) — downSizeFilterDes in
 * front of the descriptor (distributedMapping.h:996-998) and downSizeFilterICP (:1181-1185, :1200-1201): one centroid
 * per occupied leaf, ordered by the linear leaf index; non-finite points are skipped. */
int scl_voxel_grid(scl_engine* e, const void* pts, int n, int stride_bytes, float leaf, float* out_xyzi, int* n_out);
/* scl_build_insert_filtered = the producer step of distributedMapping.h:996-1003 in one call: downSizeFilterDes.filter(cloud)
 * followed by makeAndSaveDescriptorAndKey(filtered, robot, index); the filtered cloud stays on the device. out_desc (R*S
 * floats) and n_filtered may be NULL. */
int scl_build_insert_filtered(scl_engine* e, const void* pts, int n, int stride_bytes, float leaf, int8_t robot, int index,
                              float* out_desc, int* n_filtered);
/* scl_assemble_submap replaces loopFindNearKeyframes (distributedMapping.h:1163-1186): cloud c (points
 * offsets[c]..offsets[c+1]) is moved by pose c = (x, y, z, roll, pitch, yaw) as transformPointCloud does (:234-253),
 * the clouds are concatenated in order and down-sampled with scl_voxel_grid(leaf); leaf <= 0 skips the down-sampling. */
int scl_assemble_submap(scl_engine* e, const void* pts, const int* offsets, int n_clouds, int stride_bytes, const float* poses6,
                        float leaf, float* out_xyzi, int* n_out);

/* ---- one database over the GPUs of a box, in one process (csrc/sharded.cu) ------------------------------------------
 * distributed_mapping constructs ONE descriptor object (distributedMapping.h:333,402-405). scl_create_sharded is that
 * object with the keyframe database sharded by keyframe index over `ndev` devices (key mod ndev; one engine per device,
 * their exchange buffers mapped into each other by peer access). Keys, ids and results mean exactly what they mean on
 * a single engine. max_q / max_k size the exchange buffers (largest batch and candidate count of a query). Queries take
 * HOST buffers; up to scl_num_lanes() batches may be in flight (submit / wait, tickets in order).
 * With ndev = 1 it is a plain engine behind the same calls. */
typedef struct scl_sharded scl_sharded;
int scl_create_sharded(const scl_params* p, int ndev, const int* devs, int max_q, int max_k, scl_sharded** out);
int scl_sharded_destroy(scl_sharded* s);
const char* scl_sharded_last_error(scl_sharded* s);
int scl_sharded_world(scl_sharded* s);
scl_engine* scl_sharded_engine(scl_sharded* s, int rank);       /* the shard's engine (descriptor build, ICP, ... are per device) */
int scl_sharded_size(scl_sharded* s);                            /* getSize, descriptor.h:1763-1766 */
int scl_sharded_get_index(scl_sharded* s, int key, int8_t* robot, int* index);   /* getIndex, :1758-1761 */
int scl_sharded_get_descriptor(scl_sharded* s, int key, float* out_desc);
int scl_sharded_insert_batch(scl_sharded* s, const float* descs, int n, const int8_t* robots, const int32_t* indices);   /* saveDescriptorAndKey */
int scl_sharded_build_insert(scl_sharded* s, const void* pts, int n, int stride_bytes, int8_t robot, int index, float* out_desc);   /* makeAndSaveDescriptorAndKey */
int scl_sharded_query_batch(scl_sharded* s, const scl_batch_query* q, scl_batch_result* r);
int scl_sharded_query_submit(scl_sharded* s, const scl_batch_query* q, scl_batch_result* r, int* ticket);
int scl_sharded_query_wait(scl_sharded* s, int ticket);
int scl_sharded_query_intra(scl_sharded* s, int cur, int* id, float* second);    /* detectIntraLoopClosureID, :1613-1674 */
int scl_sharded_query_inter(scl_sharded* s, int cur, int* id, float* second);    /* detectInterLoopClosureID, :1676-1756 */

/* ---- the intra-robot verification as one call (distributedMapping.h:1096-1143) ---------------------------------------
 * scl_store_keyframe_cloud keeps keyframe `key`'s cloud on the device (robots[id].keyFrameArray of the reference; keys are
 * stored in order, key = number stored so far). scl_verify_intra then does what performIntraLoopClosure does after the
 * descriptor stage: loopFindNearKeyframes(cur, 0) and loopFindNearKeyframes(pre, search_num) (:1163-1186: every cloud moved
 * by its pose, merged, voxel-filtered with `leaf`), the size gates (< 300 / < 1000 points: no loop, :1102), the ICP of
 * :1108-1121 and the fitness gate of :1122 — without a cloud leaving the device. poses6 = x, y, z, roll, pitch, yaw of
 * every keyframe (cloudKeyPoses6D). out->T is icp.getFinalTransformation() (row-major 4x4); the pose between the two
 * keyframes for loop_info follows with scl_wire_loop_between(out->T, pose_cur, pose_pre, ...) (include/scl_wire.h). */
typedef struct {
    int accepted;      /* converged and fitness <= threshold */
    int converged;
    int iterations;
    float fitness;     /* icp.getFitnessScore(), the loop's noise (:1151) */
    float T[16];
    int n_src, n_tgt;  /* points of the two down-sampled clouds */
} scl_intra_result;
int scl_store_keyframe_cloud(scl_engine* e, int key, const void* pts, int n, int stride_bytes);
int scl_keyframe_clouds(scl_engine* e);
int scl_verify_intra(scl_engine* e, int key_cur, int key_pre, int search_num, const float* poses6, int n_poses, float leaf,
                     const scl_icp_params* icp, float fitness_threshold, scl_intra_result* out);

#ifdef __cplusplus
}
#endif
#endif
