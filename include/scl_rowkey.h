/*
 * scl_rowkey.h — C-ABI of the row-key candidate search (second half): the kNN stage of the second
 * descriptor family of the reference, class lidar_iris_descriptor (/root/reference/include/descriptor.h:462-1302), on the
 * same K3 kernels as the Scan Context ring keys, with R = 80 rows by default.
 *
 * What is replaced: the three places where that class builds a libnabo KD-tree over float row keys and asks it for the
 * `numCandidates` nearest ones,
 *   save()                        descriptor.h:1047-1059   per-robot key matrices, local2Global, the global index pairs
 *   detectIntraLoopClosureID      descriptor.h:1087-1148   own robot's keys [0, curPtr - numExcludeRecent), tree rebuilt per call
 *   detectInterLoopClosureID      descriptor.h:1150-1250   all other robots' keys concatenated (query from this robot) or this
 *                                                           robot's keys (query from another robot), tree rebuilt per call
 *   getIndex / getSize            descriptor.h:1252-1267
 * What stays with the caller: the LiDAR-Iris image, its log-Gabor features and compare() (descriptor.h:515-1023, OpenCV —
 * out of scope). The detect_* entry points take compare() as a callback, so that the candidate loop, the
 * strict-< minimum and the threshold test run here exactly as in the reference; the *_candidates entry points return the
 * kNN lists alone.
 *
 * Conventions as in scl_engine.h: int status codes (scl_status), HOST pointers, nothing throws, one mutex per handle,
 * no CPU fallback (scl_rowkey_create fails without a CUDA device). kNN flavour: libnabo's (sequential float accumulation,
 * ascending results, d2 <= FLT_EPSILON skipped because the reference passes no ALLOW_SELF_MATCH, missing neighbours
 * reported as index -1 / distance +inf); exact ties: lowest index first (libnabo's own tie order is unpinned).
 */
#ifndef SCL_ROWKEY_H_
#define SCL_ROWKEY_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scl_rowkey scl_rowkey;

/* constructor arguments of lidar_iris_descriptor that reach the candidate stage, descriptor.h:472-499 (same defaults) */
typedef struct {
    int rows;                /* 80: length of a row key */
    int num_exclude_recent;  /* 30 */
    int num_candidates;      /* 10 */
    double dist_thres;       /* 0.32 */
    int robot_num;           /* 1 */
    int this_id;             /* 0 */
} scl_rowkey_params;

void scl_rowkey_default_params(scl_rowkey_params* p);

/* replaces: new lidar_iris_descriptor(...) at distributedMapping.h:408 (the candidate stage of it) */
int scl_rowkey_create(const scl_rowkey_params* p, int device, scl_rowkey** out);
int scl_rowkey_destroy(scl_rowkey* h);
const char* scl_rowkey_last_error(scl_rowkey* h);

/* save(), descriptor.h:1047-1059: appends the row key to robot's matrix, local2Global[robot] gets the new global key,
 * the global index gets (robot, index). global_key (may be NULL) receives that global key. */
int scl_rowkey_save(scl_rowkey* h, const float* row_key, int8_t robot, int index, int* global_key);
/* n keys of one robot at once ([n][rows] floats); index[i] as in scl_rowkey_save, or NULL for local positions (size before + i) */
int scl_rowkey_save_batch(scl_rowkey* h, const float* row_keys, int n, int8_t robot, const int* index);
/* saveDescriptorAndKey, descriptor.h:1025-1045: the wire vector of makeAndSaveDescriptorAndKey (rows*cols image values, then the
 * rows floats of the key, :1075-1081); only the key is kept here */
int scl_rowkey_save_wire(scl_rowkey* h, const float* wire, int cols, int8_t robot, int index, int* global_key);

/* The kNN of detectIntraLoopClosureID (descriptor.h:1087-1114): query = this robot's key cur_ptr (a LOCAL position), over
 * this robot's keys [0, cur_ptr - num_exclude_recent). *n = 0 when the reference returns early (:1094-1097), else
 * num_candidates; local_idx / d2 (num_candidates entries each) in kNN order, -1 / +inf where libnabo found no neighbour. */
int scl_rowkey_intra_candidates(scl_rowkey* h, int cur_ptr, int* n, int32_t* local_idx, float* d2);
/* The kNN of detectInterLoopClosureID (descriptor.h:1150-1209): query = global key cur_ptr. concat_idx: position in the
 * concatenated key matrix the reference builds (:1164-1191); global_key = newLocal2Global[concat_idx] (-1 for none).
 * *n = 0 when fewer than num_candidates + 1 keys are available (:1194-1197). */
int scl_rowkey_inter_candidates(scl_rowkey* h, int cur_ptr, int* n, int32_t* concat_idx, int32_t* global_key, float* d2);

/* compare() of the caller (descriptor.h:964-1023) on two stored entries, each named by (robot, local position) */
typedef float (*scl_rowkey_compare_fn)(void* user, int8_t robot_a, int local_a, int8_t robot_b, int local_b, int* bias);

/* detectIntraLoopClosureID / detectInterLoopClosureID as a whole: kNN here, compare() through the callback, the candidate
 * scan with strict <, the skip of out-of-range indices and the threshold test as in descriptor.h:1116-1147 / :1211-1249.
 * id = -1, bias = 0 when there is no loop (the reference's {-1, 0.0}); the intra id is a LOCAL position of this robot, the
 * inter id a GLOBAL key (as in the reference). min_dist (may be NULL): the smallest compare() value seen, 1e7 if none. */
int scl_rowkey_detect_intra(scl_rowkey* h, int cur_ptr, scl_rowkey_compare_fn cmp, void* user, int* id, float* bias, float* min_dist);
int scl_rowkey_detect_inter(scl_rowkey* h, int cur_ptr, scl_rowkey_compare_fn cmp, void* user, int* id, float* bias, float* min_dist);

/* getIndex (descriptor.h:1252-1255; out of range: robot = -1, index = -1) and getSize (:1257-1267; id_in = -1: all) */
int scl_rowkey_get_index(scl_rowkey* h, int key, int8_t* robot, int* index);
int scl_rowkey_size(scl_rowkey* h, int id_in);

/* Batched form (what the GPU is for): Q queries at once against the key set an inter query of `from_robot` would see
 * (from_robot == this_id: all other robots concatenated; else this robot's keys), or, with from_robot = -1, against
 * robot this_id's first n_limit keys (the intra key set of cur_ptr = n_limit + num_exclude_recent). q_keys [Q][rows] host
 * floats; results [Q][K]. knn_mode as scl_set_knn_mode: 0 automatic, 1 exact CUDA-core kernel, 2 tensor-core prefilter
 * (rows = 20, 40 or 80) + exact re-rank + certificate; all give the same lists. */
int scl_rowkey_knn_batch(scl_rowkey* h, const float* q_keys, int Q, int from_robot, int n_limit, int K, int knn_mode,
                         int32_t* concat_idx, int32_t* global_key, float* d2);
/* the same on device pointers, enqueued on the handle's stream (a cudaStream_t through scl_rowkey_set_stream); no host sync */
int scl_rowkey_knn_batch_dev(scl_rowkey* h, const float* q_keys_dev, int Q, int from_robot, int n_limit, int K, int knn_mode,
                             int32_t* concat_idx_dev, float* d2_dev);
int scl_rowkey_set_stream(scl_rowkey* h, void* cuda_stream);
/* queries the tensor-core path took / had to hand to the exact kernel since creation */
int scl_rowkey_knn_stats(scl_rowkey* h, long long* tc_queries, long long* fallback_queries);

#ifdef __cplusplus
}
#endif
#endif
