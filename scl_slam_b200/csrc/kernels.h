// kernels.h — launchers of the sm_100a kernels (internal; the public boundary is include/scl_engine.h)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// K1 (+K2 epilogue): polar binning of a batch of scans. See k1_polar.cu.
// offsets_host (may be null): the same n_scans + 1 offsets on the host; batches of up to scl_polar_inline_scans() scans take them
// from there (kernel parameters) and never read offsets_dev, which may then be null.
int scl_polar_inline_scans();
// host only: the threshold tables polar_bin_kernel keeps in shared memory (ring_thr: up to 63, sec_thr: 4 x 31); 0 = ok
int scl_polar_tables_host(int R, int S, double max_radius, float* ring_thr, int* n_ring, float* s_max,
                          float* sec_thr, int* n_sec, int* sec_base, int* sec_dir);
cudaError_t scl_launch_polar(const void* pts_dev, const int* offsets_dev, const int* offsets_host, int n_scans, int max_points, int stride_bytes,
                             int R, int S, double lidar_height, double max_radius, uint32_t* gbins, int* tickets,
                             float* out_desc, float* out_keys, float* out_knorm, float* kn2max,
                             double* out_cstat /* per scan 2*S doubles (sector key | column norms, as scl_launch_ring_keys writes them) or null */,
                             int* out_ring, int* out_sector, cudaStream_t stream);
// K2: ring keys (+ squared key norms) of descriptors already in device memory.
//     kn2max (device scalar, may be null) is raised to the largest squared norm seen (atomicMax on the float bits).
//     cstat (may be null): per descriptor 2*S doubles, the column means (sector key, descriptor.h:1477-1489) and the column
//     norms (descriptor.h:1521-1522) that K4 would otherwise recompute for every pair. keys may be null (cstat only).
cudaError_t scl_launch_ring_keys(const float* desc_dev, int n, int R, int S, float* keys, float* knorm, float* kn2max, double* cstat,
                                 cudaStream_t stream);

// K3: ring-key kNN. Exact variant (CUDA cores, reference accumulation order). See k3_knn.cu.
//  qkeys [Q][R], keys [n_db][R]; out ids/d2 [Q][K] ascending by (d2, id); id_mul/id_add map local
//  key l to the reported global id l*id_mul + id_add (database sharding).
struct KnnWorkspace {
    int32_t* part_ids;   // [Q][splits][K]
    float* part_d2;      // [Q][splits][K]
    int* tickets;        // [ceil(Q / 128)] zero between launches (the last CTA of a query tile merges its splits)
    size_t capacity;     // in entries
};
int scl_knn_splits(int Q, int n_db);
//  qlist/qcount (device, may be null): only queries qlist[0..*qcount) are processed (fallback of the tensor-core path).
cudaError_t scl_launch_knn_exact(const float* qkeys, int Q, const float* keys, int n_db, int R, int K, int metric,
                                 int id_mul, int id_add, const int32_t* qlist, const int* qcount, KnnWorkspace ws,
                                 int32_t* out_ids, float* out_d2, cudaStream_t stream);

// K3 on tcgen05 (k3_knn_tc.cu): BF16x3 prefilter fed by TMA from a pre-split key image (scl_launch_key_image, 128-key
// tiles in the tcgen05 shared-memory layout), union-bound thresholds, hit queues, exact re-rank + certificate. Queries whose
// top-K could not be certified are appended to fail_list (count in *fail_count) for scl_launch_knn_exact.
struct KnnTcWorkspace {
    uint32_t* hq;        // [Qc][ranges] hit queues of scl_knn_tc_queue_bytes() each (Qc = min(Q, scl_knn_tc_max_batch()))
    int* hq_cnt;         // [Qc][ranges] groups queued
    int* slots;          // [Qc][scl_knn_tc_slot_stride()]: K' range minima by range % K' (K' <= scl_knn_tc_kprime()), then the query's direct bound
    float* err_probe;    // null, or one float raised to the largest |prefilter score error| / eps seen (tests)
    size_t capacity;     // in (query, range) pairs
};
bool scl_knn_tc_supported(int R);
int scl_knn_tc_ranges(int Q);
int scl_knn_tc_max_batch();
int scl_knn_tc_kprime();
int scl_knn_tc_slot_stride();       /* ints per query in KnnTcWorkspace::slots */
size_t scl_knn_tc_queue_bytes();
size_t scl_knn_tc_image_bytes(int R, int n_keys);
cudaError_t scl_launch_key_image(const float* keys, const float* knorm, int k_lo, int k_hi, int R, unsigned char* img, cudaStream_t stream);
// squared norms of keys [k_lo, k_hi) (prefilter only) and their maximum (kn2max, a device float raised by atomicMax on its bits); rowkey.cu
cudaError_t scl_launch_key_norms(const float* keys, int k_lo, int k_hi, int R, float* knorm, float* kn2max, cudaStream_t stream);
cudaError_t scl_launch_knn_tc(const float* qkeys, int Q, const float* keys, const unsigned char* img, const float* kn2max, int n_db, int R, int K,
                               int metric, int id_mul, int id_add, KnnTcWorkspace ws, int32_t* out_ids, float* out_d2,
                               int32_t* fail_list, int* fail_count, int* next_fail_count /* zeroed by the re-rank for the next call */,
                               bool init_state /* slots and fail_count are not known to be clean */,
                               int stages /* key tiles in flight in shared memory, 2..5: fewer leave room for kernels of other lanes on the SM */,
                               cudaStream_t stream);

// K4: shift-aligned column-cosine distance for every (query, candidate) + winner scan. See k4_scdist.cu.
//  q_desc [Q][R*S] with q_stat [Q][2*S] (scl_launch_ring_keys' cstat of the queries), or both nullptr (then queries are db
//  entries q_local[i]); db_stat [n][2*S] the per-entry cache; cand_local [Q][K] local keys (-1 = none);
//  cand_ids [Q][K] reported ids (for the self-skip rule against q_ids). exact_all != 0: every shift in FP64 (no FP32 prefilter).
cudaError_t scl_launch_scdist(const float* db_desc, const double* db_stat, const float* q_desc, const double* q_stat,
                              const int32_t* q_local, const int32_t* q_ids,
                              const int32_t* cand_local /* null: derived from cand_ids with id = local * id_mul + id_add */, const int32_t* cand_ids,
                              int id_mul, int id_add, int Q, int K, int R, int S, int search_radius,
                              double* cand_dist, int32_t* cand_shift, int32_t* best_id, double* best_dist, int32_t* best_shift,
                              int owned_per_query, int exact_all, cudaStream_t stream);

// Shard merge (multi-GPU): world blocks of [Q][K] records -> global top-K by (d2,id) + winner scan.
cudaError_t scl_launch_merge_shards(int world, int Q, int K, const int32_t* q_ids, const int32_t* all_ids, const float* all_d2,
                                    const double* all_dist, const int32_t* all_shift, int32_t* out_ids, float* out_d2,
                                    double* out_dist, int32_t* out_shift, int32_t* best_id, double* best_dist, int32_t* best_shift,
                                    cudaStream_t stream);

// Two-phase multi-GPU exchange: rank blocks are `rank_stride_bytes` apart (e.g. an all-gather buffer).
cudaError_t scl_launch_merge_topk(int world, int Q, int K, const void* ids_base, const void* d2_base, size_t rank_stride_bytes,
                                  int32_t* out_ids, float* out_d2, cudaStream_t stream);
cudaError_t scl_launch_combine_owned(int world, int Q, int K, const int32_t* q_ids, const int32_t* cand_ids, const void* dist_base,
                                     const void* shift_base, size_t rank_stride_bytes, double* out_dist, int32_t* out_shift,
                                     int32_t* best_id, double* best_dist, int32_t* best_shift, cudaStream_t stream);

// small helpers
cudaError_t scl_launch_gather_rows(const float* src, const int32_t* rows, int n, int width, float* dst, cudaStream_t stream);
// reported id -> local key ((id - id_add) / id_mul; -1 if the id belongs to another shard); missing (-1) entries
// become `missing_to`, and are also rewritten in ids_rewrite when that is non-null
cudaError_t scl_launch_ids_to_local(const int32_t* ids, int n, int id_mul, int id_add, int missing_to, int32_t* ids_rewrite,
                                    int32_t* local, cudaStream_t stream);

// K6 (k6_cloud.cu): cloud preparation — transformPointCloud + concat (distributedMapping.h:234-253,1163-1175) and
// pcl::VoxelGrid (:996-998,1181-1185). Points are 16-byte aligned x,y,z,intensity records `stride_bytes` apart; outputs
// are packed float4 (x, y, z, intensity).
cudaError_t scl_launch_transform_concat(const void* pts, const int* offsets_dev, int n_clouds, int max_points, int stride_bytes,
                                        const float* T_dev /* [n_clouds][12] */, void* out_xyzi, cudaStream_t stream);
cudaError_t scl_launch_cloud_bounds(const void* pts, int n, int stride_bytes, int* bounds6 /* order-preserving int images */, cudaStream_t stream);
size_t scl_voxel_temp_bytes(int n);
cudaError_t scl_launch_voxel_grid(const void* pts, int n, int stride_bytes, float inv_leaf, const int* min_b, int div0, int div01, int key_bits,
                                  uint32_t* keys_a, uint32_t* keys_b, int* vals_a, int* vals_b, int* head, int* ord,
                                  void* temp, size_t temp_bytes, void* out_xyzi, int* n_out, cudaStream_t stream);
cudaError_t scl_launch_pack_xyzi(const void* pts, int n, int stride_bytes, void* out_xyzi, cudaStream_t stream);

// K7 (k7_exchange.cu): the sharded query's two exchange points over NVLink peer memory, fused with their merge kernels.
// Every rank's exchange buffer has the same layout; peer[r] is rank r's buffer as mapped into this process.
struct XchgView {
    unsigned char* peer[16];
    int world, rank;
    size_t data_off[3], slot_bytes[3];    /* points 0, 1: [2 parities][world] slots; point 2 (query gather): [2 parities] areas */
    size_t flag_off;                      /* int flags[3 points][16 ranks] */
    size_t ticket_off;                    /* int tickets[3] */
};
cudaError_t scl_launch_xchg_gather_queries(const XchgView& x, int seq, const void* my_rows, size_t row0_bytes, size_t bytes, cudaStream_t stream);
cudaError_t scl_launch_xchg_merge_topk(const XchgView& x, int seq, int Q, int K, const void* my_block, int32_t* out_ids, float* out_d2, cudaStream_t stream);
cudaError_t scl_launch_xchg_combine(const XchgView& x, int seq, int Q, int K, const void* my_block, const int32_t* q_ids, const int32_t* cand_ids,
                                    double* out_dist, int32_t* out_shift, int32_t* best_id, double* best_dist, int32_t* best_shift, cudaStream_t stream);

// load every kernel of the query path on the current device (see SCL_TOUCH in common.cuh); called by scl_create
void scl_preload_k1();
void scl_preload_k3();
void scl_preload_k3_tc();
void scl_preload_k3_tc80();   /* the 80-row variant (row-key family, rowkey.cu) */
void scl_preload_k4();
void scl_preload_k7();
