/*
 * sc_oracle.h — C entry points of the CPU ORACLE for the Scan Context loop-closure path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the timed CPU baseline. The product (scl_slam_b200/) never
 * links, imports or falls back to this code.
 *
 * What it restates: class scan_context_descriptor of the reference,
 * /root/reference/include/descriptor.h:1304-1801 (each function cites its lines in
 * sc_oracle.cpp). Parity status: the restatement is checked (tests/test_oracle_*.py) against
 *   (1) oracle/_ref — the reference's own class text and its vendored nanoflann compiled here
 *       against stand-in Eigen/PCL/ROS/libnabo headers (oracle/ref_shim/), and
 *   (2) the golden fixtures under tests/golden/ generated from (1).
 * The reference itself ships no tests or golden vectors. Arithmetic that lives
 * in un-vendored third-party code — Eigen's SIMD reduction order, libnabo's kNN tie/self-match
 * rules, PCL's ICP — is "parity unpinned": here the oracle DEFINES the expected values
 * (sequential sums, lowest-index ties, libnabo's published epsilon self-match rule).
 */
#ifndef SC_ORACLE_H_
#define SC_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sco_handle sco_handle;

/* ctor mirror of descriptor.h:1307-1316 */
sco_handle* sco_create(int num_ring, int num_sector, int num_candidates, double dist_thres,
                       double lidar_height, double max_radius, int num_exclude_recent,
                       int tree_making_period, double search_ratio);
void sco_destroy(sco_handle* h);

/* descriptor.h:1404-1461. pts: n points, stride_floats floats apart, x,y,z at 0,1,2.
 * out_desc: R*S row-major float wire vector (may be NULL).
 * out_ring/out_sector: per-point 1-based bin indices, 0 when the point was dropped (may be NULL). */
void sco_make_scancontext(sco_handle* h, const float* pts, int n, int stride_floats,
                          float* out_desc, int* out_ring, int* out_sector);

/* descriptor.h:1604-1611 / 1572-1585. Return the new global key. */
int sco_make_and_save(sco_handle* h, const float* pts, int n, int stride_floats,
                      int8_t robot, int index, float* out_desc);
int sco_save(sco_handle* h, const float* desc_rowmajor, int8_t robot, int index);

/* Append n descriptors (row-major float wires, robot 0, index = key) without per-insert cost.
 * borrow != 0: the oracle keeps pointers into `wires`, which the caller must keep alive. */
int sco_bulk_load(sco_handle* h, const float* wires, int n, int borrow);

int sco_size(sco_handle* h);                                        /* :1763-1766 */
void sco_get_index(sco_handle* h, int key, int* robot, int* index); /* :1758-1761 ({-1,-1} if out of range) */
void sco_get_desc(sco_handle* h, int key, float* out_desc);         /* row-major floats */
void sco_ring_key(sco_handle* h, int key, float* out_R);            /* :1463-1475 */
void sco_sector_key(sco_handle* h, int key, double* out_S);         /* :1477-1489 */

/* descriptor.h:1538-1569 on two stored descriptors. */
void sco_distance(sco_handle* h, int key1, int key2, double* dist, int* shift);
/* same on two raw row-major float descriptors */
void sco_distance_raw(sco_handle* h, const float* d1, const float* d2, double* dist, int* shift);
int sco_fast_align(sco_handle* h, int key1, int key2);              /* :1491-1511 */
double sco_dist_direct(sco_handle* h, int key1, int key2, int shift); /* :1513-1536 on circshift(sc2, shift) */

/* descriptor.h:1613-1674 (libnabo flavour) and :1676-1756 (nanoflann flavour). */
void sco_detect_intra(sco_handle* h, int cur, int* id, float* second);
void sco_detect_inter(sco_handle* h, int cur, int* id, float* second);

/*
 * The candidate stage by itself: K nearest ring keys of entry `cur` among keys [0, n_db).
 * metric 0 = nanoflann L2_Adaptor 4-wide accumulation (nanoflann.hpp:383-408),
 * metric 1 = sequential accumulation + libnabo's "skip d2 <= FLT_EPSILON" self-match rule.
 * Ties: lowest index first. Returns the number found (slots beyond it: id -1, d2 = FLT_MAX).
 */
int sco_knn(sco_handle* h, int cur, int n_db, int k, int metric, int32_t* ids, float* d2);

/*
 * Batched throughput form used by the benches and the large parity tests: for each query
 * (a stored key) the K ring-key neighbours among [0, n_db) in kNN order, each with its
 * shift-aligned SC distance, plus the winner after the strict-< scan (:1721-1737, self skipped).
 * nthreads > 1 partitions the queries over std::threads.
 */
void sco_query_batch(sco_handle* h, const int32_t* queries, int nq, int n_db, int k, int metric,
                     int nthreads, int32_t* cand_ids, float* cand_d2, double* cand_dist,
                     int32_t* cand_shift, int32_t* best_id, double* best_dist, int32_t* best_shift);

/* The float atan used by xy2theta (descriptor.h:1357): this libm's atanf, and the fdlibm
 * restatement the CUDA kernel follows; tests check they agree bit for bit. */
float sco_atanf_libm(float x);
float sco_atanf_port(float x);

/* ---- ICP oracle (icp_oracle.cpp): PCL-default point-to-point ICP as driven by
 * distributedMapping.h:1108-1132. T is row-major 4x4. Returns iterations run. */
int sco_icp(const float* src, int n_src, const float* tgt, int n_tgt, int stride_floats,
            double max_corr_dist, int max_iter, double trans_eps, double fit_eps,
            float* T_out, float* fitness, int* converged);
/* RANSAC + SVD verification restatement (distributedMapping.h:1211-1243). Returns hypotheses tried. */
int sco_verify_ransac(const float* src, int n_src, const float* tgt, int n_tgt, int stride_floats, int max_iter, double inlier_thr,
                      double min_inlier_ratio, unsigned seed, float* T_out, int* n_corr, int* n_inliers, int* success);
/* exact nearest neighbour (brute force) — squared distances and indices */
void sco_nn(const float* src, int n_src, const float* tgt, int n_tgt, int stride_floats,
            int32_t* idx, float* d2);
/* pcl::VoxelGrid restatement (distributedMapping.h:996-998,1183-1184): centroid per leaf.
 * Returns the number of output points written to out (xyz packed, 3 floats). */
int sco_voxel_grid(const float* pts, int n, int stride_floats, float leaf, float* out);

/* ---- cloud preparation (cloud_oracle.cpp) -------------------------------------------------
 * pcl::VoxelGrid<PointXYZI> restated in PCL's float arithmetic (distributedMapping.h:996-998,1181-1185): out holds
 * 4 floats (x, y, z, intensity) per leaf, room for n points; returns the number written. */
int sco_voxel_grid_pcl(const float* pts, int n, int stride_floats, float leaf, float* out);
/* loopFindNearKeyframes (distributedMapping.h:1163-1186): transformPointCloud (:234-253) per cloud, concatenation,
 * VoxelGrid (leaf <= 0: none). poses6 = (x, y, z, roll, pitch, yaw) per cloud. */
int sco_assemble_submap(const float* pts, const int* offsets, int n_clouds, int stride_floats, const float* poses6, float leaf, float* out);

#ifdef __cplusplus
}
#endif
#endif
