// engine_internal.h — the engine object shared by engine.cu and k5_icp.cu (not a public header)
#pragma once
#include "../../include/scl_engine.h"
#include "kernels.h"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() { return static_cast<T*>(p); }
};

struct scl_engine {
    scl_params p;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::mutex mu;
    std::string err;
    int n = 0, cap = 0;
    float *d_desc = nullptr, *d_keys = nullptr, *d_knorm = nullptr;
    double* d_cstat = nullptr;             /* [cap][2*S]: sector key | column norms of every entry (K4's per-entry cache) */
    int scdist_tiles = 0;                  /* scl_set_scdist_tiles: candidate tiles per K4 CTA on an unsharded engine (0 = one per candidate) */
    bool scdist_exact_all = false;         /* scl_set_scdist_mode(1): K4 evaluates every shift in FP64 */
    float* d_kn2max = nullptr;             /* device scalar: largest squared ring-key norm in the database */
    unsigned char* d_kimg = nullptr;       /* tensor-core key image: 128-key tiles in the tcgen05 operand layout (k3_knn_tc.cu) */
    int img_n = 0, img_cap = 0;            /* keys [0, img_n) have an image; capacity in keys */
    /* Hybrid sharding (scl_set_replicated_keys_dev): ALL ring keys, in global key order, on every rank, so that K3 runs
     * query-parallel (this rank's 1 / world of the batch against every key); descriptors stay sharded for K4 */
    float *r_keys = nullptr, *r_knorm = nullptr; unsigned char* r_kimg = nullptr; int r_n = 0;
    int knn_mode = 0;                      /* 0 auto, 1 exact CUDA-core kernel, 2 tensor-core prefilter */
    int tc_stages = 2;                     /* key tiles knn_tc_kernel keeps in flight in shared memory (scl_set_tc_stages): two leave 127 KB of the SM to other lanes' kernels */
    long long stat_tc_queries = 0, stat_fallback_queries = 0;
    bool count_fallbacks = false;
    std::vector<std::pair<int8_t, int>> index;
    int rank = 0, world = 1;
    int tree_counter = 0, n_tree = 0;      /* descriptor.h:1691-1703 */
    int search_radius = 0;                 /* round(0.5*SEARCH_RATIO*S), descriptor.h:1545 */
    /* scratch of the build / insert path (engine stream) */
    DevBuf pts, offsets, gbins, tickets, stage_desc, stage_keys, stage_knorm, bins_ring, bins_sector;
    /* Query lanes: a lane is a CUDA stream plus every scratch buffer one query batch needs, so batches on different lanes
     * run concurrently (the kernels of one batch leave SMs idle: start-up, re-rank, exchange waits). Lane 0 runs on the engine
     * stream and serves the synchronous calls; scl_query_batch_submit and the *_lane calls rotate over all of them. */
    static constexpr int kLanes = 8;
    struct Lane {
        cudaStream_t stream = nullptr;
        DevBuf knn_tickets;
        DevBuf qdesc, qids, qlocal, qkeys, qknorm, qstat, part_ids, part_d2, cand_ids, cand_d2, cand_local, cand_dist, cand_shift,
            best_id, best_dist, best_shift;
        DevBuf tc_queues, tc_queue_cnt, tc_slots, tc_fail_list, tc_fail_count, tc_err_probe;
        long long tc_calls = 0;                /* tensor-core batches so far: they alternate between two fail counters */
        bool tc_state_clean = false;           /* slots and the next fail counter were reset by the last re-rank kernel */
        int tc_slots_rows = 0;                 /* rows of tc_slots known to be clean */
        int* tc_last_fail = nullptr;           /* the counter the last batch used */
        const float* qstat_of = nullptr; int qstat_rows = 0;   /* the query block whose column statistics qstat holds */
        DevBuf x_blob1, x_blob2, x_ids, x_d2;  /* sharded step: this rank's (id, d2) and (dist, shift) blocks, the merged lists */
        int xseq = 0;                          /* exchange steps taken on this lane (identical on all ranks) */
        cudaEvent_t done = nullptr;            /* pipelined host queries */
        bool busy = false;
        int scdist_owned_hint = 0;             /* K4 launch: expected candidates per query held by this shard (0 = all K) */
        DevBuf* all[27] = {&knn_tickets, &qdesc, &qids, &qlocal, &qkeys, &qknorm, &qstat, &part_ids, &part_d2, &cand_ids, &cand_d2, &cand_local,
                           &cand_dist, &cand_shift, &best_id, &best_dist, &best_shift, &tc_queues, &tc_queue_cnt, &tc_slots, &tc_fail_list,
                           &tc_fail_count, &tc_err_probe, &x_blob1, &x_blob2, &x_ids, &x_d2};
    };
    Lane lanes[kLanes];
    cudaEvent_t db_ready = nullptr;        /* recorded on the engine stream after the last change of the database */
    bool db_dirty = false;                 /* the database changed since db_ready was recorded */
    long long pipe_next = 0;               /* tickets of scl_query_batch_submit */
    /* peer-memory exchange (k7_exchange.cu): one buffer, a region per lane */
    void* xchg_buf = nullptr; size_t xchg_bytes = 0; int xchg_qk = 0, xchg_q = 0;
    XchgView xchg[kLanes] = {}; bool xchg_open = false; void* xchg_peer_map[16] = {};
    DevBuf icp_src, icp_tgt, icp_raw, icp_grid[2][5], icp_acc, icp_nn;
    /* device-resident keyframe clouds (robots[id].keyFrameArray of the reference): one append-only arena of packed
     * x, y, z, intensity records; cloud k = points [kf_off[k], kf_off[k + 1]) */
    DevBuf kf_arena; size_t kf_points = 0; std::vector<int> kf_off{0};
    DevBuf vg_in, vg_world, vg_out, vg_keys[2], vg_vals[2], vg_head, vg_ord, vg_temp, vg_misc, vg_T, vg_off;
    size_t gbins_scans = 0;
    /* per-stage event timing */
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[4];
    std::vector<cudaEvent_t> ev_pool;

    int RS() const { return p.num_ring * p.num_sector; }
};

int scl_icp_set_cloud_dev(scl_engine* e, int which, const void* xyzw_dev, int n);
int scl_icp_device(scl_engine* e, int n_src, int n_tgt, const scl_icp_params* prm, float* T_out, float* fitness, int* converged, int* iterations);

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            e->err = std::string(#call) + ": " + cudaGetErrorString(_e);                             \
            return _e == cudaErrorNotSupported ? SCL_ERR_UNSUPPORTED : (_e == cudaErrorMemoryAllocation ? SCL_ERR_NOMEM : SCL_ERR_CUDA); \
        }                                                                                            \
    } while (0)

#define FAIL(code, msg) do { e->err = (msg); return (code); } while (0)


#define LOCK() if (!e) return SCL_ERR_INVALID; std::lock_guard<std::mutex> _lk(e->mu); cudaSetDevice(e->device)
