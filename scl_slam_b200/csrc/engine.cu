// engine.cu — the C-ABI of include/scl_engine.h over the sm_100a kernels.
//
// Host-side mirror of class scan_context_descriptor (/root/reference/include/descriptor.h:
// 1304-1801): same constructor parameters, same insert / query / accessor semantics, with the
// keyframe database resident in HBM:
//   d_desc  [cap][R*S] float32  row-major wire image of every descriptor (the reference keeps
//                               MatrixXd; every value is a float widened to double, so FP32 is lossless)
//   d_keys  [cap][R]   float32  ring keys, one contiguous row per keyframe (the reference: one
//                               COLUMN per keyframe in polarcontextRowKey, re-allocated per insert)
//   d_knorm [cap]      float32  squared key norms for the tensor-core prefilter
// (robot, index) metadata stays on the host (descriptor.h:1599,1758-1761).
// There is no CPU fallback anywhere in this file: every data-path call launches kernels.
#include "engine_internal.h"
#include "common.cuh"

namespace {

typedef scl_engine::Lane Lane;

struct StageTimer {
    scl_engine* e; int stage; cudaStream_t s; cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(scl_engine* e_, int stage_, cudaStream_t s_) : e(e_), stage(stage_), s(s_)
    {
        if (!e->profiling) return;
        a = take(); b = take();
        cudaEventRecord(a, s);
    }
    ~StageTimer()
    {
        if (!a) return;
        cudaEventRecord(b, s);
        e->ev[stage].emplace_back(a, b);
    }
    cudaEvent_t take()
    {
        if (!e->ev_pool.empty()) { cudaEvent_t x = e->ev_pool.back(); e->ev_pool.pop_back(); return x; }
        cudaEvent_t x; cudaEventCreate(&x); return x;
    }
};

int grow(scl_engine* e, int need)
{
    if (need <= e->cap) return SCL_OK;
    int ncap = e->cap ? e->cap : 1024;
    while (ncap < need) ncap = ncap < (1 << 28) ? ncap * 2 : ncap + (1 << 26);
    const size_t RS = e->RS(), R = e->p.num_ring;
    const size_t S2 = 2 * (size_t)e->p.num_sector;
    float *nd = nullptr, *nk = nullptr, *nn = nullptr; double* nc = nullptr;
    if (cudaMalloc(&nd, (size_t)ncap * RS * 4) != cudaSuccess || cudaMalloc(&nk, (size_t)ncap * R * 4) != cudaSuccess ||
        cudaMalloc(&nn, (size_t)ncap * 4) != cudaSuccess || cudaMalloc(&nc, (size_t)ncap * S2 * 8) != cudaSuccess) {
        cudaFree(nd); cudaFree(nk); cudaFree(nn); cudaFree(nc);          /* the database stays as it was */
        (void)cudaGetLastError();
        FAIL(SCL_ERR_CUDA, "out of device memory growing the keyframe database");
    }
    if (e->n > 0) {
        CK(cudaMemcpyAsync(nd, e->d_desc, (size_t)e->n * RS * 4, cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaMemcpyAsync(nk, e->d_keys, (size_t)e->n * R * 4, cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaMemcpyAsync(nn, e->d_knorm, (size_t)e->n * 4, cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaMemcpyAsync(nc, e->d_cstat, (size_t)e->n * S2 * 8, cudaMemcpyDeviceToDevice, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    for (int l = 1; l < scl_engine::kLanes; l++) if (e->lanes[l].stream) CK(cudaStreamSynchronize(e->lanes[l].stream));   /* queries in flight read the old arrays */
    cudaFree(e->d_desc); cudaFree(e->d_keys); cudaFree(e->d_knorm); cudaFree(e->d_cstat);
    e->d_desc = nd; e->d_keys = nk; e->d_knorm = nn; e->d_cstat = nc; e->cap = ncap;
    return SCL_OK;
}

void append_index(scl_engine* e, int n, const int8_t* robots, const int32_t* indices)
{
    for (int i = 0; i < n; i++)
        e->index.emplace_back(robots ? robots[i] : (int8_t)0, indices ? indices[i] : e->n + i);
}

// polar binning of a batch already in device memory
int build_dev(scl_engine* e, const void* pts_dev, const int32_t* offsets_host, int n_scans, int stride_bytes, int insert,
              const int8_t* robots, const int32_t* indices, float* out_desc_dev, int* ring_dev, int* sector_dev,
              float** where_desc)
{
    const int R = e->p.num_ring, S = e->p.num_sector;
    const size_t RS = e->RS();
    if (n_scans <= 0) return SCL_OK;
    int max_points = 0;
    for (int i = 0; i < n_scans; i++) {
        const int c = offsets_host[i + 1] - offsets_host[i];
        if (c < 0) FAIL(SCL_ERR_INVALID, "offsets must be non-decreasing");
        if (c > max_points) max_points = c;
    }
    if (insert) { int rc = grow(e, e->n + n_scans); if (rc) return rc; }
    const bool inline_offsets = n_scans <= scl_polar_inline_scans();     /* the offsets travel as kernel parameters: no copy in front of the launch */
    if (!inline_offsets) {
        CK(e->offsets.ensure((size_t)(n_scans + 1) * 4));
        CK(cudaMemcpyAsync(e->offsets.p, offsets_host, (size_t)(n_scans + 1) * 4, cudaMemcpyHostToDevice, e->stream));
    }
    if ((size_t)n_scans > e->gbins_scans) {
        CK(e->gbins.ensure((size_t)n_scans * RS * 4));
        CK(e->tickets.ensure((size_t)n_scans * 4));
        CK(cudaMemsetAsync(e->gbins.p, 0, e->gbins.cap, e->stream));   /* kernels leave it zeroed afterwards */
        CK(cudaMemsetAsync(e->tickets.p, 0, e->tickets.cap, e->stream));
        e->gbins_scans = e->gbins.cap / (RS * 4);
        if (e->tickets.cap / 4 < e->gbins_scans) e->gbins_scans = e->tickets.cap / 4;
    }
    float *od, *ok, *on; double* oc = nullptr;
    if (insert) {
        od = e->d_desc + (size_t)e->n * RS; ok = e->d_keys + (size_t)e->n * R; on = e->d_knorm + e->n;
        oc = e->d_cstat + (size_t)e->n * 2 * S;
    } else {
        if (!out_desc_dev) CK(e->stage_desc.ensure((size_t)n_scans * RS * 4));
        CK(e->stage_keys.ensure((size_t)n_scans * R * 4));
        CK(e->stage_knorm.ensure((size_t)n_scans * 4));
        od = out_desc_dev ? out_desc_dev : e->stage_desc.as<float>();      /* no insert: the kernel writes the caller's buffer directly */
        ok = e->stage_keys.as<float>(); on = e->stage_knorm.as<float>();
    }
    StageTimer st(e, 3, e->stream);
    e->db_dirty = e->db_dirty || insert;
    CK(scl_launch_polar(pts_dev, inline_offsets ? nullptr : e->offsets.as<int>(), offsets_host, n_scans, max_points, stride_bytes, R, S, e->p.lidar_height, e->p.max_radius,
                        e->gbins.as<uint32_t>(), e->tickets.as<int>(), od, ok, on, insert ? e->d_kn2max : nullptr,
                        oc /* the per-entry cache K4 reads: sector key + column norms of the new entries (descriptor.h:1541-1542 recomputes them per pair), written by the kernel's epilogue */,
                        ring_dev, sector_dev, e->stream));
    if (out_desc_dev && od != out_desc_dev) CK(cudaMemcpyAsync(out_desc_dev, od, (size_t)n_scans * RS * 4, cudaMemcpyDeviceToDevice, e->stream));
    if (where_desc) *where_desc = od;
    if (insert) { append_index(e, n_scans, robots, indices); e->n += n_scans; }
    return SCL_OK;
}

int build_host(scl_engine* e, const void* pts, const int32_t* offsets, int n_scans, int stride_bytes, int insert,
               const int8_t* robots, const int32_t* indices, float* out_desc, int32_t* out_ring, int32_t* out_sector)
{
    if (n_scans < 0 || stride_bytes < 12 || (stride_bytes & 3)) FAIL(SCL_ERR_INVALID, "stride_bytes must be >= 12 and a multiple of 4");
    if (n_scans == 0) return SCL_OK;
    if (!offsets) FAIL(SCL_ERR_INVALID, "offsets is NULL");
    const size_t total = (size_t)offsets[n_scans];
    if (total > 0 && !pts) FAIL(SCL_ERR_INVALID, "pts is NULL");
    /* the copy covers whole strides except after the last point, whose tail may not exist */
    const size_t bytes = total ? (total - 1) * (size_t)stride_bytes + 12 : 0;
    CK(e->pts.ensure(total * (size_t)stride_bytes + 16));
    if (bytes) CK(cudaMemcpyAsync(e->pts.p, pts, bytes, cudaMemcpyHostToDevice, e->stream));
    int *ring = nullptr, *sector = nullptr;
    if (out_ring && out_sector && total) {
        CK(e->bins_ring.ensure(total * 4)); CK(e->bins_sector.ensure(total * 4));
        ring = e->bins_ring.as<int>(); sector = e->bins_sector.as<int>();
    }
    float* where = nullptr;
    int rc = build_dev(e, e->pts.p, offsets, n_scans, stride_bytes, insert, robots, indices, nullptr, ring, sector, &where);
    if (rc) return rc;
    if (out_desc) CK(cudaMemcpyAsync(out_desc, where, (size_t)n_scans * e->RS() * 4, cudaMemcpyDeviceToHost, e->stream));
    if (ring) {
        CK(cudaMemcpyAsync(out_ring, ring, total * 4, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(out_sector, sector, total * 4, cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    return SCL_OK;
}

// The tensor-core key image follows the database lazily: keys appended since the last tensor-core query are
// converted before the next one (same stream, so ordered after the inserts that produced them).
int sync_key_image(scl_engine* e)
{
    const int R = e->p.num_ring;
    if (e->img_cap < e->cap) {
        /* derived data: on growth the image is simply rebuilt from the keys */
        if (e->d_kimg) {
            CK(cudaStreamSynchronize(e->stream));
            for (int l = 1; l < scl_engine::kLanes; l++) if (e->lanes[l].stream) CK(cudaStreamSynchronize(e->lanes[l].stream));
            cudaFree(e->d_kimg); e->d_kimg = nullptr;
        }
        const size_t bytes = scl_knn_tc_image_bytes(R, e->cap);
        CK(cudaMalloc(&e->d_kimg, bytes));
        CK(cudaMemsetAsync(e->d_kimg, 0, bytes, e->stream));
        e->img_cap = e->cap; e->img_n = 0;
    }
    if (e->img_n < e->n) {
        CK(scl_launch_key_image(e->d_keys, e->d_knorm, e->img_n, e->n, R, e->d_kimg, e->stream));
        e->img_n = e->n;
        e->db_dirty = true;
    }
    return SCL_OK;
}

// A lane other than lane 0 runs on its own stream: it must see every insert (and key-image update) enqueued on the engine
// stream so far. One event, re-recorded only when the database changed since it was last recorded.
int lane_begin(scl_engine* e, Lane& ln)
{
    if (!ln.stream) {
        if (&ln == &e->lanes[0]) ln.stream = e->stream;
        else CK(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
    }
    if (&ln == &e->lanes[0]) { ln.stream = e->stream; return SCL_OK; }
    if (!e->db_ready) { CK(cudaEventCreateWithFlags(&e->db_ready, cudaEventDisableTiming)); e->db_dirty = true; }
    if (e->db_dirty) { CK(cudaEventRecord(e->db_ready, e->stream)); e->db_dirty = false; }
    CK(cudaStreamWaitEvent(ln.stream, e->db_ready, 0));
    return SCL_OK;
}

// K2 (query ring keys) + K3 (ring-key kNN) on device pointers; results: reported ids + float distances
int knn_dev(scl_engine* e, Lane& ln, const float* q_desc, const int32_t* q_ids, int Q, int K, int n_db, int metric,
            int32_t* cand_ids, float* cand_d2, int32_t** q_local_out)
{
    const int R = e->p.num_ring, S = e->p.num_sector;
    if (q_local_out) *q_local_out = nullptr;
    if (Q <= 0) return SCL_OK;
    if (K < 1 || K > 32) FAIL(SCL_ERR_INVALID, "K must be in 1..32");
    if (metric != 0 && metric != 1) FAIL(SCL_ERR_INVALID, "metric must be 0 or 1");
    const int n_db_in = n_db;
    if (n_db < 0) n_db = 0;
    if (n_db > e->n) n_db = e->n;
    if (!q_desc && !q_ids) FAIL(SCL_ERR_INVALID, "q_desc and q_ids are both NULL");
    if (!q_desc && e->world != 1) FAIL(SCL_ERR_INVALID, "queries by key need q_desc on a sharded engine");
    /* K3 variant: on a database worth streaming the tensor-core prefilter wins from four queries up (measured on 1 M keys:
     * 182 us per call at Q = 9..128 against 800..1800 us for the exact kernel, and against 199 / 234 us for the thread-per-key
     * kernel at Q = 4 / 8; on 20 k keys and batches of 1024: 15.1 M against 5.0 M queries/s for the whole query); up to three queries and
     * databases under 16 384 keys (fewer key tiles than the union bound needs ranges) take the exact CUDA-core kernels. */
    /* Hybrid sharding: with every ring key on every rank, this rank searches ALL keys (n_db is then the GLOBAL bound) for its
     * 1 / world of the batch and reports global ids; the other queries' lists stay empty (-1) and the exchange's per-query merge
     * of `world` lists returns the owner's. No per-rank K3 on all Q queries, no re-rank of all Q on every rank. */
    const bool hybrid = e->world > 1 && e->r_n > 0 && q_desc != nullptr;
    int q_lo = 0, Qk = Q, nk = n_db, id_mul = e->world, id_add = e->rank;
    const float* keys = e->d_keys; const unsigned char* kimg = e->d_kimg; const float* kn2max = e->d_kn2max;
    if (hybrid) {
        const int per = (Q + e->world - 1) / e->world;
        q_lo = e->rank * per < Q ? e->rank * per : Q;
        Qk = Q - q_lo < per ? Q - q_lo : per;
        nk = n_db_in < e->r_n ? (n_db_in < 0 ? 0 : n_db_in) : e->r_n;
        keys = e->r_keys; kimg = e->r_kimg; kn2max = e->d_kn2max + 1; id_mul = 1; id_add = 0;
    }
    const bool use_tc = scl_knn_tc_supported(R) && K <= scl_knn_tc_kprime() - 2 &&
                        (e->knn_mode == 2 || (e->knn_mode == 0 && Qk > 3 && nk >= 16384));
    if (use_tc && !hybrid) { int rc = sync_key_image(e); if (rc) return rc; }            /* on the engine stream */
    if (!hybrid) kimg = e->d_kimg;                   /* sync_key_image may have (re)allocated it */
    { int rc = lane_begin(e, ln); if (rc) return rc; }
    const size_t QK = (size_t)Q * K;
    CK(ln.cand_local.ensure(QK * 4));
    CK(ln.qkeys.ensure((size_t)Q * R * 4));
    int32_t* q_local = nullptr;
    if (q_desc) {
        CK(ln.qknorm.ensure((size_t)Q * 4));
        CK(ln.qstat.ensure((size_t)Q * 2 * S * 8));
    } else {
        CK(ln.qlocal.ensure((size_t)Q * 4));
        q_local = ln.qlocal.as<int32_t>();
    }
    const int splits = scl_knn_splits(Qk > 0 ? Qk : 1, nk);
    KnnWorkspace ws;
    CK(ln.part_ids.ensure((size_t)Q * splits * K * 4));
    CK(ln.part_d2.ensure((size_t)Q * splits * K * 4));
    {
        const void* old = ln.knn_tickets.p;
        CK(ln.knn_tickets.ensure(((size_t)Q / 128 + 16) * 4));          /* per 128-query tile, or per query of a handful */
        if (old != ln.knn_tickets.p) CK(cudaMemsetAsync(ln.knn_tickets.p, 0, ln.knn_tickets.cap, ln.stream));   /* the kernel leaves them zero */
    }
    ws.part_ids = ln.part_ids.as<int32_t>(); ws.part_d2 = ln.part_d2.as<float>(); ws.tickets = ln.knn_tickets.as<int>(); ws.capacity = (size_t)Q * splits * K;   /* sized for all Q; a hybrid rank uses Qk of them */
    if (q_desc) {
        StageTimer st(e, 0, ln.stream);
        CK(scl_launch_ring_keys(q_desc, Q, R, S, ln.qkeys.as<float>(), ln.qknorm.as<float>(), nullptr, ln.qstat.as<double>(), ln.stream));
        ln.qstat_of = q_desc; ln.qstat_rows = Q;
    } else {
        CK(scl_launch_ids_to_local(q_ids, Q, e->world, e->rank, -1, nullptr, q_local, ln.stream));
        CK(scl_launch_gather_rows(e->d_keys, q_local, Q, R, ln.qkeys.as<float>(), ln.stream));
    }
    if (hybrid) CK(cudaMemsetAsync(cand_ids, 0xff, QK * 4, ln.stream));            /* -1: the lists of the other ranks' queries are empty */
    const float* qk = ln.qkeys.as<float>() + (size_t)q_lo * R;
    int32_t* o_ids = cand_ids + (size_t)q_lo * K; float* o_d2 = cand_d2 + (size_t)q_lo * K;
    if (Qk > 0) {
        StageTimer st(e, 1, ln.stream);
        if (use_tc) {
            const int Qc = Qk < scl_knn_tc_max_batch() ? Qk : scl_knn_tc_max_batch();
            const int ranges = scl_knn_tc_ranges(Qc);
            const size_t pairs = (size_t)Qc * ranges;
            CK(ln.tc_queues.ensure(pairs * scl_knn_tc_queue_bytes())); CK(ln.tc_queue_cnt.ensure(pairs * 4));
            const void* old_slots = ln.tc_slots.p; const void* old_cnt = ln.tc_fail_count.p;
            CK(ln.tc_fail_list.ensure((size_t)Q * 4)); CK(ln.tc_fail_count.ensure(128));
            CK(ln.tc_slots.ensure((size_t)Qc * scl_knn_tc_slot_stride() * 4));
            const bool init_state = !ln.tc_state_clean || old_slots != ln.tc_slots.p || old_cnt != ln.tc_fail_count.p || Qc > ln.tc_slots_rows;
            if (old_cnt != ln.tc_fail_count.p) CK(cudaMemsetAsync(ln.tc_fail_count.p, 0, 128, ln.stream));   /* two call counters + the running total */
            /* counter block (ints): [0] / [16] = the fail counters of even / odd calls; [8] + [24] = uncertified queries since
             * creation (the re-rank kernel adds to the int 8 places after the counter it is told to zero) */
            int* fail_cur = ln.tc_fail_count.as<int>() + 16 * (ln.tc_calls & 1);
            int* fail_next = ln.tc_fail_count.as<int>() + 16 * ((ln.tc_calls + 1) & 1);
            ln.tc_calls++;
            ln.tc_state_clean = false;
            float* probe = nullptr;
            if (e->count_fallbacks) {
                if (!ln.tc_err_probe.p) { CK(ln.tc_err_probe.ensure(64)); CK(cudaMemsetAsync(ln.tc_err_probe.p, 0, 64, ln.stream)); }
                probe = ln.tc_err_probe.as<float>();
            }
            KnnTcWorkspace tw{ln.tc_queues.as<uint32_t>(), ln.tc_queue_cnt.as<int>(), ln.tc_slots.as<int>(), probe, pairs};
            CK(scl_launch_knn_tc(qk, Qk, keys, kimg, kn2max, nk, R, K, metric, id_mul, id_add, tw,
                                 o_ids, o_d2, ln.tc_fail_list.as<int32_t>(), fail_cur, fail_next, init_state, e->tc_stages, ln.stream));
            ln.tc_state_clean = true; ln.tc_slots_rows = Qc;
            /* uncertified queries (normally none) are redone exactly; CTAs beyond the list length exit at once */
            CK(scl_launch_knn_exact(qk, Qk, keys, nk, R, K, metric, id_mul, id_add,
                                    ln.tc_fail_list.as<int32_t>(), fail_cur, ws, o_ids, o_d2, ln.stream));
            ln.tc_last_fail = fail_cur;
            e->stat_tc_queries += Qk;
        } else {
            CK(scl_launch_knn_exact(qk, Qk, keys, nk, R, K, metric, id_mul, id_add, nullptr, nullptr, ws, o_ids, o_d2, ln.stream));
        }
    }
    if (use_tc && e->count_fallbacks) {
        int nfail = 0;
        CK(cudaMemcpyAsync(&nfail, ln.tc_last_fail, 4, cudaMemcpyDeviceToHost, ln.stream));
        CK(cudaStreamSynchronize(ln.stream));            /* developer mode: surfaces kernel faults at the call that caused them */
        (void)nfail;
    }
    if (q_local_out) *q_local_out = q_local;
    return SCL_OK;
}

// K4 on device pointers: SC distance of every (query, candidate) this engine owns (+ the winner scan)
int scdist_dev(scl_engine* e, Lane& ln, const float* q_desc, const int32_t* q_local, const int32_t* q_ids, int Q, int K, int32_t* cand_ids,
               int missing_to_zero, double* cand_dist, int32_t* cand_shift, int32_t* best_id, double* best_dist, int32_t* best_shift,
               bool q_stat_ready)
{
    const int R = e->p.num_ring, S = e->p.num_sector;
    if (Q <= 0) return SCL_OK;
    { int rc = lane_begin(e, ln); if (rc) return rc; }
    const size_t QK = (size_t)Q * K;
    /* reported id -> row of this engine's descriptor array; on an unsharded engine the two are the same (missing = -1) */
    const int32_t* cand_local = e->world != 1 ? nullptr : cand_ids;     /* a shard derives the local keys inside K4 */
    if (missing_to_zero) {
        CK(ln.cand_local.ensure(QK * 4));
        CK(scl_launch_ids_to_local(cand_ids, (int)QK, e->world, e->rank, missing_to_zero ? 0 : -1, missing_to_zero ? cand_ids : nullptr,
                                   ln.cand_local.as<int32_t>(), ln.stream));
        cand_local = ln.cand_local.as<int32_t>();
    }
    const double* q_stat = nullptr;
    if (q_desc) {
        if (!q_stat_ready) {
            /* queries that did not come through knn_dev on this lane: their column statistics first */
            CK(ln.qstat.ensure((size_t)Q * 2 * S * 8));
            CK(scl_launch_ring_keys(q_desc, Q, R, S, nullptr, nullptr, nullptr, ln.qstat.as<double>(), ln.stream));
            ln.qstat_of = q_desc; ln.qstat_rows = Q;
        }
        q_stat = ln.qstat.as<double>();
    }
    StageTimer st(e, 2, ln.stream);
    CK(scl_launch_scdist(e->d_desc, e->d_cstat, q_desc, q_stat, q_local, q_ids, cand_local, cand_ids, e->world, e->rank, Q, K, R, S, e->search_radius,
                         cand_dist, cand_shift, best_id, best_dist, best_shift, e->world > 1 ? ln.scdist_owned_hint : e->scdist_tiles, e->scdist_exact_all ? 1 : 0, ln.stream));
    return SCL_OK;
}

// the batched query on device pointers; any result pointer may be null (scratch is used)
int query_dev(scl_engine* e, Lane& ln, const float* q_desc, const int32_t* q_ids, int Q, int K, int n_db, int metric, int missing_to_zero,
              int32_t* cand_ids, float* cand_d2, double* cand_dist, int32_t* cand_shift,
              int32_t* best_id, double* best_dist, int32_t* best_shift)
{
    if (Q <= 0) return SCL_OK;
    if (K < 1 || K > 32) FAIL(SCL_ERR_INVALID, "K must be in 1..32");
    const size_t QK = (size_t)Q * K;
    if (!cand_ids) { CK(ln.cand_ids.ensure(QK * 4)); cand_ids = ln.cand_ids.as<int32_t>(); }
    if (!cand_d2) { CK(ln.cand_d2.ensure(QK * 4)); cand_d2 = ln.cand_d2.as<float>(); }
    int32_t* q_local = nullptr;
    int rc = knn_dev(e, ln, q_desc, q_ids, Q, K, n_db, metric, cand_ids, cand_d2, &q_local);
    if (rc) return rc;
    return scdist_dev(e, ln, q_desc, q_local, q_ids, Q, K, cand_ids, missing_to_zero, cand_dist, cand_shift, best_id, best_dist, best_shift,
                      /* q_stat_ready = */ q_desc != nullptr);
}

// device-side part of a host-buffer query: kernels + the device-to-host copies of whatever r asks for (no sync)
int query_host_enqueue(scl_engine* e, Lane& ln, const float* dq, const int32_t* di, const scl_batch_query* q, scl_batch_result* r, int missing_to_zero)
{
    const int Q = q->Q, K = q->K;
    const size_t QK = (size_t)Q * K;
    CK(ln.cand_ids.ensure(QK * 4)); CK(ln.cand_d2.ensure(QK * 4)); CK(ln.cand_dist.ensure(QK * 8)); CK(ln.cand_shift.ensure(QK * 4));
    CK(ln.best_id.ensure((size_t)Q * 4)); CK(ln.best_dist.ensure((size_t)Q * 8)); CK(ln.best_shift.ensure((size_t)Q * 4));
    int rc = query_dev(e, ln, dq, di, Q, K, q->n_db, q->metric, missing_to_zero, ln.cand_ids.as<int32_t>(), ln.cand_d2.as<float>(),
                       ln.cand_dist.as<double>(), ln.cand_shift.as<int32_t>(), ln.best_id.as<int32_t>(), ln.best_dist.as<double>(),
                       ln.best_shift.as<int32_t>());
    if (rc) return rc;
    if (r->cand_ids) CK(cudaMemcpyAsync(r->cand_ids, ln.cand_ids.p, QK * 4, cudaMemcpyDeviceToHost, ln.stream));
    if (r->cand_d2) CK(cudaMemcpyAsync(r->cand_d2, ln.cand_d2.p, QK * 4, cudaMemcpyDeviceToHost, ln.stream));
    if (r->cand_dist) CK(cudaMemcpyAsync(r->cand_dist, ln.cand_dist.p, QK * 8, cudaMemcpyDeviceToHost, ln.stream));
    if (r->cand_shift) CK(cudaMemcpyAsync(r->cand_shift, ln.cand_shift.p, QK * 4, cudaMemcpyDeviceToHost, ln.stream));
    if (r->best_id) CK(cudaMemcpyAsync(r->best_id, ln.best_id.p, (size_t)Q * 4, cudaMemcpyDeviceToHost, ln.stream));
    if (r->best_dist) CK(cudaMemcpyAsync(r->best_dist, ln.best_dist.p, (size_t)Q * 8, cudaMemcpyDeviceToHost, ln.stream));
    if (r->best_shift) CK(cudaMemcpyAsync(r->best_shift, ln.best_shift.p, (size_t)Q * 4, cudaMemcpyDeviceToHost, ln.stream));
    return SCL_OK;
}

int check_host_query(scl_engine* e, const scl_batch_query* q, scl_batch_result* r)
{
    if (!q || !r) FAIL(SCL_ERR_INVALID, "null query/result");
    if (q->Q <= 0) return SCL_OK;
    if (q->K < 1 || q->K > 32) FAIL(SCL_ERR_INVALID, "K must be in 1..32");
    if (q->q_ids && !q->q_desc)
        for (int i = 0; i < q->Q; i++)
            if (q->q_ids[i] < 0 || q->q_ids[i] >= e->n) FAIL(SCL_ERR_RANGE, "query key out of range");
    return SCL_OK;
}

// host queries -> the lane's staging buffers (on the lane's stream: the copy of one lane overlaps the kernels of the others)
int stage_host_query(scl_engine* e, Lane& ln, const scl_batch_query* q, const float** dq, const int32_t** di)
{
    const int Q = q->Q;
    const size_t RS = e->RS();
    *dq = nullptr; *di = nullptr;
    { int rc = lane_begin(e, ln); if (rc) return rc; }
    if (q->q_desc) {
        CK(ln.qdesc.ensure((size_t)Q * RS * 4));
        CK(cudaMemcpyAsync(ln.qdesc.p, q->q_desc, (size_t)Q * RS * 4, cudaMemcpyHostToDevice, ln.stream));
        *dq = ln.qdesc.as<float>();
    }
    if (q->q_ids) {
        CK(ln.qids.ensure((size_t)Q * 4));
        CK(cudaMemcpyAsync(ln.qids.p, q->q_ids, (size_t)Q * 4, cudaMemcpyHostToDevice, ln.stream));
        *di = ln.qids.as<int32_t>();
    }
    return SCL_OK;
}

int query_host(scl_engine* e, const scl_batch_query* q, scl_batch_result* r, int missing_to_zero)
{
    int rc = check_host_query(e, q, r); if (rc) return rc;
    if (q->Q <= 0) return SCL_OK;
    Lane& ln = e->lanes[0];                          /* behind any pipelined batch of lane 0 in stream order */
    const float* dq = nullptr; const int32_t* di = nullptr;
    rc = stage_host_query(e, ln, q, &dq, &di); if (rc) return rc;
    rc = query_host_enqueue(e, ln, dq, di, q, r, missing_to_zero); if (rc) return rc;
    CK(cudaStreamSynchronize(ln.stream));
    return SCL_OK;
}

} // namespace

extern "C" {

void scl_default_params(scl_params* p)
{
    p->num_ring = 20; p->num_sector = 60; p->num_candidates = 3; p->dist_thres = 0.14; p->lidar_height = 1.65;
    p->max_radius = 80.0; p->num_exclude_recent = 100; p->tree_making_period = 10; p->search_ratio = 0.1;
}

void scl_default_icp_params(scl_icp_params* p)
{
    p->max_corr_dist = 100.0; p->max_iterations = 50; p->trans_eps = 1e-6; p->fitness_eps = 1e-6;
}

int scl_create(const scl_params* p, int device, scl_engine** out)
{
    if (!p || !out) return SCL_ERR_INVALID;
    *out = nullptr;
    if (p->num_ring < 1 || p->num_sector < 1 || p->num_candidates < 1 || p->num_candidates > 32 || p->tree_making_period < 1 ||
        !(p->max_radius > 0))
        return SCL_ERR_INVALID;
    /* what EVERY kernel of the path can run, so that a build can never fail later for a geometry that was accepted here:
     * the kNN kernels are built for 10 / 20 / 40 rings, K1's threshold tables hold 4 x 31 sector boundaries, K4's masks 128 sectors */
    if (p->num_ring != 10 && p->num_ring != 20 && p->num_ring != 40) return SCL_ERR_UNSUPPORTED;
    if (p->num_sector > 124 || p->num_ring * p->num_sector > 8192) return SCL_ERR_UNSUPPORTED;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return SCL_ERR_CUDA; /* no CPU fallback */
    if (cudaSetDevice(device) != cudaSuccess) return SCL_ERR_CUDA;
    scl_engine* e = new scl_engine();
    e->p = *p; e->device = device;
    e->search_radius = (int)std::round(0.5 * p->search_ratio * p->num_sector);
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { delete e; return SCL_ERR_CUDA; }
    if (cudaMalloc(&e->d_kn2max, 64) != cudaSuccess || cudaMemset(e->d_kn2max, 0, 64) != cudaSuccess) { delete e; return SCL_ERR_CUDA; }
    {
        /* once per device: load the query path's kernels now, so that no launch on it ever has to synchronise the context */
        static SclOncePerDevice once;
        if (once.first()) { scl_preload_k1(); scl_preload_k3(); scl_preload_k3_tc(); scl_preload_k4(); scl_preload_k7(); (void)cudaGetLastError(); }
    }
    *out = e;
    return SCL_OK;
}

int scl_destroy(scl_engine* e)
{
    if (!e) return SCL_ERR_INVALID;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        cudaSetDevice(e->device);
        cudaStreamSynchronize(e->stream);
        for (int l = 0; l < scl_engine::kLanes; l++) {
            Lane& ln = e->lanes[l];
            if (l > 0 && ln.stream) { cudaStreamSynchronize(ln.stream); cudaStreamDestroy(ln.stream); }
            if (ln.done) cudaEventDestroy(ln.done);
            for (DevBuf* b : ln.all) b->release();
        }
        if (e->db_ready) cudaEventDestroy(e->db_ready);
        for (int r = 0; r < 16; r++) if (e->xchg_peer_map[r]) cudaIpcCloseMemHandle(e->xchg_peer_map[r]);
        if (e->xchg_buf) cudaFree(e->xchg_buf);
        cudaFree(e->d_desc); cudaFree(e->d_keys); cudaFree(e->d_knorm); cudaFree(e->d_cstat); cudaFree(e->d_kn2max); cudaFree(e->d_kimg);
        if (e->r_keys) cudaFree(e->r_keys);
        if (e->r_knorm) cudaFree(e->r_knorm);
        if (e->r_kimg) cudaFree(e->r_kimg);
        DevBuf* bufs[] = {&e->pts, &e->offsets, &e->gbins, &e->tickets, &e->stage_desc, &e->stage_keys, &e->stage_knorm,
                          &e->bins_ring, &e->bins_sector, &e->kf_arena, &e->icp_src, &e->icp_tgt, &e->icp_raw, &e->icp_acc, &e->icp_nn,
                          &e->icp_grid[0][0], &e->icp_grid[0][1], &e->icp_grid[0][2], &e->icp_grid[0][3], &e->icp_grid[0][4],
                          &e->icp_grid[1][0], &e->icp_grid[1][1], &e->icp_grid[1][2], &e->icp_grid[1][3], &e->icp_grid[1][4]};
        for (DevBuf* b : bufs) b->release();
        DevBuf* cloud_bufs[] = {&e->vg_in, &e->vg_world, &e->vg_out, &e->vg_keys[0], &e->vg_keys[1], &e->vg_vals[0], &e->vg_vals[1], &e->vg_head,
                                &e->vg_ord, &e->vg_temp, &e->vg_misc, &e->vg_T, &e->vg_off};
        for (DevBuf* b : cloud_bufs) b->release();
        for (auto& v : e->ev) for (auto& pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
        for (cudaEvent_t x : e->ev_pool) cudaEventDestroy(x);
        if (e->own_stream) cudaStreamDestroy(e->stream);
    }
    delete e;
    return SCL_OK;
}

const char* scl_last_error(scl_engine* e) { return e ? e->err.c_str() : "null engine"; }

int scl_set_stream(scl_engine* e, void* s)
{
    LOCK();
    CK(cudaStreamSynchronize(e->stream));
    if (e->own_stream) cudaStreamDestroy(e->stream);
    e->stream = static_cast<cudaStream_t>(s); e->own_stream = false;
    e->lanes[0].stream = e->stream;
    e->db_dirty = true;
    return SCL_OK;
}

int scl_set_knn_mode(scl_engine* e, int mode, int count_fallbacks)
{
    LOCK();
    if (mode < 0 || mode > 2) FAIL(SCL_ERR_INVALID, "mode must be 0 (auto), 1 (exact) or 2 (tensor core)");
    if (mode == 2 && !scl_knn_tc_supported(e->p.num_ring)) FAIL(SCL_ERR_UNSUPPORTED, "tensor-core kNN is built for 20 and 40 rings");
    e->knn_mode = mode; e->count_fallbacks = count_fallbacks != 0;
    return SCL_OK;
}

int scl_set_tc_stages(scl_engine* e, int stages)
{
    LOCK();
    if (stages < 2 || stages > 5) FAIL(SCL_ERR_INVALID, "2..5 key tiles in flight");
    e->tc_stages = stages;
    return SCL_OK;
}

int scl_set_scdist_mode(scl_engine* e, int mode)
{
    LOCK();
    if (mode != 0 && mode != 1) FAIL(SCL_ERR_INVALID, "mode must be 0 (FP32 prefilter + exact evaluation of the shifts that can win) or 1 (every shift exactly)");
    e->scdist_exact_all = mode == 1;
    return SCL_OK;
}

int scl_set_scdist_tiles(scl_engine* e, int tiles)
{
    LOCK();
    if (tiles != 0 && (tiles < 2 || tiles > 16)) FAIL(SCL_ERR_INVALID, "0 (one tile per candidate) or 2..16 candidate tiles per CTA");
    e->scdist_tiles = tiles;
    return SCL_OK;
}

int scl_knn_stats(scl_engine* e, long long* tc_queries, long long* fallback_queries)
{
    LOCK();
    if (tc_queries) *tc_queries = e->stat_tc_queries;
    if (fallback_queries) {
        /* counted on the device by the re-rank kernel (ints [8] and [24] of each lane's counter block, one per call parity) */
        long long total = 0;
        for (int l = 0; l < scl_engine::kLanes; l++) {
            Lane& ln = e->lanes[l];
            if (!ln.tc_fail_count.p) continue;
            int h[32] = {0};
            CK(cudaStreamSynchronize(ln.stream));
            CK(cudaMemcpy(h, ln.tc_fail_count.p, sizeof(h), cudaMemcpyDeviceToHost));
            total += (long long)h[8] + (long long)h[24];
        }
        *fallback_queries = total;
    }
    return SCL_OK;
}

int scl_set_profiling(scl_engine* e, int on) { LOCK(); e->profiling = on != 0; return SCL_OK; }

int scl_stage_time(scl_engine* e, int stage, double* ms, int* launches)
{
    LOCK();
    if (stage < 0 || stage > 3 || !ms || !launches) FAIL(SCL_ERR_INVALID, "bad stage");
    CK(cudaStreamSynchronize(e->stream));
    for (int l = 1; l < scl_engine::kLanes; l++) if (e->lanes[l].stream) CK(cudaStreamSynchronize(e->lanes[l].stream));
    double total = 0.0; int n = 0;
    for (auto& pr : e->ev[stage]) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) { total += t; n++; }
        e->ev_pool.push_back(pr.first); e->ev_pool.push_back(pr.second);
    }
    e->ev[stage].clear();
    *ms = total; *launches = n;
    return SCL_OK;
}

int scl_reserve(scl_engine* e, int capacity) { LOCK(); return grow(e, capacity); }

int scl_set_shard(scl_engine* e, int rank, int world)
{
    LOCK();
    if (world < 1 || rank < 0 || rank >= world) FAIL(SCL_ERR_INVALID, "bad shard");
    e->rank = rank; e->world = world;
    return SCL_OK;
}

int scl_export_keys_dev(scl_engine* e, float* keys_out_dev, int n)
{
    LOCK();
    if (!keys_out_dev || n < 0 || n > e->n) FAIL(SCL_ERR_RANGE, "n must be within the engine's size");
    if (n) CK(cudaMemcpyAsync(keys_out_dev, e->d_keys, (size_t)n * e->p.num_ring * 4, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SCL_OK;
}

int scl_set_replicated_keys_dev(scl_engine* e, const float* keys_dev, int n_total)
{
    LOCK();
    const int R = e->p.num_ring;
    if (n_total < 0 || (n_total > 0 && !keys_dev)) FAIL(SCL_ERR_INVALID, "keys_dev is NULL or n_total < 0");
    if (n_total > 0 && !scl_knn_tc_supported(R) && R != 10) FAIL(SCL_ERR_UNSUPPORTED, "key length");
    CK(cudaStreamSynchronize(e->stream));
    for (int l = 1; l < scl_engine::kLanes; l++) if (e->lanes[l].stream) CK(cudaStreamSynchronize(e->lanes[l].stream));
    if (e->r_keys) cudaFree(e->r_keys);
    if (e->r_knorm) cudaFree(e->r_knorm);
    if (e->r_kimg) cudaFree(e->r_kimg);
    e->r_keys = e->r_knorm = nullptr; e->r_kimg = nullptr; e->r_n = 0;
    if (n_total == 0) return SCL_OK;
    CK(cudaMalloc(&e->r_keys, (size_t)n_total * R * 4));
    CK(cudaMalloc(&e->r_knorm, (size_t)n_total * 4));
    CK(cudaMemcpyAsync(e->r_keys, keys_dev, (size_t)n_total * R * 4, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemsetAsync(e->d_kn2max + 1, 0, 4, e->stream));
    CK(scl_launch_key_norms(e->r_keys, 0, n_total, R, e->r_knorm, e->d_kn2max + 1, e->stream));
    if (scl_knn_tc_supported(R)) {
        const size_t bytes = scl_knn_tc_image_bytes(R, n_total);
        CK(cudaMalloc(&e->r_kimg, bytes));
        CK(cudaMemsetAsync(e->r_kimg, 0, bytes, e->stream));
        CK(scl_launch_key_image(e->r_keys, e->r_knorm, 0, n_total, R, e->r_kimg, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    e->r_n = n_total;
    return SCL_OK;
}

int scl_polar_tables(const scl_params* p, float* ring_thr, int32_t* n_ring, float* s_max, float* sec_thr, int32_t* n_sec, int32_t* sec_base, int32_t* sec_dir)
{
    if (!p) return SCL_ERR_INVALID;
    const int rc = scl_polar_tables_host(p->num_ring, p->num_sector, p->max_radius, ring_thr, n_ring, s_max, sec_thr, n_sec, sec_base, sec_dir);
    return rc == 0 ? SCL_OK : (rc == 1 ? SCL_ERR_INVALID : SCL_ERR_UNSUPPORTED);
}

int scl_build_insert(scl_engine* e, const void* pts, int n, int stride_bytes, int8_t robot, int index, float* out_desc)
{
    LOCK();
    if (n < 0) FAIL(SCL_ERR_INVALID, "n < 0");
    const int32_t off[2] = {0, n};
    return build_host(e, pts, off, 1, stride_bytes, 1, &robot, &index, out_desc, nullptr, nullptr);
}

int scl_make_scancontext(scl_engine* e, const void* pts, int n, int stride_bytes, float* out_desc, int32_t* out_ring, int32_t* out_sector)
{
    LOCK();
    if (n < 0) FAIL(SCL_ERR_INVALID, "n < 0");
    const int32_t off[2] = {0, n};
    return build_host(e, pts, off, 1, stride_bytes, 0, nullptr, nullptr, out_desc, out_ring, out_sector);
}

int scl_build_batch(scl_engine* e, const void* pts, const int32_t* offsets, int n_scans, int stride_bytes, int insert,
                    const int8_t* robots, const int32_t* indices, float* out_desc)
{
    LOCK();
    return build_host(e, pts, offsets, n_scans, stride_bytes, insert, robots, indices, out_desc, nullptr, nullptr);
}

int scl_build_batch_dev(scl_engine* e, const void* pts_dev, const int32_t* offsets, int n_scans, int stride_bytes, int insert,
                        const int8_t* robots, const int32_t* indices, float* out_desc_dev)
{
    LOCK();
    if (n_scans < 0 || stride_bytes < 12 || (stride_bytes & 3)) FAIL(SCL_ERR_INVALID, "bad arguments");
    return build_dev(e, pts_dev, offsets, n_scans, stride_bytes, insert, robots, indices, out_desc_dev, nullptr, nullptr, nullptr);
}

int scl_insert_batch(scl_engine* e, const float* descs, int n, const int8_t* robots, const int32_t* indices)
{
    LOCK();
    if (n < 0 || (n > 0 && !descs)) FAIL(SCL_ERR_INVALID, "bad arguments");
    if (n == 0) return SCL_OK;
    int rc = grow(e, e->n + n); if (rc) return rc;
    const size_t RS = e->RS();
    float* dst = e->d_desc + (size_t)e->n * RS;
    CK(cudaMemcpyAsync(dst, descs, (size_t)n * RS * 4, cudaMemcpyHostToDevice, e->stream));   /* wire decode, descriptor.h:1575-1582 */
    CK(scl_launch_ring_keys(dst, n, e->p.num_ring, e->p.num_sector, e->d_keys + (size_t)e->n * e->p.num_ring, e->d_knorm + e->n, e->d_kn2max,
                            e->d_cstat + (size_t)e->n * 2 * e->p.num_sector, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    append_index(e, n, robots, indices); e->n += n;
    e->db_dirty = true;
    return SCL_OK;
}

int scl_insert(scl_engine* e, const float* desc, int8_t robot, int index) { return scl_insert_batch(e, desc, 1, &robot, &index); }

int scl_insert_batch_dev(scl_engine* e, const float* descs_dev, int n, const int8_t* robots, const int32_t* indices)
{
    LOCK();
    if (n < 0 || (n > 0 && !descs_dev)) FAIL(SCL_ERR_INVALID, "bad arguments");
    if (n == 0) return SCL_OK;
    int rc = grow(e, e->n + n); if (rc) return rc;
    const size_t RS = e->RS();
    float* dst = e->d_desc + (size_t)e->n * RS;
    CK(cudaMemcpyAsync(dst, descs_dev, (size_t)n * RS * 4, cudaMemcpyDeviceToDevice, e->stream));
    CK(scl_launch_ring_keys(dst, n, e->p.num_ring, e->p.num_sector, e->d_keys + (size_t)e->n * e->p.num_ring, e->d_knorm + e->n, e->d_kn2max,
                            e->d_cstat + (size_t)e->n * 2 * e->p.num_sector, e->stream));
    append_index(e, n, robots, indices); e->n += n;
    e->db_dirty = true;
    return SCL_OK;
}

int scl_get_index(scl_engine* e, int key, int8_t* robot, int* index)
{
    LOCK();
    if (!robot || !index) FAIL(SCL_ERR_INVALID, "null output");
    if (key < 0 || key >= e->n) { *robot = -1; *index = -1; return SCL_OK; }   /* the reference's getIndex(-1) is UB */
    *robot = e->index[key].first; *index = e->index[key].second;
    return SCL_OK;
}

int scl_size(scl_engine* e) { if (!e) return -1; std::lock_guard<std::mutex> lk(e->mu); return e->n; }

int scl_get_descriptor(scl_engine* e, int key, float* out)
{
    LOCK();
    if (key < 0 || key >= e->n || !out) FAIL(SCL_ERR_RANGE, "key out of range");
    CK(cudaMemcpyAsync(out, e->d_desc + (size_t)key * e->RS(), (size_t)e->RS() * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SCL_OK;
}

int scl_get_ring_key(scl_engine* e, int key, float* out)
{
    LOCK();
    if (key < 0 || key >= e->n || !out) FAIL(SCL_ERR_RANGE, "key out of range");
    CK(cudaMemcpyAsync(out, e->d_keys + (size_t)key * e->p.num_ring, (size_t)e->p.num_ring * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SCL_OK;
}

int scl_query_batch(scl_engine* e, const scl_batch_query* q, scl_batch_result* r) { LOCK(); return query_host(e, q, r, 0); }

int scl_query_batch_dev(scl_engine* e, const scl_batch_query* q, scl_batch_result* r)
{
    LOCK();
    if (!q || !r) FAIL(SCL_ERR_INVALID, "null query/result");
    return query_dev(e, e->lanes[0], q->q_desc, q->q_ids, q->Q, q->K, q->n_db, q->metric, 0, r->cand_ids, r->cand_d2, r->cand_dist, r->cand_shift,
                     r->best_id, r->best_dist, r->best_shift);
}

int scl_num_lanes(void) { return scl_engine::kLanes; }

int scl_query_batch_dev_lane(scl_engine* e, int lane, const scl_batch_query* q, scl_batch_result* r)
{
    LOCK();
    if (!q || !r) FAIL(SCL_ERR_INVALID, "null query/result");
    if (lane < 0 || lane >= scl_engine::kLanes) FAIL(SCL_ERR_INVALID, "no such lane");
    return query_dev(e, e->lanes[lane], q->q_desc, q->q_ids, q->Q, q->K, q->n_db, q->metric, 0, r->cand_ids, r->cand_d2, r->cand_dist, r->cand_shift,
                     r->best_id, r->best_dist, r->best_shift);
}

int scl_lanes_fork(scl_engine* e, void* stream)
{
    LOCK();
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CK(cudaEventRecord(ev, static_cast<cudaStream_t>(stream)));
    for (int l = 0; l < scl_engine::kLanes; l++) {
        Lane& ln = e->lanes[l];
        int rc = lane_begin(e, ln); if (rc) { cudaEventDestroy(ev); return rc; }
        if (ln.stream != static_cast<cudaStream_t>(stream)) CK(cudaStreamWaitEvent(ln.stream, ev, 0));
    }
    CK(cudaEventDestroy(ev));                        /* released once the waits that captured it have passed */
    return SCL_OK;
}

int scl_lanes_join(scl_engine* e, void* stream)
{
    LOCK();
    for (int l = 0; l < scl_engine::kLanes; l++) {
        Lane& ln = e->lanes[l];
        if (!ln.stream || ln.stream == static_cast<cudaStream_t>(stream)) continue;
        cudaEvent_t ev;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CK(cudaEventRecord(ev, ln.stream));
        CK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), ev, 0));
        CK(cudaEventDestroy(ev));
    }
    return SCL_OK;
}

int scl_lane_sync(scl_engine* e, int lane)
{
    if (!e) return SCL_ERR_INVALID;
    cudaStream_t s;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        if (lane < 0 || lane >= scl_engine::kLanes) FAIL(SCL_ERR_INVALID, "no such lane");
        s = e->lanes[lane].stream;
        if (!s && lane == 0) s = e->stream;          /* lane 0 is the engine's own stream, also before its first batch */
    }
    cudaSetDevice(e->device);
    if (s && cudaStreamSynchronize(s) != cudaSuccess) { std::lock_guard<std::mutex> lk(e->mu); FAIL(SCL_ERR_CUDA, "cudaStreamSynchronize failed"); }
    return SCL_OK;
}

namespace {
// ticket -> lane; at most kLanes batches in flight
int pipe_take_lane(scl_engine* e, Lane** out, int* ticket)
{
    const int l = (int)(e->pipe_next % scl_engine::kLanes);
    Lane& ln = e->lanes[l];
    if (ln.busy) FAIL(SCL_ERR_INVALID, "every lane has a batch in flight: wait for the oldest one first");
    if (!ln.done) CK(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
    *out = &ln;
    *ticket = (int)(e->pipe_next & 0x7fffffff);
    return SCL_OK;
}
} // namespace

int scl_query_batch_submit(scl_engine* e, const scl_batch_query* q, scl_batch_result* r, int* ticket)
{
    LOCK();
    if (!ticket) FAIL(SCL_ERR_INVALID, "null ticket");
    int rc = check_host_query(e, q, r); if (rc) return rc;
    Lane* lnp = nullptr; int t = 0;
    rc = pipe_take_lane(e, &lnp, &t); if (rc) return rc;
    Lane& ln = *lnp;
    if (q->Q > 0) {
        const float* dq = nullptr; const int32_t* di = nullptr;
        rc = stage_host_query(e, ln, q, &dq, &di); if (rc) return rc;
        rc = query_host_enqueue(e, ln, dq, di, q, r, 0); if (rc) return rc;
    } else { rc = lane_begin(e, ln); if (rc) return rc; }
    CK(cudaEventRecord(ln.done, ln.stream));
    ln.busy = true;
    *ticket = t;
    e->pipe_next++;
    return SCL_OK;
}

int scl_query_batch_wait(scl_engine* e, int ticket)
{
    if (!e) return SCL_ERR_INVALID;
    cudaEvent_t ev;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        const int l = ticket % scl_engine::kLanes;
        if (ticket < 0 || !e->lanes[l].busy) FAIL(SCL_ERR_INVALID, "no batch in flight for this ticket");
        ev = e->lanes[l].done;
    }
    cudaSetDevice(e->device);
    cudaError_t err = cudaEventSynchronize(ev);              /* outside the lock: inserts and submits may proceed meanwhile */
    std::lock_guard<std::mutex> lk(e->mu);
    e->lanes[ticket % scl_engine::kLanes].busy = false;
    if (err != cudaSuccess) { e->err = std::string("cudaEventSynchronize: ") + cudaGetErrorString(err); return SCL_ERR_CUDA; }
    return SCL_OK;
}

int scl_merge_shards_dev(scl_engine* e, int world, int Q, int K, const int32_t* q_ids, const int32_t* all_ids, const float* all_d2,
                         const double* all_dist, const int32_t* all_shift, scl_batch_result* m)
{
    LOCK();
    if (!m) FAIL(SCL_ERR_INVALID, "null result");
    CK(scl_launch_merge_shards(world, Q, K, q_ids, all_ids, all_d2, all_dist, all_shift, m->cand_ids, m->cand_d2, m->cand_dist,
                               m->cand_shift, m->best_id, m->best_dist, m->best_shift, e->stream));
    return SCL_OK;
}

int scl_knn_batch_dev(scl_engine* e, const scl_batch_query* q, int32_t* ids_dev, float* d2_dev)
{
    LOCK();
    if (!q || !ids_dev || !d2_dev) FAIL(SCL_ERR_INVALID, "null argument");
    return knn_dev(e, e->lanes[0], q->q_desc, q->q_ids, q->Q, q->K, q->n_db, q->metric, ids_dev, d2_dev, nullptr);
}

int scl_merge_topk_dev(scl_engine* e, int world, int Q, int K, const void* ids_base, const void* d2_base, uint64_t rank_stride_bytes,
                       int32_t* out_ids, float* out_d2)
{
    LOCK();
    if (!ids_base || !d2_base || !out_ids || !out_d2) FAIL(SCL_ERR_INVALID, "null argument");
    CK(scl_launch_merge_topk(world, Q, K, ids_base, d2_base, (size_t)rank_stride_bytes, out_ids, out_d2, e->stream));
    return SCL_OK;
}

namespace {
int scdist_owned(scl_engine* e, Lane& ln, const float* q_desc_dev, const int32_t* q_ids_dev, int Q, int K, const int32_t* cand_ids_dev,
                 double* dist_dev, int32_t* shift_dev, bool stats_ready /* knn_dev has just run on this lane for these very queries */)
{
    /* the global candidate lists are spread over the shards: about K / world of a query's candidates live here */
    ln.scdist_owned_hint = e->world > 1 ? (K + e->world - 1) / e->world + (e->world <= 2 ? 1 : 0) : 0;
    const int rc = scdist_dev(e, ln, q_desc_dev, nullptr, q_ids_dev, Q, K, const_cast<int32_t*>(cand_ids_dev), 0, dist_dev, shift_dev, nullptr, nullptr, nullptr,
                              stats_ready && ln.qstat_of == q_desc_dev && ln.qstat_rows >= Q);
    ln.scdist_owned_hint = 0;
    return rc;
}
} // namespace

int scl_scdist_owned_dev(scl_engine* e, const float* q_desc_dev, const int32_t* q_ids_dev, int Q, int K, const int32_t* cand_ids_dev,
                         double* dist_dev, int32_t* shift_dev)
{
    LOCK();
    if (!q_desc_dev || !cand_ids_dev || !dist_dev || !shift_dev) FAIL(SCL_ERR_INVALID, "null argument");
    if (K < 1 || K > 32) FAIL(SCL_ERR_INVALID, "K must be in 1..32");
    return scdist_owned(e, e->lanes[0], q_desc_dev, q_ids_dev, Q, K, cand_ids_dev, dist_dev, shift_dev, false);
}

int scl_combine_owned_dev(scl_engine* e, int world, int Q, int K, const int32_t* q_ids, const int32_t* cand_ids, const void* dist_base,
                          const void* shift_base, uint64_t rank_stride_bytes, scl_batch_result* m)
{
    LOCK();
    if (!cand_ids || !dist_base || !shift_base || !m) FAIL(SCL_ERR_INVALID, "null argument");
    CK(scl_launch_combine_owned(world, Q, K, q_ids, cand_ids, dist_base, shift_base, (size_t)rank_stride_bytes, m->cand_dist, m->cand_shift,
                                m->best_id, m->best_dist, m->best_shift, e->stream));
    return SCL_OK;
}

// ---- peer-memory exchange (k7_exchange.cu) ------------------------------------------------------------------------
namespace {
size_t xchg_layout(scl_engine* e, int world, int max_q, int max_k)
{
    /* one region per lane: [flags 3 x 16 ints | tickets | point 0 slots | point 1 slots | query areas] */
    const size_t qk = (size_t)max_q * max_k;
    const size_t s0 = (qk * 8 + 15) / 16 * 16, s1 = (qk * 12 + 15) / 16 * 16;
    const size_t s2 = ((size_t)max_q * e->RS() * 4 + 15) / 16 * 16;
    size_t o = 0;
    for (int l = 0; l < scl_engine::kLanes; l++) {
        XchgView& x = e->xchg[l];
        x.world = world; x.rank = -1;
        x.flag_off = o; o += 256;
        x.ticket_off = o; o += 256;
        x.slot_bytes[0] = s0; x.data_off[0] = o; o += 2 * (size_t)world * s0;
        x.slot_bytes[1] = s1; x.data_off[1] = o; o += 2 * (size_t)world * s1;
        x.slot_bytes[2] = s2; x.data_off[2] = o; o += 2 * s2;
    }
    return o;
}
} // namespace

int scl_xchg_bytes(scl_engine* e, int world, int max_q, int max_k, uint64_t* bytes)
{
    LOCK();
    if (world < 2 || world > 16 || max_q < 1 || max_k < 1 || !bytes) FAIL(SCL_ERR_INVALID, "bad arguments");
    XchgView keep[scl_engine::kLanes];
    memcpy(keep, e->xchg, sizeof(keep));
    *bytes = xchg_layout(e, world, max_q, max_k);
    memcpy(e->xchg, keep, sizeof(keep));
    return SCL_OK;
}

namespace {
// Every buffer a sharded step of up to max_q queries and max_k candidates touches on a lane, allocated NOW: a step must
// never allocate (cudaMalloc / cudaFree may synchronise the device) between launching an exchange kernel, which waits for
// a peer, and giving that peer its work — with one host thread driving several devices that would be a deadlock.
int prealloc_lane(scl_engine* e, Lane& ln, int max_q, int max_k)
{
    const size_t Q = (size_t)max_q, QK = Q * max_k, R = e->p.num_ring, S = e->p.num_sector;
    int rc = lane_begin(e, ln); if (rc) return rc;
    CK(ln.qkeys.ensure(Q * R * 4)); CK(ln.qknorm.ensure(Q * 4)); CK(ln.qstat.ensure(Q * 2 * S * 8)); CK(ln.qlocal.ensure(Q * 4)); CK(ln.qids.ensure(Q * 4));
    CK(ln.cand_local.ensure(QK * 4)); CK(ln.cand_ids.ensure(QK * 4)); CK(ln.cand_d2.ensure(QK * 4)); CK(ln.cand_dist.ensure(QK * 8)); CK(ln.cand_shift.ensure(QK * 4));
    CK(ln.best_id.ensure(Q * 4)); CK(ln.best_dist.ensure(Q * 8)); CK(ln.best_shift.ensure(Q * 4));
    CK(ln.part_ids.ensure(Q * 256 * max_k * 4)); CK(ln.part_d2.ensure(Q * 256 * max_k * 4));           /* up to 256 key splits in the exact kernel */
    CK(ln.knn_tickets.ensure((Q / 128 + 16) * 4 + Q * 4));
    CK(cudaMemsetAsync(ln.knn_tickets.p, 0, ln.knn_tickets.cap, ln.stream));
    const int Qc = max_q < scl_knn_tc_max_batch() ? max_q : scl_knn_tc_max_batch();
    const size_t pairs = (size_t)Qc * scl_knn_tc_ranges(Qc);
    /* the tensor-core kNN runs from four queries up; one query group uses all 148 ranges */
    const size_t pairs_max = std::max(pairs, (size_t)std::min(Qc, 256) * SCL_NUM_SMS);
    CK(ln.tc_queues.ensure(pairs_max * scl_knn_tc_queue_bytes())); CK(ln.tc_queue_cnt.ensure(pairs_max * 4));
    CK(ln.tc_fail_list.ensure(Q * 4)); CK(ln.tc_fail_count.ensure(128)); CK(ln.tc_slots.ensure((size_t)Qc * scl_knn_tc_slot_stride() * 4));
    CK(cudaMemsetAsync(ln.tc_fail_count.p, 0, 128, ln.stream));
    ln.tc_state_clean = false;
    CK(ln.x_blob1.ensure(QK * 8)); CK(ln.x_blob2.ensure(QK * 12)); CK(ln.x_ids.ensure(QK * 4)); CK(ln.x_d2.ensure(QK * 4));
    if (!ln.done) CK(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
    CK(cudaStreamSynchronize(ln.stream));
    return SCL_OK;
}
} // namespace

int scl_xchg_create(scl_engine* e, int world, int max_q, int max_k, unsigned char* handle64)
{
    LOCK();
    if (world < 2 || world > 16 || max_q < 1 || max_k < 1 || max_k > 32 || !handle64) FAIL(SCL_ERR_INVALID, "bad arguments");
    if (e->xchg_buf) FAIL(SCL_ERR_INVALID, "exchange buffer exists already");
    for (int l = 0; l < scl_engine::kLanes; l++) { int rc = prealloc_lane(e, e->lanes[l], max_q, max_k); if (rc) return rc; }
    e->xchg_bytes = xchg_layout(e, world, max_q, max_k);
    e->xchg_qk = max_q * max_k; e->xchg_q = max_q;
    CK(cudaMalloc(&e->xchg_buf, e->xchg_bytes));
    CK(cudaMemset(e->xchg_buf, 0, e->xchg_bytes));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, e->xchg_buf));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return SCL_OK;
}

int scl_xchg_open(scl_engine* e, int world, int rank, const unsigned char* handles)
{
    LOCK();
    if (!e->xchg_buf || world != e->xchg[0].world || rank < 0 || rank >= world || !handles) FAIL(SCL_ERR_INVALID, "bad arguments");
    unsigned char* peer[16] = {};
    for (int r = 0; r < world; r++) {
        if (r == rank) { peer[r] = static_cast<unsigned char*>(e->xchg_buf); continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        e->xchg_peer_map[r] = p;
        peer[r] = static_cast<unsigned char*>(p);
    }
    for (int l = 0; l < scl_engine::kLanes; l++) { memcpy(e->xchg[l].peer, peer, sizeof(peer)); e->xchg[l].rank = rank; }
    e->xchg_open = true;
    return SCL_OK;
}

/* the same within ONE process (scl_sharded.cu: one engine per device of a box): the peers' buffers are plain device
 * pointers once peer access is enabled */
int scl_xchg_open_local(scl_engine* e, int world, int rank, void* const* peer_bufs)
{
    LOCK();
    if (!e->xchg_buf || world != e->xchg[0].world || rank < 0 || rank >= world || !peer_bufs) FAIL(SCL_ERR_INVALID, "bad arguments");
    for (int l = 0; l < scl_engine::kLanes; l++) {
        for (int r = 0; r < world; r++) e->xchg[l].peer[r] = static_cast<unsigned char*>(r == rank ? e->xchg_buf : peer_bufs[r]);
        e->xchg[l].rank = rank;
    }
    e->xchg_open = true;
    return SCL_OK;
}

void* scl_xchg_buffer(scl_engine* e) { return e ? e->xchg_buf : nullptr; }

int scl_xchg_close(scl_engine* e)
{
    LOCK();
    cudaStreamSynchronize(e->stream);
    for (int l = 1; l < scl_engine::kLanes; l++) if (e->lanes[l].stream) cudaStreamSynchronize(e->lanes[l].stream);
    for (int r = 0; r < 16; r++) if (e->xchg_peer_map[r]) { cudaIpcCloseMemHandle(e->xchg_peer_map[r]); e->xchg_peer_map[r] = nullptr; }
    if (e->xchg_buf) { cudaFree(e->xchg_buf); e->xchg_buf = nullptr; }
    e->xchg_open = false;
    return SCL_OK;
}

int scl_xchg_merge_topk_dev(scl_engine* e, int seq, int Q, int K, const void* my_block_dev, int32_t* out_ids, float* out_d2)
{
    LOCK();
    if (!e->xchg_open) FAIL(SCL_ERR_INVALID, "exchange not open");
    if (seq < 1 || Q < 0 || K < 1 || (long long)Q * K > e->xchg_qk || ((long long)Q * K) % 4 || !my_block_dev || !out_ids || !out_d2)
        FAIL(SCL_ERR_INVALID, "bad arguments (Q*K must be a multiple of 4 within the size given to scl_xchg_create)");
    CK(scl_launch_xchg_merge_topk(e->xchg[0], seq, Q, K, my_block_dev, out_ids, out_d2, e->stream));
    return SCL_OK;
}

int scl_xchg_combine_dev(scl_engine* e, int seq, int Q, int K, const void* my_block_dev, const int32_t* q_ids, const int32_t* cand_ids,
                         scl_batch_result* m)
{
    LOCK();
    if (!e->xchg_open) FAIL(SCL_ERR_INVALID, "exchange not open");
    if (seq < 1 || Q < 0 || K < 1 || (long long)Q * K > e->xchg_qk || ((long long)Q * K) % 4 || !my_block_dev || !cand_ids || !m)
        FAIL(SCL_ERR_INVALID, "bad arguments (Q*K must be a multiple of 4 within the size given to scl_xchg_create)");
    CK(scl_launch_xchg_combine(e->xchg[0], seq, Q, K, my_block_dev, q_ids, cand_ids, m->cand_dist, m->cand_shift, m->best_id, m->best_dist,
                               m->best_shift, e->stream));
    return SCL_OK;
}

// ---- the sharded query as ONE call per batch and lane -----------------------------------------------------
namespace {
// device side of one sharded step on lane ln: K2 + K3 on the shard, exchange + global top-K, K4 on the owned candidates,
// exchange + winner scan. q_desc: the full batch in device memory. Results: device pointers (scratch where null).
int shard_step(scl_engine* e, Lane& ln, int lane, const float* q_desc, const int32_t* q_ids, int Q, int K, int n_db, int metric, scl_batch_result* r)
{
    if (!e->xchg_open) FAIL(SCL_ERR_INVALID, "exchange not open (scl_xchg_create / scl_xchg_open)");
    if (Q < 1 || K < 1 || K > 32 || (long long)Q * K > e->xchg_qk || ((long long)Q * K) % 4) FAIL(SCL_ERR_INVALID, "Q*K must be a multiple of 4 within the size given to scl_xchg_create");
    const size_t QK = (size_t)Q * K;
    CK(ln.x_blob1.ensure(QK * 8)); CK(ln.x_blob2.ensure(QK * 12)); CK(ln.x_ids.ensure(QK * 4)); CK(ln.x_d2.ensure(QK * 4));
    int32_t* loc_ids = ln.x_blob1.as<int32_t>(); float* loc_d2 = reinterpret_cast<float*>(ln.x_blob1.as<unsigned char>() + QK * 4);
    double* own_dist = ln.x_blob2.as<double>(); int32_t* own_shift = reinterpret_cast<int32_t*>(ln.x_blob2.as<unsigned char>() + QK * 8);
    int32_t* g_ids = r->cand_ids ? r->cand_ids : ln.x_ids.as<int32_t>();
    float* g_d2 = r->cand_d2 ? r->cand_d2 : ln.x_d2.as<float>();
    int rc = knn_dev(e, ln, q_desc, nullptr, Q, K, n_db, metric, loc_ids, loc_d2, nullptr); if (rc) return rc;
    const int seq = ++ln.xseq;
    CK(scl_launch_xchg_merge_topk(e->xchg[lane], seq, Q, K, ln.x_blob1.p, g_ids, g_d2, ln.stream));
    rc = scdist_owned(e, ln, q_desc, q_ids, Q, K, g_ids, own_dist, own_shift, true); if (rc) return rc;
    CK(scl_launch_xchg_combine(e->xchg[lane], seq, Q, K, ln.x_blob2.p, q_ids, g_ids, r->cand_dist, r->cand_shift, r->best_id, r->best_dist,
                               r->best_shift, ln.stream));
    return SCL_OK;
}
} // namespace

int scl_shard_query_dev(scl_engine* e, int lane, const scl_batch_query* q, scl_batch_result* r)
{
    LOCK();
    if (!q || !r || !q->q_desc) FAIL(SCL_ERR_INVALID, "a sharded query needs device descriptors and a result");
    if (lane < 0 || lane >= scl_engine::kLanes) FAIL(SCL_ERR_INVALID, "no such lane");
    return shard_step(e, e->lanes[lane], lane, q->q_desc, q->q_ids, q->Q, q->K, q->n_db, q->metric, r);
}

int scl_shard_query_submit(scl_engine* e, const scl_batch_query* q, scl_batch_result* r, int* ticket)
{
    LOCK();
    if (!q || !r || !ticket || !q->q_desc) FAIL(SCL_ERR_INVALID, "a sharded query needs host descriptors, a result and a ticket");
    if (!e->xchg_open) FAIL(SCL_ERR_INVALID, "exchange not open (scl_xchg_create / scl_xchg_open)");
    const int Q = q->Q, K = q->K, world = e->world, rank = e->rank;
    if (Q < 1 || Q > e->xchg_q) FAIL(SCL_ERR_INVALID, "Q must be within the size given to scl_xchg_create");
    Lane* lnp = nullptr; int t = 0;
    int rc = pipe_take_lane(e, &lnp, &t); if (rc) return rc;
    Lane& ln = *lnp;
    const int lane = (int)(e->pipe_next % scl_engine::kLanes);
    rc = lane_begin(e, ln); if (rc) return rc;
    /* every rank uploads 1 / world of the batch (its rows) straight into its own query area and the gather kernel stores
     * them into every peer's: the host link carries each descriptor once, NVLink the rest */
    const size_t RS4 = (size_t)e->RS() * 4;
    const XchgView& x = e->xchg[lane];
    const int seq_next = ln.xseq + 1;
    unsigned char* area = static_cast<unsigned char*>(e->xchg_buf) + x.data_off[2] + (size_t)(seq_next & 1) * x.slot_bytes[2];
    const int per = (Q + world - 1) / world;
    const int row0 = rank * per < Q ? rank * per : Q, rows = Q - row0 < per ? Q - row0 : per;     /* the last ranks may have nothing to bring */
    if (rows > 0)
        CK(cudaMemcpyAsync(area + (size_t)row0 * RS4, reinterpret_cast<const unsigned char*>(q->q_desc) + (size_t)row0 * RS4, (size_t)rows * RS4,
                           cudaMemcpyHostToDevice, ln.stream));
    const int32_t* di = nullptr;
    if (q->q_ids) {                                   /* the keys of the queries (self-skip rule, descriptor.h:1731): a few bytes, every rank */
        CK(ln.qids.ensure((size_t)Q * 4));
        CK(cudaMemcpyAsync(ln.qids.p, q->q_ids, (size_t)Q * 4, cudaMemcpyHostToDevice, ln.stream));
        di = ln.qids.as<int32_t>();
    }
    /* the gather uses its own step counter space: exchange point 2 of step seq_next */
    CK(scl_launch_xchg_gather_queries(x, seq_next, area + (size_t)row0 * RS4, (size_t)row0 * RS4, (size_t)rows * RS4, ln.stream));
    const size_t QK = (size_t)Q * K;
    CK(ln.cand_dist.ensure(QK * 8)); CK(ln.cand_shift.ensure(QK * 4));
    CK(ln.best_id.ensure((size_t)Q * 4)); CK(ln.best_dist.ensure((size_t)Q * 8)); CK(ln.best_shift.ensure((size_t)Q * 4));
    CK(ln.cand_ids.ensure(QK * 4)); CK(ln.cand_d2.ensure(QK * 4));
    scl_batch_result d{ln.cand_ids.as<int32_t>(), ln.cand_d2.as<float>(), ln.cand_dist.as<double>(), ln.cand_shift.as<int32_t>(),
                       ln.best_id.as<int32_t>(), ln.best_dist.as<double>(), ln.best_shift.as<int32_t>()};
    rc = shard_step(e, ln, lane, reinterpret_cast<const float*>(area), di, Q, K, q->n_db, q->metric, &d); if (rc) return rc;
    if (r->cand_ids) CK(cudaMemcpyAsync(r->cand_ids, d.cand_ids, QK * 4, cudaMemcpyDeviceToHost, ln.stream));
    if (r->cand_d2) CK(cudaMemcpyAsync(r->cand_d2, d.cand_d2, QK * 4, cudaMemcpyDeviceToHost, ln.stream));
    if (r->cand_dist) CK(cudaMemcpyAsync(r->cand_dist, d.cand_dist, QK * 8, cudaMemcpyDeviceToHost, ln.stream));
    if (r->cand_shift) CK(cudaMemcpyAsync(r->cand_shift, d.cand_shift, QK * 4, cudaMemcpyDeviceToHost, ln.stream));
    if (r->best_id) CK(cudaMemcpyAsync(r->best_id, d.best_id, (size_t)Q * 4, cudaMemcpyDeviceToHost, ln.stream));
    if (r->best_dist) CK(cudaMemcpyAsync(r->best_dist, d.best_dist, (size_t)Q * 8, cudaMemcpyDeviceToHost, ln.stream));
    if (r->best_shift) CK(cudaMemcpyAsync(r->best_shift, d.best_shift, (size_t)Q * 4, cudaMemcpyDeviceToHost, ln.stream));
    CK(cudaEventRecord(ln.done, ln.stream));
    ln.busy = true;
    *ticket = t;
    e->pipe_next++;
    return SCL_OK;
}

int scl_query_intra(scl_engine* e, int cur, int* id, float* second)
{
    LOCK();
    if (!id || !second) FAIL(SCL_ERR_INVALID, "null output");
    *id = -1; *second = 0.0f;
    if (cur < 0 || cur >= e->n) FAIL(SCL_ERR_RANGE, "query key out of range");
    const int K = e->p.num_candidates;
    if (cur < e->p.num_exclude_recent + K + 1) return SCL_OK;               /* descriptor.h:1620 */
    int32_t ids[32]; double dist[32]; int32_t shift[32];
    scl_batch_query q{nullptr, &cur, 1, K, cur - e->p.num_exclude_recent, 1};  /* :1627; libnabo flavour */
    scl_batch_result r{ids, nullptr, dist, shift, nullptr, nullptr, nullptr};
    int rc = query_host(e, &q, &r, 0); if (rc) return rc;
    float minDis = 10000000.0f; int minIndex = -1, minBias = 0;             /* :1637-1659: minDis is a float there */
    for (int i = 0; i < K; i++) {
        if (ids[i] < 0) continue;
        if (dist[i] < (double)minDis) { minDis = (float)dist[i]; minIndex = ids[i]; minBias = shift[i]; }
    }
    if ((double)minDis < e->p.dist_thres) { *id = minIndex; *second = (float)minBias; }
    return SCL_OK;
}

int scl_query_inter(scl_engine* e, int cur, int* id, float* second)
{
    LOCK();
    if (!id || !second) FAIL(SCL_ERR_INVALID, "null output");
    *id = -1; *second = 0.0f;
    if (cur < 0 || cur >= e->n) FAIL(SCL_ERR_RANGE, "query key out of range");
    const int K = e->p.num_candidates;
    if (e->n < e->p.num_exclude_recent + 1) return SCL_OK;                  /* descriptor.h:1684 */
    if (e->tree_counter % e->p.tree_making_period == 0) e->n_tree = e->n - e->p.num_exclude_recent;   /* :1691-1699 */
    e->tree_counter++;
    int32_t ids[32]; double dist[32]; int32_t shift[32];
    scl_batch_query q{nullptr, &cur, 1, K, e->n_tree, 0};
    scl_batch_result r{ids, nullptr, dist, shift, nullptr, nullptr, nullptr};
    int rc = query_host(e, &q, &r, 1 /* unfilled result slots read entry 0, :1710,1723 */); if (rc) return rc;
    double min_dist = 10000000; int nn_align = 0, nn_idx = -1;
    for (int i = 0; i < K; i++)
        if (dist[i] < min_dist && ids[i] != cur) { min_dist = dist[i]; nn_align = shift[i]; nn_idx = ids[i]; }
    if (min_dist < e->p.dist_thres) *id = nn_idx;
    const double unit = 360.0 / double(e->p.num_sector);
    *second = (float)(nn_align * unit * M_PI / 180.0);                      /* :1752 */
    return SCL_OK;
}

// ---- cloud preparation (K6) -------------------------------------------------------------------------------------
namespace {
inline float ordered_to_float(int i) { const int b = i ^ ((i >> 31) & 0x7fffffff); float f; memcpy(&f, &b, 4); return f; }

// pcl::VoxelGrid on a device cloud (16-byte aligned records); result: packed float4 in e->vg_out, count in *n_out (host)
int voxel_grid_dev(scl_engine* e, const void* d_pts, int n, int stride, float leaf, int* n_out)
{
    *n_out = 0;
    if (n <= 0) return SCL_OK;
    if (!(leaf > 0.0f)) FAIL(SCL_ERR_INVALID, "leaf size must be positive");
    CK(e->vg_misc.ensure(64));
    int* d_bounds = e->vg_misc.as<int>();
    int* d_nout = d_bounds + 8;
    CK(scl_launch_cloud_bounds(d_pts, n, stride, d_bounds, e->stream));
    int hb[6];
    CK(cudaMemcpyAsync(hb, d_bounds, sizeof(hb), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(e->vg_out.ensure((size_t)n * 16));
    if (hb[0] == 0x7fffffff) return SCL_OK;                          /* no finite point */
    float mn[3], mx[3];
    for (int k = 0; k < 3; k++) { mn[k] = ordered_to_float(hb[k]); mx[k] = ordered_to_float(hb[3 + k]); }
    const float inv = 1.0f / leaf;                                   /* voxel_grid.h: inverse_leaf_size_ = 1 / leaf_size_ */
    /* voxel_grid.hpp: refuse grids whose linear index would overflow an int (PCL then returns the input unchanged) */
    const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1, dz = (long long)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > 2147483647LL) {
        CK(scl_launch_pack_xyzi(d_pts, n, stride, e->vg_out.p, e->stream));
        *n_out = n;
        return SCL_OK;
    }
    int min_b[3], max_b[3];
    for (int k = 0; k < 3; k++) { min_b[k] = (int)floorf(mn[k] * inv); max_b[k] = (int)floorf(mx[k] * inv); }
    const int div0 = max_b[0] - min_b[0] + 1, div1 = max_b[1] - min_b[1] + 1;
    const size_t tb = scl_voxel_temp_bytes(n);
    CK(e->vg_temp.ensure(tb));
    for (int k = 0; k < 2; k++) { CK(e->vg_keys[k].ensure((size_t)n * 4)); CK(e->vg_vals[k].ensure((size_t)n * 4)); }
    CK(e->vg_head.ensure((size_t)n * 4)); CK(e->vg_ord.ensure((size_t)n * 4));
    CK(scl_launch_voxel_grid(d_pts, n, stride, inv, min_b, div0, div0 * div1, 32, e->vg_keys[0].as<uint32_t>(), e->vg_keys[1].as<uint32_t>(),
                             e->vg_vals[0].as<int>(), e->vg_vals[1].as<int>(), e->vg_head.as<int>(), e->vg_ord.as<int>(), e->vg_temp.p, tb,
                             e->vg_out.p, d_nout, e->stream));
    CK(cudaMemcpyAsync(n_out, d_nout, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SCL_OK;
}
} // namespace

int scl_voxel_grid(scl_engine* e, const void* pts, int n, int stride_bytes, float leaf, float* out_xyzi, int* n_out)
{
    LOCK();
    if (!n_out || (n > 0 && (!pts || !out_xyzi))) FAIL(SCL_ERR_INVALID, "null argument");
    if (stride_bytes < 16 || stride_bytes % 16) FAIL(SCL_ERR_UNSUPPORTED, "points must be 16-byte aligned x,y,z,intensity records");
    *n_out = 0;
    if (n <= 0) return SCL_OK;
    CK(e->vg_in.ensure((size_t)n * stride_bytes));
    CK(cudaMemcpyAsync(e->vg_in.p, pts, (size_t)n * stride_bytes, cudaMemcpyHostToDevice, e->stream));
    int rc = voxel_grid_dev(e, e->vg_in.p, n, stride_bytes, leaf, n_out); if (rc) return rc;
    if (*n_out > 0) {
        CK(cudaMemcpyAsync(out_xyzi, e->vg_out.p, (size_t)*n_out * 16, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    return SCL_OK;
}

// The reference's producer step as one call (distributedMapping.h:996-1003): downSizeFilterDes.filter(keyFrame) and then
// makeAndSaveDescriptorAndKey(filtered, robot, index). The filtered cloud never leaves the device.
int scl_build_insert_filtered(scl_engine* e, const void* pts, int n, int stride_bytes, float leaf, int8_t robot, int index,
                              float* out_desc, int* n_filtered)
{
    LOCK();
    if (n < 0) FAIL(SCL_ERR_INVALID, "n < 0");
    if (n > 0 && !pts) FAIL(SCL_ERR_INVALID, "null cloud");
    if (stride_bytes < 16 || stride_bytes % 16) FAIL(SCL_ERR_UNSUPPORTED, "points must be 16-byte aligned x,y,z,intensity records");
    int m = 0;
    if (n > 0) {
        CK(e->vg_in.ensure((size_t)n * stride_bytes));
        CK(cudaMemcpyAsync(e->vg_in.p, pts, (size_t)n * stride_bytes, cudaMemcpyHostToDevice, e->stream));
        int rc = voxel_grid_dev(e, e->vg_in.p, n, stride_bytes, leaf, &m); if (rc) return rc;
    } else {
        CK(e->vg_out.ensure(16));
    }
    if (n_filtered) *n_filtered = m;
    const int32_t off[2] = {0, m};
    float* where = nullptr;
    int rc = build_dev(e, e->vg_out.p, off, 1, 16, 1, &robot, &index, nullptr, nullptr, nullptr, &where); if (rc) return rc;
    if (out_desc) CK(cudaMemcpyAsync(out_desc, where, (size_t)e->RS() * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SCL_OK;
}

int scl_assemble_submap(scl_engine* e, const void* pts, const int* offsets, int n_clouds, int stride_bytes, const float* poses6,
                        float leaf, float* out_xyzi, int* n_out)
{
    LOCK();
    if (!n_out || !offsets || n_clouds < 0 || (n_clouds > 0 && (!poses6 || !pts))) FAIL(SCL_ERR_INVALID, "null argument");
    if (stride_bytes < 16 || stride_bytes % 16) FAIL(SCL_ERR_UNSUPPORTED, "points must be 16-byte aligned x,y,z,intensity records");
    *n_out = 0;
    const int total = n_clouds > 0 ? offsets[n_clouds] : 0;
    for (int c = 0; c < n_clouds; c++)
        if (offsets[c] < 0 || offsets[c + 1] < offsets[c]) FAIL(SCL_ERR_INVALID, "offsets must start at >= 0 and be non-decreasing");
    if (total <= 0) return SCL_OK;
    if (!out_xyzi) FAIL(SCL_ERR_INVALID, "null output");
    /* pcl::getTransformation(x, y, z, roll, pitch, yaw) in float with libm, as the reference calls it (:241) */
    std::vector<float> T((size_t)n_clouds * 12);
    int max_points = 0;
    for (int c = 0; c < n_clouds; c++) {
        const float* p = poses6 + (size_t)c * 6;
        const float A = cosf(p[5]), B = sinf(p[5]), C = cosf(p[4]), D = sinf(p[4]), E = cosf(p[3]), F = sinf(p[3]), DE = D * E, DF = D * F;
        float* t = &T[(size_t)c * 12];
        t[0] = A * C; t[1] = A * DF - B * E; t[2] = B * F + A * DE; t[3] = p[0];
        t[4] = B * C; t[5] = A * E + B * DF; t[6] = B * DE - A * F; t[7] = p[1];
        t[8] = -D;    t[9] = C * F;          t[10] = C * E;         t[11] = p[2];
        max_points = std::max(max_points, offsets[c + 1] - offsets[c]);
    }
    CK(e->vg_in.ensure((size_t)total * stride_bytes)); CK(e->vg_world.ensure((size_t)total * 16));
    CK(e->vg_T.ensure(T.size() * 4)); CK(e->vg_off.ensure((size_t)(n_clouds + 1) * 4));
    CK(cudaMemcpyAsync(e->vg_in.p, pts, (size_t)total * stride_bytes, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->vg_T.p, T.data(), T.size() * 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->vg_off.p, offsets, (size_t)(n_clouds + 1) * 4, cudaMemcpyHostToDevice, e->stream));
    CK(scl_launch_transform_concat(e->vg_in.p, e->vg_off.as<int>(), n_clouds, max_points, stride_bytes, e->vg_T.as<float>(), e->vg_world.p, e->stream));
    if (leaf > 0.0f) {
        int rc = voxel_grid_dev(e, e->vg_world.p, total, 16, leaf, n_out); if (rc) return rc;   /* T is consumed before voxel_grid_dev synchronises */
        if (*n_out > 0) CK(cudaMemcpyAsync(out_xyzi, e->vg_out.p, (size_t)*n_out * 16, cudaMemcpyDeviceToHost, e->stream));
    } else {
        *n_out = total;
        CK(cudaMemcpyAsync(out_xyzi, e->vg_world.p, (size_t)total * 16, cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    return SCL_OK;
}

// ---- device-resident keyframe clouds + the intra-robot verification as one call -------------------------------------------
int scl_store_keyframe_cloud(scl_engine* e, int key, const void* pts, int n, int stride_bytes)
{
    LOCK();
    if (n < 0 || (n > 0 && !pts)) FAIL(SCL_ERR_INVALID, "bad cloud");
    if (stride_bytes < 16 || stride_bytes % 16) FAIL(SCL_ERR_UNSUPPORTED, "points must be 16-byte aligned x,y,z,intensity records");
    if (key != (int)e->kf_off.size() - 1) FAIL(SCL_ERR_INVALID, "keyframe clouds are stored in key order (key = number stored so far), like keyFrameArray.push_back");
    const size_t need = (e->kf_points + (size_t)n) * 16;
    if (need > e->kf_arena.cap) {
        /* grow the arena (append-only; doubling) keeping what is stored */
        DevBuf bigger;
        CK(bigger.ensure(need * 2 + (1 << 20)));
        if (e->kf_points) CK(cudaMemcpyAsync(bigger.p, e->kf_arena.p, e->kf_points * 16, cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->kf_arena.release();
        e->kf_arena = bigger;
    }
    if (n > 0) {
        CK(e->vg_in.ensure((size_t)n * stride_bytes));
        CK(cudaMemcpyAsync(e->vg_in.p, pts, (size_t)n * stride_bytes, cudaMemcpyHostToDevice, e->stream));
        CK(scl_launch_pack_xyzi(e->vg_in.p, n, stride_bytes, static_cast<unsigned char*>(e->kf_arena.p) + e->kf_points * 16, e->stream));
        CK(cudaStreamSynchronize(e->stream));            /* the caller's cloud may go away */
    }
    e->kf_points += (size_t)n;
    e->kf_off.push_back((int)e->kf_points);
    return SCL_OK;
}

int scl_keyframe_clouds(scl_engine* e) { if (!e) return -1; std::lock_guard<std::mutex> lk(e->mu); return (int)e->kf_off.size() - 1; }

namespace {
// loopFindNearKeyframes (distributedMapping.h:1163-1186) on the stored clouds: keys key-search .. key+search (clipped), each
// moved by its pose (transformPointCloud, :234-253), concatenated, down-sampled (downSizeFilterICP). Result: packed float4
// in e->vg_out (leaf > 0) or e->vg_world (leaf <= 0), count in *n_out.
int near_keyframes_dev(scl_engine* e, int key, int search, const float* poses6, int n_poses, float leaf, int* n_out, const void** where)
{
    *n_out = 0; *where = nullptr;
    const int n_clouds_all = (int)e->kf_off.size() - 1;
    int k0 = key - search, k1 = key + search;
    if (k0 < 0) k0 = 0;
    if (k1 >= n_clouds_all) k1 = n_clouds_all - 1;
    if (k1 < k0) return SCL_OK;
    if (k1 >= n_poses) FAIL(SCL_ERR_RANGE, "a stored keyframe has no pose");
    const int nc = k1 - k0 + 1;
    std::vector<float> T((size_t)nc * 12);
    std::vector<int> off((size_t)nc + 1);
    int max_points = 0;
    for (int c = 0; c < nc; c++) {
        const float* p = poses6 + (size_t)(k0 + c) * 6;
        /* pcl::getTransformation(x, y, z, roll, pitch, yaw) in float with libm, as the reference calls it (:241) */
        const float A = cosf(p[5]), B = sinf(p[5]), C = cosf(p[4]), D = sinf(p[4]), E = cosf(p[3]), F = sinf(p[3]), DE = D * E, DF = D * F;
        float* t = &T[(size_t)c * 12];
        t[0] = A * C; t[1] = A * DF - B * E; t[2] = B * F + A * DE; t[3] = p[0];
        t[4] = B * C; t[5] = A * E + B * DF; t[6] = B * DE - A * F; t[7] = p[1];
        t[8] = -D;    t[9] = C * F;          t[10] = C * E;         t[11] = p[2];
        off[c] = e->kf_off[k0 + c] - e->kf_off[k0];
        max_points = std::max(max_points, e->kf_off[k0 + c + 1] - e->kf_off[k0 + c]);
    }
    off[nc] = e->kf_off[k1 + 1] - e->kf_off[k0];
    const int total = off[nc];
    if (total <= 0) return SCL_OK;
    CK(e->vg_world.ensure((size_t)total * 16));
    CK(e->vg_T.ensure(T.size() * 4)); CK(e->vg_off.ensure((size_t)(nc + 1) * 4));
    CK(cudaMemcpyAsync(e->vg_T.p, T.data(), T.size() * 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->vg_off.p, off.data(), (size_t)(nc + 1) * 4, cudaMemcpyHostToDevice, e->stream));
    CK(scl_launch_transform_concat(static_cast<unsigned char*>(e->kf_arena.p) + (size_t)e->kf_off[k0] * 16, e->vg_off.as<int>(), nc, max_points, 16,
                                   e->vg_T.as<float>(), e->vg_world.p, e->stream));
    if (leaf > 0.0f) {
        int rc = voxel_grid_dev(e, e->vg_world.p, total, 16, leaf, n_out); if (rc) return rc;    /* synchronises: T and off are consumed */
        *where = e->vg_out.p;
    } else {
        CK(cudaStreamSynchronize(e->stream));
        *n_out = total; *where = e->vg_world.p;
    }
    return SCL_OK;
}
} // namespace

int scl_verify_intra(scl_engine* e, int key_cur, int key_pre, int search_num, const float* poses6, int n_poses, float leaf,
                     const scl_icp_params* icp, float fitness_threshold, scl_intra_result* out)
{
    LOCK();
    if (!poses6 || !icp || !out || search_num < 0) FAIL(SCL_ERR_INVALID, "bad arguments");
    const int n_clouds = (int)e->kf_off.size() - 1;
    if (key_cur < 0 || key_cur >= n_clouds || key_pre < 0 || key_pre >= n_clouds || key_cur >= n_poses) FAIL(SCL_ERR_RANGE, "keyframe out of range");
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < 4; i++) out->T[5 * i] = 1.0f;
    out->fitness = 3.402823466e+38f;
    /* the current keyframe (searchNum 0) and the history submap, both in the world frame and down-sampled; the clouds never leave the device */
    int n_src = 0, n_tgt = 0; const void* where = nullptr;
    int rc = near_keyframes_dev(e, key_cur, 0, poses6, n_poses, leaf, &n_src, &where); if (rc) return rc;
    rc = scl_icp_set_cloud_dev(e, 0, where, n_src); if (rc) return rc;
    rc = near_keyframes_dev(e, key_pre, search_num, poses6, n_poses, leaf, &n_tgt, &where); if (rc) return rc;
    rc = scl_icp_set_cloud_dev(e, 1, where, n_tgt); if (rc) return rc;
    out->n_src = n_src; out->n_tgt = n_tgt;
    if (n_src < 300 || n_tgt < 1000) return SCL_OK;                  /* distributedMapping.h:1102: too little to verify, no loop */
    rc = scl_icp_device(e, n_src, n_tgt, icp, out->T, &out->fitness, &out->converged, &out->iterations); if (rc) return rc;
    out->accepted = (out->converged && !(out->fitness > fitness_threshold)) ? 1 : 0;   /* :1122 */
    return SCL_OK;
}

} // extern "C"
