// sharded.cu — one Scan Context database over the GPUs of a box, inside ONE process (scl_create_sharded & co,
// include/scl_engine.h). distributed_mapping constructs a single descriptor object
// (/root/reference/include/distributedMapping.h:333,402-405); this is that object when the keyframe database is
// sharded by keyframe index (north_star): one scl_engine per device holding the keys with key mod world == rank, their
// exchange buffers mapped into each other (peer access, no IPC needed inside a process), and every query driven as the
// same sharded step the multi-process form runs (scl_shard_query_submit): each device uploads 1 / world of the batch,
// the gather kernel spreads it over NVLink, K2 + K3 per shard, exchange + global top-K, K4 by the owners, exchange +
// winner scan. One host thread drives all devices: nothing it calls blocks before every device has its work.
#include "engine_internal.h"

#include <algorithm>

struct scl_sharded {
    int world = 0;
    std::vector<scl_engine*> eng;
    std::vector<int> devs;
    std::mutex mu;
    std::string err;
    scl_params p;
    int n = 0;                                         /* keyframes over all shards */
    std::vector<std::pair<int8_t, int>> index;         /* (robot, index) of every global key (descriptor.h:1599) */
    int max_q = 0, max_k = 0;
    int tree_counter = 0, n_tree = 0;                  /* descriptor.h:1691-1703 */
    struct Slot {                                      /* one batch in flight: page-locked staging on the host */
        float* q = nullptr; size_t q_bytes = 0;
        int32_t* ids = nullptr;
        unsigned char* res = nullptr; size_t res_bytes = 0;
        scl_batch_result user{}; int Q = 0, Qp = 0, K = 0; bool busy = false;
        int tickets[16] = {};
    };
    Slot slot[scl_engine::kLanes];
    long long next = 0;
};

#define SFAIL(code, msg) do { s->err = (msg); return (code); } while (0)

namespace {
int fail_from(scl_sharded* s, scl_engine* e, int rc) { s->err = scl_last_error(e); return rc; }

size_t res_layout(int Qp, int K, size_t off[7])
{
    const size_t QK = (size_t)Qp * K;
    const size_t sz[7] = {QK * 4, QK * 4, QK * 8, QK * 4, (size_t)Qp * 4, (size_t)Qp * 8, (size_t)Qp * 4};
    size_t o = 0;
    for (int i = 0; i < 7; i++) { off[i] = o; o += (sz[i] + 15) / 16 * 16; }
    return o;
}

int local_bound(int n_db, int rank, int world) { return n_db <= rank ? 0 : (n_db - rank + world - 1) / world; }   /* keys < n_db held by `rank` */
} // namespace

extern "C" {

int scl_create_sharded(const scl_params* p, int ndev, const int* devs, int max_q, int max_k, scl_sharded** out)
{
    if (!p || !out || ndev < 1 || ndev > 16 || !devs || max_q < 1 || max_k < 1 || max_k > 32) return SCL_ERR_INVALID;
    *out = nullptr;
    scl_sharded* s = new scl_sharded();
    s->world = ndev; s->p = *p; s->max_q = (max_q + 3) / 4 * 4; s->max_k = max_k;
    s->devs.assign(devs, devs + ndev);
    int rc = SCL_OK;
    for (int r = 0; r < ndev && rc == SCL_OK; r++) {
        scl_engine* e = nullptr;
        rc = scl_create(p, devs[r], &e);
        if (rc == SCL_OK) { s->eng.push_back(e); rc = scl_set_shard(e, r, ndev); }
    }
    if (rc == SCL_OK && ndev > 1) {
        for (int a = 0; a < ndev && rc == SCL_OK; a++) {
            if (cudaSetDevice(devs[a]) != cudaSuccess) { rc = SCL_ERR_CUDA; break; }
            for (int b = 0; b < ndev; b++) {
                if (a == b || devs[a] == devs[b]) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devs[a], devs[b]) != cudaSuccess || !can) { rc = SCL_ERR_UNSUPPORTED; break; }
                const cudaError_t ce = cudaDeviceEnablePeerAccess(devs[b], 0);
                if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) { rc = SCL_ERR_CUDA; break; }
                (void)cudaGetLastError();
            }
        }
        unsigned char handle[64];
        std::vector<void*> bufs(ndev);
        for (int r = 0; r < ndev && rc == SCL_OK; r++) {
            rc = scl_xchg_create(s->eng[r], ndev, s->max_q, max_k, handle);
            bufs[r] = scl_xchg_buffer(s->eng[r]);
        }
        for (int r = 0; r < ndev && rc == SCL_OK; r++) rc = scl_xchg_open_local(s->eng[r], ndev, r, bufs.data());
    }
    if (rc != SCL_OK) {
        for (scl_engine* e : s->eng) scl_destroy(e);
        delete s;
        return rc;
    }
    *out = s;
    return SCL_OK;
}

int scl_sharded_destroy(scl_sharded* s)
{
    if (!s) return SCL_ERR_INVALID;
    for (auto& sl : s->slot) {
        if (sl.q) cudaFreeHost(sl.q);
        if (sl.ids) cudaFreeHost(sl.ids);
        if (sl.res) cudaFreeHost(sl.res);
    }
    for (scl_engine* e : s->eng) scl_destroy(e);
    delete s;
    return SCL_OK;
}

const char* scl_sharded_last_error(scl_sharded* s) { return s ? s->err.c_str() : "null handle"; }
int scl_sharded_world(scl_sharded* s) { return s ? s->world : -1; }
scl_engine* scl_sharded_engine(scl_sharded* s, int rank) { return (s && rank >= 0 && rank < s->world) ? s->eng[rank] : nullptr; }
int scl_sharded_size(scl_sharded* s) { if (!s) return -1; std::lock_guard<std::mutex> lk(s->mu); return s->n; }

int scl_sharded_get_index(scl_sharded* s, int key, int8_t* robot, int* index)
{
    if (!s || !robot || !index) return SCL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(s->mu);
    if (key < 0 || key >= s->n) { *robot = -1; *index = -1; return SCL_OK; }       /* the reference's getIndex(-1) is UB */
    *robot = s->index[key].first; *index = s->index[key].second;
    return SCL_OK;
}

/* saveDescriptorAndKey (descriptor.h:1572-1585) for n descriptors: global key g = size + i goes to shard g mod world */
int scl_sharded_insert_batch(scl_sharded* s, const float* descs, int n, const int8_t* robots, const int32_t* indices)
{
    if (!s || n < 0 || (n > 0 && !descs)) return SCL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(s->mu);
    const size_t RS = (size_t)s->p.num_ring * s->p.num_sector;
    std::vector<float> rows; std::vector<int8_t> rb; std::vector<int32_t> ix;
    for (int r = 0; r < s->world; r++) {
        rows.clear(); rb.clear(); ix.clear();
        for (int i = 0; i < n; i++) {
            const int g = s->n + i;
            if (g % s->world != r) continue;
            rows.insert(rows.end(), descs + (size_t)i * RS, descs + (size_t)(i + 1) * RS);
            rb.push_back(robots ? robots[i] : (int8_t)0); ix.push_back(indices ? indices[i] : g);
        }
        if (rb.empty()) continue;
        const int rc = scl_insert_batch(s->eng[r], rows.data(), (int)rb.size(), rb.data(), ix.data());
        if (rc) return fail_from(s, s->eng[r], rc);
    }
    for (int i = 0; i < n; i++) s->index.emplace_back(robots ? robots[i] : (int8_t)0, indices ? indices[i] : s->n + i);
    s->n += n;
    return SCL_OK;
}

/* makeAndSaveDescriptorAndKey (descriptor.h:1604-1611): built and kept by the shard that owns the next key */
int scl_sharded_build_insert(scl_sharded* s, const void* pts, int n, int stride_bytes, int8_t robot, int index, float* out_desc)
{
    if (!s) return SCL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(s->mu);
    scl_engine* e = s->eng[s->n % s->world];
    const int rc = scl_build_insert(e, pts, n, stride_bytes, robot, index, out_desc);
    if (rc) return fail_from(s, e, rc);
    s->index.emplace_back(robot, index);
    s->n++;
    return SCL_OK;
}

int scl_sharded_get_descriptor(scl_sharded* s, int key, float* out_desc)
{
    if (!s || !out_desc) return SCL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(s->mu);
    if (key < 0 || key >= s->n) SFAIL(SCL_ERR_RANGE, "key out of range");
    scl_engine* e = s->eng[key % s->world];
    const int rc = scl_get_descriptor(e, key / s->world, out_desc);
    return rc ? fail_from(s, e, rc) : SCL_OK;
}

/* The batched query (scl_query_batch's meaning) on the sharded database, pipelined: up to scl_num_lanes() batches in flight.
 * q->q_desc: Q descriptors in host memory (any kind; they are staged through page-locked memory); q->q_ids optional. */
int scl_sharded_query_submit(scl_sharded* s, const scl_batch_query* q, scl_batch_result* r, int* ticket)
{
    if (!s || !q || !r || !ticket || !q->q_desc) return SCL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(s->mu);
    const int Q = q->Q, K = q->K;
    if (Q < 1 || K < 1 || K > s->max_k) SFAIL(SCL_ERR_INVALID, "Q >= 1 and K within the size given to scl_create_sharded");
    const int Qp = (Q + 3) / 4 * 4;                    /* the exchange moves 16-byte pieces: Q * K a multiple of 4 */
    if (Qp > s->max_q) SFAIL(SCL_ERR_INVALID, "batch larger than the size given to scl_create_sharded");
    scl_sharded::Slot& sl = s->slot[s->next % scl_engine::kLanes];
    if (sl.busy) SFAIL(SCL_ERR_INVALID, "every lane has a batch in flight: wait for the oldest one first");
    const size_t RS = (size_t)s->p.num_ring * s->p.num_sector;
    const size_t qb = (size_t)Qp * RS * 4;
    if (qb > sl.q_bytes) {
        if (sl.q) cudaFreeHost(sl.q);
        if (sl.ids) cudaFreeHost(sl.ids);
        sl.q = nullptr; sl.ids = nullptr; sl.q_bytes = 0;
        if (cudaMallocHost(&sl.q, qb) != cudaSuccess || cudaMallocHost(&sl.ids, (size_t)Qp * 4) != cudaSuccess) SFAIL(SCL_ERR_NOMEM, "page-locked staging");
        sl.q_bytes = qb;
    }
    size_t off[7];
    const size_t rb = res_layout(Qp, K, off);
    if (rb > sl.res_bytes) {
        if (sl.res) cudaFreeHost(sl.res);
        sl.res = nullptr; sl.res_bytes = 0;
        if (cudaMallocHost(&sl.res, rb) != cudaSuccess) SFAIL(SCL_ERR_NOMEM, "page-locked staging");
        sl.res_bytes = rb;
    }
    memcpy(sl.q, q->q_desc, (size_t)Q * RS * 4);
    if (Qp > Q) memset(sl.q + (size_t)Q * RS, 0, (size_t)(Qp - Q) * RS * 4);        /* padding queries: empty descriptors */
    if (q->q_ids) { memcpy(sl.ids, q->q_ids, (size_t)Q * 4); for (int i = Q; i < Qp; i++) sl.ids[i] = -1; }
    sl.user = *r; sl.Q = Q; sl.Qp = Qp; sl.K = K;
    scl_batch_result dev0{reinterpret_cast<int32_t*>(sl.res + off[0]), reinterpret_cast<float*>(sl.res + off[1]), reinterpret_cast<double*>(sl.res + off[2]),
                          reinterpret_cast<int32_t*>(sl.res + off[3]), reinterpret_cast<int32_t*>(sl.res + off[4]), reinterpret_cast<double*>(sl.res + off[5]),
                          reinterpret_cast<int32_t*>(sl.res + off[6])};
    scl_batch_result none{};
    const int n_db = q->n_db < 0 ? 0 : (q->n_db > s->n ? s->n : q->n_db);
    for (int rk = 0; rk < s->world; rk++) {
        scl_batch_query qq{sl.q, q->q_ids ? sl.ids : nullptr, Qp, K, local_bound(n_db, rk, s->world), q->metric};
        int rc;
        if (s->world == 1) rc = scl_query_batch_submit(s->eng[0], &qq, &dev0, &sl.tickets[0]);
        else rc = scl_shard_query_submit(s->eng[rk], &qq, rk == 0 ? &dev0 : &none, &sl.tickets[rk]);
        if (rc) return fail_from(s, s->eng[rk], rc);       /* (a failure half-way leaves the ranks out of step: the handle is unusable) */
    }
    sl.busy = true;
    *ticket = (int)(s->next & 0x7fffffff);
    s->next++;
    return SCL_OK;
}

int scl_sharded_query_wait(scl_sharded* s, int ticket)
{
    if (!s || ticket < 0) return SCL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(s->mu);
    scl_sharded::Slot& sl = s->slot[ticket % scl_engine::kLanes];
    if (!sl.busy) SFAIL(SCL_ERR_INVALID, "no batch in flight for this ticket");
    for (int rk = 0; rk < s->world; rk++) {
        const int rc = scl_query_batch_wait(s->eng[rk], sl.tickets[rk]);
        if (rc) return fail_from(s, s->eng[rk], rc);
    }
    sl.busy = false;
    size_t off[7];
    res_layout(sl.Qp, sl.K, off);
    const size_t QK = (size_t)sl.Q * sl.K;
    const scl_batch_result& u = sl.user;
    if (u.cand_ids) memcpy(u.cand_ids, sl.res + off[0], QK * 4);
    if (u.cand_d2) memcpy(u.cand_d2, sl.res + off[1], QK * 4);
    if (u.cand_dist) memcpy(u.cand_dist, sl.res + off[2], QK * 8);
    if (u.cand_shift) memcpy(u.cand_shift, sl.res + off[3], QK * 4);
    if (u.best_id) memcpy(u.best_id, sl.res + off[4], (size_t)sl.Q * 4);
    if (u.best_dist) memcpy(u.best_dist, sl.res + off[5], (size_t)sl.Q * 8);
    if (u.best_shift) memcpy(u.best_shift, sl.res + off[6], (size_t)sl.Q * 4);
    return SCL_OK;
}

int scl_sharded_query_batch(scl_sharded* s, const scl_batch_query* q, scl_batch_result* r)
{
    int t = 0;
    const int rc = scl_sharded_query_submit(s, q, r, &t);
    return rc ? rc : scl_sharded_query_wait(s, t);
}

namespace {
int query_one(scl_sharded* s, int cur, int n_db, int metric, int32_t* ids, double* dist, int32_t* shift)
{
    const int K = s->p.num_candidates;
    std::vector<float> d((size_t)s->p.num_ring * s->p.num_sector);
    int rc = scl_sharded_get_descriptor(s, cur, d.data()); if (rc) return rc;
    scl_batch_query q{d.data(), &cur, 1, K, n_db, metric};
    scl_batch_result r{ids, nullptr, dist, shift, nullptr, nullptr, nullptr};
    return scl_sharded_query_batch(s, &q, &r);
}
} // namespace

/* detectIntraLoopClosureID (descriptor.h:1613-1674) on the sharded database */
int scl_sharded_query_intra(scl_sharded* s, int cur, int* id, float* second)
{
    if (!s || !id || !second) return SCL_ERR_INVALID;
    *id = -1; *second = 0.0f;
    const int n = scl_sharded_size(s), K = s->p.num_candidates;
    if (cur < 0 || cur >= n) SFAIL(SCL_ERR_RANGE, "query key out of range");
    if (K > s->max_k) SFAIL(SCL_ERR_INVALID, "num_candidates exceeds the max_k given to scl_create_sharded");
    if (cur < s->p.num_exclude_recent + K + 1) return SCL_OK;               /* :1620 */
    int32_t ids[32]; double dist[32]; int32_t shift[32];
    const int rc = query_one(s, cur, cur - s->p.num_exclude_recent, 1, ids, dist, shift); if (rc) return rc;   /* :1627; libnabo flavour */
    float minDis = 10000000.0f; int minIndex = -1, minBias = 0;             /* :1637-1659: minDis is a float there */
    for (int i = 0; i < K; i++) {
        if (ids[i] < 0) continue;
        if (dist[i] < (double)minDis) { minDis = (float)dist[i]; minIndex = ids[i]; minBias = shift[i]; }
    }
    if ((double)minDis < s->p.dist_thres) { *id = minIndex; *second = (float)minBias; }
    return SCL_OK;
}

/* detectInterLoopClosureID (descriptor.h:1676-1756) on the sharded database. Where the searched range holds fewer than K
 * keys the reference reads result slots its tree never filled (:1710,1723); here they simply stay empty. */
int scl_sharded_query_inter(scl_sharded* s, int cur, int* id, float* second)
{
    if (!s || !id || !second) return SCL_ERR_INVALID;
    *id = -1; *second = 0.0f;
    const int n = scl_sharded_size(s), K = s->p.num_candidates;
    if (cur < 0 || cur >= n) SFAIL(SCL_ERR_RANGE, "query key out of range");
    if (K > s->max_k) SFAIL(SCL_ERR_INVALID, "num_candidates exceeds the max_k given to scl_create_sharded");
    if (n < s->p.num_exclude_recent + 1) return SCL_OK;                     /* :1684 */
    if (s->tree_counter % s->p.tree_making_period == 0) s->n_tree = n - s->p.num_exclude_recent;   /* :1691-1699 */
    s->tree_counter++;
    int32_t ids[32]; double dist[32]; int32_t shift[32];
    const int rc = query_one(s, cur, s->n_tree, 0, ids, dist, shift); if (rc) return rc;
    double min_dist = 10000000; int nn_align = 0, nn_idx = -1;
    for (int i = 0; i < K; i++)
        if (ids[i] >= 0 && dist[i] < min_dist && ids[i] != cur) { min_dist = dist[i]; nn_align = shift[i]; nn_idx = ids[i]; }
    if (min_dist < s->p.dist_thres) *id = nn_idx;
    const double unit = 360.0 / double(s->p.num_sector);
    *second = (float)(nn_align * unit * M_PI / 180.0);                      /* :1752 */
    return SCL_OK;
}

} // extern "C"
