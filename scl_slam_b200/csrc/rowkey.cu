// rowkey.cu — the row-key candidate search of the reference's second descriptor family (class lidar_iris_descriptor,
// /root/reference/include/descriptor.h:462-1302) behind include/scl_rowkey.h: per-robot key matrices in HBM, the two key-set
// selection rules of detectIntraLoopClosureID (:1087-1114) and detectInterLoopClosureID (:1150-1209), and the K3 kernels
// (k3_knn.cu exact, k3_knn_tc.cu tensor-core prefilter + exact re-rank) in libnabo's flavour for the kNN itself.
// The reference rebuilds a KD-tree over a fresh copy of the selected key columns on EVERY call (:1099-1104, :1159-1199);
// here a query is one pass over the resident key arrays, and the concatenation of several robots' matrices (:1164-1179) is
// a top-K merge of per-robot lists whose ids carry the concatenation offsets.
#include "common.cuh"
#include "kernels.h"
#include "engine_internal.h"
#include "../../include/scl_rowkey.h"

#include <cfloat>
#include <cmath>
#include <mutex>
#include <string>
#include <vector>

namespace {

// squared norms of keys [k_lo, k_hi) (tensor-core prefilter only) and the largest of them (non-negative floats order as ints)
__global__ void rowkey_norm_kernel(const float* __restrict__ keys, int k_lo, int k_hi, int R, float* __restrict__ knorm, float* __restrict__ kn2max)
{
    const int k = k_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_hi) return;
    float n = 0.0f;
    for (int d = 0; d < R; d++) { const float v = __ldg(keys + (size_t)k * R + d); n = fmaf(v, v, n); }
    knorm[k] = n;
    if (n == n) atomicMax(reinterpret_cast<int*>(kn2max), __float_as_int(n));
}

// per-query lists of the selected key sets -> one list: every set's ids already carry its concatenation offset; this adds
// newLocal2Global (descriptor.h:1175,1188) for the caller
__global__ void rowkey_map_kernel(const int32_t* __restrict__ concat_idx, int n, const int32_t* __restrict__ l2g, int n_l2g, int32_t* __restrict__ global_key)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = concat_idx[i];
    global_key[i] = (c >= 0 && c < n_l2g) ? l2g[c] : -1;
}

}  // namespace

cudaError_t scl_launch_key_norms(const float* keys, int k_lo, int k_hi, int R, float* knorm, float* kn2max, cudaStream_t stream)
{
    if (k_hi <= k_lo) return cudaSuccess;
    rowkey_norm_kernel<<<(k_hi - k_lo + 127) / 128, 128, 0, stream>>>(keys, k_lo, k_hi, R, knorm, kn2max);
    return cudaGetLastError();
}

namespace {

struct KeyStore {
    float* keys = nullptr; float* knorm = nullptr; unsigned char* kimg = nullptr;
    int n = 0, cap = 0, norm_n = 0, img_n = 0, img_cap = 0;
};

}  // namespace

struct scl_rowkey {
    scl_rowkey_params p;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::mutex mu;
    std::string err;
    std::vector<KeyStore> robots;
    std::vector<std::vector<int>> local2global;              /* descriptor.h:1299 */
    std::vector<std::pair<int8_t, int>> index;               /* irisFeatureIndexs, descriptor.h:1300 */
    float* d_kn2max = nullptr;
    DevBuf qkeys, part_ids, part_d2, tickets, blk_ids, blk_d2, out_ids, out_d2, out_gk, l2g;
    DevBuf tc_queues, tc_queue_cnt, tc_slots, tc_fail_list, tc_fail_count;
    long long tc_calls = 0; bool tc_state_clean = false; int tc_slots_rows = 0;
    long long stat_tc = 0;
    std::vector<int> l2g_host; int l2g_from = -2; size_t l2g_sizes_sig = 0;   /* the mapping currently in l2g */
    DevBuf* all[15] = {&qkeys, &part_ids, &part_d2, &tickets, &blk_ids, &blk_d2, &out_ids, &out_d2, &out_gk, &l2g,
                       &tc_queues, &tc_queue_cnt, &tc_slots, &tc_fail_list, &tc_fail_count};
};

#define RK_LOCK() std::lock_guard<std::mutex> lk(e->mu); cudaSetDevice(e->device)

namespace {

struct KeySet { int robot; int n; int offset; };

// the key sets a query sees, in concatenation order, and the concatenated local-to-global map
void select_sets(scl_rowkey* e, int from_robot, int n_limit, std::vector<KeySet>& sets, int* total)
{
    sets.clear();
    int off = 0;
    if (from_robot < 0) {
        /* intra: this robot's first n_limit keys (descriptor.h:1101-1103) */
        int n = e->robots[e->p.this_id].n;
        if (n_limit < n) n = n_limit;
        if (n > 0) { sets.push_back(KeySet{e->p.this_id, n, 0}); off = n; }
    } else if (from_robot == e->p.this_id) {
        /* a query of this robot: every other robot that has keys, ascending (descriptor.h:1164-1179) */
        for (int i = 0; i < e->p.robot_num; i++)
            if (i != e->p.this_id && !e->local2global[i].empty()) { sets.push_back(KeySet{i, (int)e->local2global[i].size(), off}); off += (int)e->local2global[i].size(); }
    } else {
        /* a query of another robot: this robot's keys (descriptor.h:1180-1191) */
        if (!e->local2global[e->p.this_id].empty()) { sets.push_back(KeySet{e->p.this_id, (int)e->local2global[e->p.this_id].size(), 0}); off = (int)e->local2global[e->p.this_id].size(); }
    }
    *total = off;
}

int grow(scl_rowkey* e, KeyStore& s, int need)
{
    if (need <= s.cap) return SCL_OK;
    int cap = s.cap ? s.cap : 1024;
    while (cap < need) cap *= 2;
    const int R = e->p.rows;
    float *nk = nullptr, *nn = nullptr;
    CK(cudaMalloc(&nk, (size_t)cap * R * 4));
    cudaError_t er = cudaMalloc(&nn, (size_t)cap * 4);
    if (er != cudaSuccess) { cudaFree(nk); CK(er); }
    if (s.n > 0) {
        er = cudaMemcpyAsync(nk, s.keys, (size_t)s.n * R * 4, cudaMemcpyDeviceToDevice, e->stream);
        if (er == cudaSuccess && s.norm_n > 0) er = cudaMemcpyAsync(nn, s.knorm, (size_t)s.norm_n * 4, cudaMemcpyDeviceToDevice, e->stream);
        if (er == cudaSuccess) er = cudaStreamSynchronize(e->stream);
        if (er != cudaSuccess) { cudaFree(nk); cudaFree(nn); CK(er); }
    }
    if (s.keys) cudaFree(s.keys);
    if (s.knorm) cudaFree(s.knorm);
    s.keys = nk; s.knorm = nn; s.cap = cap;
    return SCL_OK;
}

int append(scl_rowkey* e, int robot, const float* keys_host, int n)
{
    KeyStore& s = e->robots[robot];
    { int rc = grow(e, s, s.n + n); if (rc) return rc; }
    CK(cudaMemcpyAsync(s.keys + (size_t)s.n * e->p.rows, keys_host, (size_t)n * e->p.rows * 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));                       /* the caller's buffer is free on return, as in the reference */
    s.n += n;
    return SCL_OK;
}

// norms and tensor-core image of the keys appended since the last tensor-core query
int sync_image(scl_rowkey* e, KeyStore& s)
{
    const int R = e->p.rows;
    if (s.norm_n < s.n) {
        CK(scl_launch_key_norms(s.keys, s.norm_n, s.n, R, s.knorm, e->d_kn2max, e->stream));
        s.norm_n = s.n;
    }
    if (s.img_cap < s.cap) {
        if (s.kimg) { CK(cudaStreamSynchronize(e->stream)); cudaFree(s.kimg); s.kimg = nullptr; }
        const size_t bytes = scl_knn_tc_image_bytes(R, s.cap);
        CK(cudaMalloc(&s.kimg, bytes));
        CK(cudaMemsetAsync(s.kimg, 0, bytes, e->stream));
        s.img_cap = s.cap; s.img_n = 0;
    }
    if (s.img_n < s.n) {
        CK(scl_launch_key_image(s.keys, s.knorm, s.img_n, s.n, R, s.kimg, e->stream));
        s.img_n = s.n;
    }
    return SCL_OK;
}

// kNN of Q device-resident query keys over the selected sets; results (concatenated index, d2) in out_ids / out_d2 (device)
int knn_sets(scl_rowkey* e, const float* qkeys_dev, int Q, const std::vector<KeySet>& sets, int K, int knn_mode, int32_t* out_ids, float* out_d2)
{
    const int R = e->p.rows;
    if (K < 1 || K > 32) FAIL(SCL_ERR_INVALID, "K must be in 1..32");
    if (sets.size() > 16) FAIL(SCL_ERR_UNSUPPORTED, "more than 16 key sets in one query");
    const size_t QK = (size_t)Q * K;
    const int n_sets = (int)sets.size();
    if (n_sets > 1) { CK(e->blk_ids.ensure(QK * 4 * n_sets)); CK(e->blk_d2.ensure(QK * 4 * n_sets)); }
    int max_n = 0;
    for (const KeySet& s : sets) if (s.n > max_n) max_n = s.n;
    const int splits = scl_knn_splits(Q, max_n);
    CK(e->part_ids.ensure((size_t)Q * splits * K * 4));
    CK(e->part_d2.ensure((size_t)Q * splits * K * 4));
    {
        const void* old = e->tickets.p;
        CK(e->tickets.ensure(((size_t)Q / 128 + 16) * 4));
        if (old != e->tickets.p) CK(cudaMemsetAsync(e->tickets.p, 0, e->tickets.cap, e->stream));
    }
    KnnWorkspace ws{e->part_ids.as<int32_t>(), e->part_d2.as<float>(), e->tickets.as<int>(), (size_t)Q * splits * K};
    for (int i = 0; i < n_sets; i++) {
        const KeySet& ks = sets[i];
        KeyStore& st = e->robots[ks.robot];
        int32_t* ids = n_sets > 1 ? e->blk_ids.as<int32_t>() + (size_t)i * QK : out_ids;
        float* d2 = n_sets > 1 ? e->blk_d2.as<float>() + (size_t)i * QK : out_d2;
        /* the same rule as the Scan Context engine (engine.cu, knn_dev): the tensor-core prefilter from four queries up on a
         * key set worth streaming; it needs K <= K' - 2 */
        const bool use_tc = scl_knn_tc_supported(R) && K <= scl_knn_tc_kprime() - 2 && (knn_mode == 2 || (knn_mode == 0 && Q > 3 && ks.n >= 16384));
        if (use_tc) {
            { int rc = sync_image(e, st); if (rc) return rc; }
            const int Qc = Q < scl_knn_tc_max_batch() ? Q : scl_knn_tc_max_batch();
            const int ranges = scl_knn_tc_ranges(Qc);
            const size_t pairs = (size_t)Qc * ranges;
            CK(e->tc_queues.ensure(pairs * scl_knn_tc_queue_bytes())); CK(e->tc_queue_cnt.ensure(pairs * 4));
            const void* old_slots = e->tc_slots.p; const void* old_cnt = e->tc_fail_count.p;
            CK(e->tc_fail_list.ensure((size_t)Q * 4)); CK(e->tc_fail_count.ensure(128));
            CK(e->tc_slots.ensure((size_t)Qc * scl_knn_tc_slot_stride() * 4));
            const bool init_state = !e->tc_state_clean || old_slots != e->tc_slots.p || old_cnt != e->tc_fail_count.p || Qc > e->tc_slots_rows;
            if (old_cnt != e->tc_fail_count.p) CK(cudaMemsetAsync(e->tc_fail_count.p, 0, 128, e->stream));
            int* fail_cur = e->tc_fail_count.as<int>() + 16 * (e->tc_calls & 1);
            int* fail_next = e->tc_fail_count.as<int>() + 16 * ((e->tc_calls + 1) & 1);
            e->tc_calls++;
            e->tc_state_clean = false;
            KnnTcWorkspace tw{e->tc_queues.as<uint32_t>(), e->tc_queue_cnt.as<int>(), e->tc_slots.as<int>(), nullptr, pairs};
            CK(scl_launch_knn_tc(qkeys_dev, Q, st.keys, st.kimg, e->d_kn2max, ks.n, R, K, 1, 1, ks.offset, tw, ids, d2,
                                 e->tc_fail_list.as<int32_t>(), fail_cur, fail_next, init_state, 3, e->stream));
            e->tc_state_clean = true; e->tc_slots_rows = Qc;
            CK(scl_launch_knn_exact(qkeys_dev, Q, st.keys, ks.n, R, K, 1, 1, ks.offset, e->tc_fail_list.as<int32_t>(), fail_cur, ws, ids, d2, e->stream));
            e->stat_tc += Q;
        } else {
            CK(scl_launch_knn_exact(qkeys_dev, Q, st.keys, ks.n, R, K, 1, 1, ks.offset, nullptr, nullptr, ws, ids, d2, e->stream));
        }
    }
    if (n_sets > 1)
        CK(scl_launch_merge_topk(n_sets, Q, K, e->blk_ids.p, e->blk_d2.p, QK * 4, out_ids, out_d2, e->stream));
    return SCL_OK;
}

// libnabo reports a missing neighbour as index -1 / distance +inf; the exact nanoflann-style kernels as FLT_MAX
void fix_missing(int n, int32_t* ids, float* d2)
{
    for (int i = 0; i < n; i++) if (ids[i] < 0) { ids[i] = -1; d2[i] = INFINITY; }
}

// the candidate list of one stored key against the selected sets, on the host
int one_query(scl_rowkey* e, int q_robot, int q_local, int from_robot, int n_limit, std::vector<int32_t>& idx, std::vector<float>& d2, int* total)
{
    const int K = e->p.num_candidates, R = e->p.rows;
    std::vector<KeySet> sets;
    select_sets(e, from_robot, n_limit, sets, total);
    idx.assign(K, -1); d2.assign(K, INFINITY);
    if (sets.empty()) return SCL_OK;
    CK(e->out_ids.ensure((size_t)K * 4)); CK(e->out_d2.ensure((size_t)K * 4));
    const float* q = e->robots[q_robot].keys + (size_t)q_local * R;
    { int rc = knn_sets(e, q, 1, sets, K, 0, e->out_ids.as<int32_t>(), e->out_d2.as<float>()); if (rc) return rc; }
    CK(cudaMemcpyAsync(idx.data(), e->out_ids.p, (size_t)K * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(d2.data(), e->out_d2.p, (size_t)K * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    fix_missing(K, idx.data(), d2.data());
    return SCL_OK;
}

int intra_list(scl_rowkey* e, int cur_ptr, int* n, std::vector<int32_t>& idx, std::vector<float>& d2)
{
    const int K = e->p.num_candidates;
    *n = 0;
    if (cur_ptr < 0 || cur_ptr >= e->robots[e->p.this_id].n) FAIL(SCL_ERR_RANGE, "cur_ptr is not a key of this robot");
    if (cur_ptr < e->p.num_exclude_recent + K + 1) return SCL_OK;                 /* descriptor.h:1094-1097 */
    int total = 0;
    { int rc = one_query(e, e->p.this_id, cur_ptr, -1, cur_ptr - e->p.num_exclude_recent, idx, d2, &total); if (rc) return rc; }
    *n = K;
    return SCL_OK;
}

int inter_list(scl_rowkey* e, int cur_ptr, int* n, std::vector<int32_t>& idx, std::vector<float>& d2, std::vector<KeySet>& sets)
{
    const int K = e->p.num_candidates;
    *n = 0;
    if (cur_ptr < 0 || cur_ptr >= (int)e->index.size()) FAIL(SCL_ERR_RANGE, "cur_ptr is not a global key");
    const int cur_robot = e->index[cur_ptr].first, cur_index = e->index[cur_ptr].second;    /* descriptor.h:1153-1155 */
    if (cur_robot < 0 || cur_robot >= e->p.robot_num || cur_index < 0 || cur_index >= e->robots[cur_robot].n)
        FAIL(SCL_ERR_RANGE, "the index saved with this key is not a position in its robot's key matrix");
    int total = 0;
    select_sets(e, cur_robot, 0, sets, &total);
    if (total < K + 1) return SCL_OK;                                               /* descriptor.h:1194-1197 */
    { int rc = one_query(e, cur_robot, cur_index, cur_robot, 0, idx, d2, &total); if (rc) return rc; }
    *n = K;
    return SCL_OK;
}

// concatenated position -> (robot, local position)
bool locate(const std::vector<KeySet>& sets, int c, int8_t* robot, int* local)
{
    for (const KeySet& s : sets) if (c >= s.offset && c < s.offset + s.n) { *robot = (int8_t)s.robot; *local = c - s.offset; return true; }
    return false;
}

}  // namespace

extern "C" {

void scl_rowkey_default_params(scl_rowkey_params* p)
{
    if (!p) return;
    p->rows = 80; p->num_exclude_recent = 30; p->num_candidates = 10; p->dist_thres = 0.32; p->robot_num = 1; p->this_id = 0;   /* descriptor.h:472-485 */
}

int scl_rowkey_create(const scl_rowkey_params* p, int device, scl_rowkey** out)
{
    if (!p || !out) return SCL_ERR_INVALID;
    *out = nullptr;
    if (p->rows < 4 || p->rows % 4 != 0 || !(p->rows == 10 || p->rows == 20 || p->rows == 40 || p->rows == 80)) return SCL_ERR_UNSUPPORTED;   /* the K3 kernels' key lengths */
    if (p->robot_num < 1 || p->robot_num > 17 || p->this_id < 0 || p->this_id >= p->robot_num) return SCL_ERR_INVALID;
    if (p->num_candidates < 1 || p->num_candidates > 32 || p->num_exclude_recent < 0) return SCL_ERR_INVALID;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return SCL_ERR_CUDA;     /* no CPU fallback */
    if (cudaSetDevice(device) != cudaSuccess) return SCL_ERR_CUDA;
    scl_rowkey* e = new scl_rowkey();
    e->p = *p; e->device = device;
    e->robots.resize(p->robot_num); e->local2global.resize(p->robot_num);
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc(&e->d_kn2max, 4) != cudaSuccess ||
        cudaMemsetAsync(e->d_kn2max, 0, 4, e->stream) != cudaSuccess) {
        if (e->stream) cudaStreamDestroy(e->stream);
        delete e;
        return SCL_ERR_CUDA;
    }
    scl_preload_k3();
    if (scl_knn_tc_supported(p->rows)) { if (p->rows == 80) scl_preload_k3_tc80(); else scl_preload_k3_tc(); }
    SCL_TOUCH(rowkey_norm_kernel); SCL_TOUCH(rowkey_map_kernel);
    *out = e;
    return SCL_OK;
}

int scl_rowkey_destroy(scl_rowkey* e)
{
    if (!e) return SCL_OK;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (KeyStore& s : e->robots) { if (s.keys) cudaFree(s.keys); if (s.knorm) cudaFree(s.knorm); if (s.kimg) cudaFree(s.kimg); }
    for (DevBuf* b : e->all) b->release();
    if (e->d_kn2max) cudaFree(e->d_kn2max);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return SCL_OK;
}

const char* scl_rowkey_last_error(scl_rowkey* e) { return e ? e->err.c_str() : "null handle"; }

int scl_rowkey_set_stream(scl_rowkey* e, void* cuda_stream)
{
    if (!e) return SCL_ERR_INVALID;
    RK_LOCK();
    CK(cudaStreamSynchronize(e->stream));
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    e->stream = static_cast<cudaStream_t>(cuda_stream); e->own_stream = false;
    return SCL_OK;
}

int scl_rowkey_save_batch(scl_rowkey* e, const float* row_keys, int n, int8_t robot, const int* index)
{
    if (!e) return SCL_ERR_INVALID;
    RK_LOCK();
    if (!row_keys || n < 0) FAIL(SCL_ERR_INVALID, "row_keys is NULL or n < 0");
    if (robot < 0 || robot >= e->p.robot_num) FAIL(SCL_ERR_RANGE, "robot out of range");
    if (n == 0) return SCL_OK;
    const int before = e->robots[robot].n;
    { int rc = append(e, robot, row_keys, n); if (rc) return rc; }
    for (int i = 0; i < n; i++) {                                 /* descriptor.h:1056-1059 */
        e->local2global[robot].push_back((int)e->index.size());
        e->index.push_back(std::make_pair(robot, index ? index[i] : before + i));
    }
    e->l2g_from = -2;
    return SCL_OK;
}

int scl_rowkey_save(scl_rowkey* e, const float* row_key, int8_t robot, int index, int* global_key)
{
    const int rc = scl_rowkey_save_batch(e, row_key, 1, robot, &index);
    if (rc == SCL_OK && global_key) { std::lock_guard<std::mutex> lk(e->mu); *global_key = (int)e->index.size() - 1; }
    return rc;
}

int scl_rowkey_save_wire(scl_rowkey* e, const float* wire, int cols, int8_t robot, int index, int* global_key)
{
    if (!e) return SCL_ERR_INVALID;
    if (!wire || cols < 1) { std::lock_guard<std::mutex> lk(e->mu); FAIL(SCL_ERR_INVALID, "wire is NULL or cols < 1"); }
    return scl_rowkey_save(e, wire + (size_t)e->p.rows * cols, robot, index, global_key);      /* descriptor.h:1038-1041 */
}

int scl_rowkey_intra_candidates(scl_rowkey* e, int cur_ptr, int* n, int32_t* local_idx, float* d2)
{
    if (!e) return SCL_ERR_INVALID;
    RK_LOCK();
    if (!n || !local_idx || !d2) FAIL(SCL_ERR_INVALID, "NULL output");
    std::vector<int32_t> idx; std::vector<float> dd;
    { int rc = intra_list(e, cur_ptr, n, idx, dd); if (rc) return rc; }
    for (int i = 0; i < *n; i++) { local_idx[i] = idx[i]; d2[i] = dd[i]; }
    return SCL_OK;
}

int scl_rowkey_inter_candidates(scl_rowkey* e, int cur_ptr, int* n, int32_t* concat_idx, int32_t* global_key, float* d2)
{
    if (!e) return SCL_ERR_INVALID;
    RK_LOCK();
    if (!n || !concat_idx || !d2) FAIL(SCL_ERR_INVALID, "NULL output");
    std::vector<int32_t> idx; std::vector<float> dd; std::vector<KeySet> sets;
    { int rc = inter_list(e, cur_ptr, n, idx, dd, sets); if (rc) return rc; }
    for (int i = 0; i < *n; i++) {
        concat_idx[i] = idx[i]; d2[i] = dd[i];
        if (global_key) {
            int8_t r = -1; int l = -1;
            global_key[i] = locate(sets, idx[i], &r, &l) ? e->local2global[r][l] : -1;
        }
    }
    return SCL_OK;
}

int scl_rowkey_detect_intra(scl_rowkey* e, int cur_ptr, scl_rowkey_compare_fn cmp, void* user, int* id, float* bias, float* min_dist)
{
    if (!e) return SCL_ERR_INVALID;
    RK_LOCK();
    if (!cmp || !id || !bias) FAIL(SCL_ERR_INVALID, "NULL callback or output");
    *id = -1; *bias = 0.0f;
    if (min_dist) *min_dist = 10000000.0f;
    int n = 0;
    std::vector<int32_t> idx; std::vector<float> dd;
    { int rc = intra_list(e, cur_ptr, &n, idx, dd); if (rc) return rc; }
    if (n == 0) return SCL_OK;
    float min_dis = 10000000.0f; int min_index = -1, min_bias = 0;          /* descriptor.h:1109-1111 */
    const size_t own = e->local2global[e->p.this_id].size();
    for (int i = 0; i < n; i++) {
        if ((size_t)idx[i] >= own) continue;                                /* :1119, -1 included (it converts to a huge unsigned there too) */
        int b = 0;
        const float dis = cmp(user, (int8_t)e->p.this_id, cur_ptr, (int8_t)e->p.this_id, idx[i], &b);
        if (dis < min_dis) { min_dis = dis; min_index = idx[i]; min_bias = b; }
    }
    if (min_dist) *min_dist = min_dis;
    if (min_dis < e->p.dist_thres) { *id = min_index; *bias = (float)min_bias; }   /* :1138-1142 */
    return SCL_OK;
}

int scl_rowkey_detect_inter(scl_rowkey* e, int cur_ptr, scl_rowkey_compare_fn cmp, void* user, int* id, float* bias, float* min_dist)
{
    if (!e) return SCL_ERR_INVALID;
    RK_LOCK();
    if (!cmp || !id || !bias) FAIL(SCL_ERR_INVALID, "NULL callback or output");
    *id = -1; *bias = 0.0f;
    if (min_dist) *min_dist = 10000000.0f;
    int n = 0;
    std::vector<int32_t> idx; std::vector<float> dd; std::vector<KeySet> sets;
    { int rc = inter_list(e, cur_ptr, &n, idx, dd, sets); if (rc) return rc; }
    if (n == 0) return SCL_OK;
    const int cur_robot = e->index[cur_ptr].first, cur_index = e->index[cur_ptr].second;
    float min_dis = 10000000.0f; int min_index = -1, min_bias = 0;
    for (int i = 0; i < n; i++) {
        int8_t r = -1; int l = -1;
        if (!locate(sets, idx[i], &r, &l)) continue;                        /* :1214-1218 */
        int b = 0;
        const float dis = cmp(user, (int8_t)cur_robot, cur_index, r, l, &b);
        if (dis < min_dis) { min_dis = dis; min_index = e->local2global[r][l]; min_bias = b; }   /* :1226-1231 */
    }
    if (min_dist) *min_dist = min_dis;
    if (min_dis < e->p.dist_thres) { *id = min_index; *bias = (float)min_bias; }   /* :1235-1238 */
    return SCL_OK;
}

int scl_rowkey_get_index(scl_rowkey* e, int key, int8_t* robot, int* index)
{
    if (!e) return SCL_ERR_INVALID;
    RK_LOCK();
    const bool ok = key >= 0 && key < (int)e->index.size();
    if (robot) *robot = ok ? e->index[key].first : (int8_t)-1;
    if (index) *index = ok ? e->index[key].second : -1;
    return SCL_OK;
}

int scl_rowkey_size(scl_rowkey* e, int id_in)
{
    if (!e) return 0;
    RK_LOCK();
    if (id_in == -1) return (int)e->index.size();                            /* descriptor.h:1259-1262 */
    if (id_in < 0 || id_in >= e->p.robot_num) return 0;
    return (int)e->local2global[id_in].size();
}

int scl_rowkey_knn_batch_dev(scl_rowkey* e, const float* q_keys_dev, int Q, int from_robot, int n_limit, int K, int knn_mode,
                             int32_t* concat_idx_dev, float* d2_dev)
{
    if (!e) return SCL_ERR_INVALID;
    RK_LOCK();
    if (Q <= 0) return SCL_OK;
    if (!q_keys_dev || !concat_idx_dev || !d2_dev) FAIL(SCL_ERR_INVALID, "NULL pointer");
    if (from_robot < -1 || from_robot >= e->p.robot_num) FAIL(SCL_ERR_RANGE, "from_robot out of range");
    if (knn_mode < 0 || knn_mode > 2) FAIL(SCL_ERR_INVALID, "knn_mode must be 0, 1 or 2");
    std::vector<KeySet> sets; int total = 0;
    select_sets(e, from_robot, n_limit, sets, &total);
    if (sets.empty()) {
        CK(cudaMemsetAsync(concat_idx_dev, 0xff, (size_t)Q * K * 4, e->stream));
        CK(cudaMemsetAsync(d2_dev, 0x7f, (size_t)Q * K * 4, e->stream));
        return SCL_OK;
    }
    return knn_sets(e, q_keys_dev, Q, sets, K, knn_mode, concat_idx_dev, d2_dev);
}

int scl_rowkey_knn_batch(scl_rowkey* e, const float* q_keys, int Q, int from_robot, int n_limit, int K, int knn_mode,
                         int32_t* concat_idx, int32_t* global_key, float* d2)
{
    if (!e) return SCL_ERR_INVALID;
    {
        RK_LOCK();
        if (Q <= 0) return SCL_OK;
        if (!q_keys || !concat_idx || !d2) FAIL(SCL_ERR_INVALID, "NULL pointer");
        if (K < 1 || K > 32) FAIL(SCL_ERR_INVALID, "K must be in 1..32");
        CK(e->qkeys.ensure((size_t)Q * e->p.rows * 4));
        CK(e->out_ids.ensure((size_t)Q * K * 4)); CK(e->out_d2.ensure((size_t)Q * K * 4));
        CK(cudaMemcpyAsync(e->qkeys.p, q_keys, (size_t)Q * e->p.rows * 4, cudaMemcpyHostToDevice, e->stream));
    }
    { int rc = scl_rowkey_knn_batch_dev(e, e->qkeys.as<float>(), Q, from_robot, n_limit, K, knn_mode, e->out_ids.as<int32_t>(), e->out_d2.as<float>()); if (rc) return rc; }
    RK_LOCK();
    const size_t QK = (size_t)Q * K;
    if (global_key) {
        /* newLocal2Global of the selected sets (descriptor.h:1175,1188), applied on the device */
        std::vector<KeySet> sets; int total = 0;
        select_sets(e, from_robot, n_limit, sets, &total);
        std::vector<int> map((size_t)total);
        for (const KeySet& s : sets) for (int i = 0; i < s.n; i++) map[(size_t)s.offset + i] = e->local2global[s.robot][i];
        CK(e->l2g.ensure((size_t)(total > 0 ? total : 1) * 4)); CK(e->out_gk.ensure(QK * 4));
        if (total > 0) CK(cudaMemcpyAsync(e->l2g.p, map.data(), (size_t)total * 4, cudaMemcpyHostToDevice, e->stream));
        rowkey_map_kernel<<<(int)((QK + 255) / 256), 256, 0, e->stream>>>(e->out_ids.as<int32_t>(), (int)QK, e->l2g.as<int32_t>(), total, e->out_gk.as<int32_t>());
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(global_key, e->out_gk.p, QK * 4, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));                       /* `map` is pageable and on this stack */
    }
    CK(cudaMemcpyAsync(concat_idx, e->out_ids.p, QK * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(d2, e->out_d2.p, QK * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    fix_missing((int)QK, concat_idx, d2);
    return SCL_OK;
}

int scl_rowkey_knn_stats(scl_rowkey* e, long long* tc_queries, long long* fallback_queries)
{
    if (!e) return SCL_ERR_INVALID;
    RK_LOCK();
    if (tc_queries) *tc_queries = e->stat_tc;
    if (fallback_queries) {
        *fallback_queries = 0;
        if (e->tc_fail_count.p) {
            int h[32];
            CK(cudaStreamSynchronize(e->stream));
            CK(cudaMemcpy(h, e->tc_fail_count.p, sizeof(h), cudaMemcpyDeviceToHost));
            *fallback_queries = (long long)h[8] + h[24];           /* the re-rank kernel's running totals, one per call parity */
        }
    }
    return SCL_OK;
}

}  // extern "C"
