/*
 * ref_driver.cpp — builds oracle/_ref/libscl_ref.so: the reference's OWN scan_context_descriptor
 * class text and its vendored nanoflann, compiled where they lie under /root/reference, behind
 * the same sco_* C entry points as the restatement (oracle/sc_oracle.h) so one test harness
 * drives both. TEST INFRASTRUCTURE ONLY; see oracle/Makefile (target _ref/libscl_ref.so) for the recipe.
 *
 * The two .inc files are cut from /root/reference/include/descriptor.h at build time into a
 * temporary directory (never into the repo): class scan_descriptor (:21-36) and class
 * scan_context_descriptor (:1304-1801). Eigen/PCL/ROS/libnabo are replaced by
 * oracle/ref_shim/ref_standins.h (our code). nanoflann.hpp and KDTreeVectorOfVectorsAdaptor.h
 * are the reference's vendored files, included unchanged.
 *
 * Fix-ups applied from OUTSIDE the class (its members are public), so that the shipped-broken
 * nanoflann path can run at all:
 *   - after every insert, polarcontext_invkeys_mat_ gets the float ring key (the push the
 *     reference has commented out at descriptor.h:1596);
 *   - PC_UNIT_SECTORANGLE, PC_UNIT_RINGGAP, tree_making_period_conter are initialised
 *     (the constructor shadows them, descriptor.h:1332-1334).
 */
#include "ref_standins.h"

#include "KDTreeVectorOfVectorsAdaptor.h" /* /root/reference/include, pulls nanoflann.hpp */

#include <cstring>
#include <deque>
#include <thread>

using namespace std; /* descriptor.h:19 — selects the float overloads of sqrt/atan in the class */

#include "scan_descriptor.inc"
#include "scan_context_descriptor.inc"

#include "../sc_oracle.h"

typedef KDTreeVectorOfVectorsAdaptor<std::vector<std::vector<float>>, float> ref_tree_t;

struct sco_handle {
    scan_context_descriptor* sc;
    int R, S;
    /* cached tree for sco_knn / sco_query_batch */
    std::vector<std::vector<float>> tree_keys;
    std::unique_ptr<ref_tree_t> tree;
    int tree_n;
    /* bulk mode (sco_bulk_load): descriptors live here as float wires instead of in
     * sc->polarcontexts_, because the class's own insert is O(N) per call (descriptor.h:1597) */
    std::vector<const float*> wires;
    std::deque<std::vector<float>> owned;
};

static void after_insert(sco_handle* h)
{
    scan_context_descriptor* sc = h->sc;
    const int n = (int)sc->polarcontexts_.size();
    std::vector<float> key(h->R);
    for (int r = 0; r < h->R; r++) key[r] = sc->polarcontextRowKey(r, n - 1);
    sc->polarcontext_invkeys_mat_.push_back(key);
}

static pcl::PointCloud<pcl::PointXYZI> to_cloud(const float* pts, int n, int stride)
{
    pcl::PointCloud<pcl::PointXYZI> c;
    c.points.resize(n);
    for (int i = 0; i < n; i++) {
        c.points[i].x = pts[(size_t)i * stride]; c.points[i].y = pts[(size_t)i * stride + 1];
        c.points[i].z = pts[(size_t)i * stride + 2]; c.points[i].intensity = 0;
    }
    return c;
}

static Eigen::MatrixXd to_mat(sco_handle* h, const float* d)
{
    Eigen::MatrixXd m(h->R, h->S);
    for (int r = 0; r < h->R; r++) for (int c = 0; c < h->S; c++) m(r, c) = d[(size_t)r * h->S + c];
    return m;
}

static Eigen::MatrixXd get_mat(sco_handle* h, int id)
{
    if (!h->wires.empty()) return to_mat(h, h->wires[id]);
    return h->sc->polarcontexts_[id];
}

static void ensure_tree(sco_handle* h, int n_db)
{
    if (h->tree && h->tree_n == n_db) return;
    h->tree.reset();
    h->tree_keys.assign(h->sc->polarcontext_invkeys_mat_.begin(), h->sc->polarcontext_invkeys_mat_.begin() + n_db);
    h->tree = std::make_unique<ref_tree_t>(h->R, h->tree_keys, 10 /* max leaf, descriptor.h:1699 */);
    h->tree_n = n_db;
}

extern "C" {

sco_handle* sco_create(int R, int S, int K, double thr, double lidar_h, double max_r, int excl, int period, double ratio)
{
    sco_handle* h = new sco_handle();
    h->sc = new scan_context_descriptor(R, S, K, thr, lidar_h, max_r, excl, period, ratio);
    h->sc->PC_UNIT_SECTORANGLE = 360.0 / double(S);
    h->sc->PC_UNIT_RINGGAP = max_r / double(R);
    h->sc->tree_making_period_conter = 0;
    h->R = R; h->S = S; h->tree_n = -1;
    return h;
}
void sco_destroy(sco_handle* h) { delete h->sc; delete h; }

void sco_make_scancontext(sco_handle* h, const float* pts, int n, int stride, float* out_desc, int* out_ring, int* out_sector)
{
    (void)out_ring; (void)out_sector; /* the reference does not expose per-point bins */
    std::vector<float> vT;
    h->sc->makeScancontext(to_cloud(pts, n, stride), &vT);
    if (out_desc) std::memcpy(out_desc, vT.data(), sizeof(float) * vT.size());
}

int sco_make_and_save(sco_handle* h, const float* pts, int n, int stride, int8_t robot, int index, float* out_desc)
{
    std::vector<float> vT = h->sc->makeAndSaveDescriptorAndKey(to_cloud(pts, n, stride), robot, index);
    after_insert(h);
    if (out_desc) std::memcpy(out_desc, vT.data(), sizeof(float) * vT.size());
    return h->sc->getSize() - 1;
}

int sco_save(sco_handle* h, const float* wire, int8_t robot, int index)
{
    h->sc->saveDescriptorAndKey(wire, robot, index);
    after_insert(h);
    return h->sc->getSize() - 1;
}

int sco_bulk_load(sco_handle* h, const float* wires, int n, int borrow)
{
    if (h->sc->getSize() != 0 && h->wires.empty()) return -1; /* bulk mode only on a fresh handle */
    const size_t rs = (size_t)h->R * h->S;
    const int base = (int)h->wires.size();
    h->wires.resize(base + n);
    h->sc->polarcontext_invkeys_mat_.resize(base + n);
    h->sc->polarcontext_indexs_.resize(base + n);
    std::vector<std::vector<float>> own(borrow ? 0 : n);
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    auto work = [&](int i0, int i1) {
        for (int i = i0; i < i1; i++) {
            const float* w = wires + (size_t)i * rs;
            if (!borrow) { own[i].assign(w, w + rs); w = own[i].data(); }
            h->wires[base + i] = w;
            const Eigen::MatrixXf key = h->sc->makeRingkeyFromScancontext(to_mat(h, w)); /* descriptor.h:1589 */
            std::vector<float> kv(h->R);
            for (int r = 0; r < h->R; r++) kv[r] = key(r, 0);
            h->sc->polarcontext_invkeys_mat_[base + i] = kv;
            h->sc->polarcontext_indexs_[base + i] = std::make_pair((int8_t)0, base + i);
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < hw; t++) th.emplace_back(work, (int)((long long)n * t / hw), (int)((long long)n * (t + 1) / hw));
    for (auto& t : th) t.join();
    for (auto& v : own) h->owned.push_back(std::move(v));
    return base + n;
}

int sco_size(sco_handle* h) { return h->sc->getSize(); }

void sco_get_index(sco_handle* h, int key, int* robot, int* index)
{
    if (key < 0 || key >= h->sc->getSize()) { *robot = -1; *index = -1; return; } /* the reference is UB here */
    const std::pair<int8_t, int> p = h->sc->getIndex(key);
    *robot = p.first; *index = p.second;
}

void sco_get_desc(sco_handle* h, int key, float* out)
{
    const Eigen::MatrixXd m = get_mat(h, key);
    for (int r = 0; r < h->R; r++) for (int c = 0; c < h->S; c++) out[(size_t)r * h->S + c] = (float)m(r, c);
}

void sco_ring_key(sco_handle* h, int key, float* out)
{
    for (int r = 0; r < h->R; r++) out[r] = h->sc->polarcontext_invkeys_mat_[key][r];
}

void sco_sector_key(sco_handle* h, int key, double* out)
{
    const Eigen::MatrixXd v = h->sc->makeSectorkeyFromScancontext(get_mat(h, key));
    for (int c = 0; c < h->S; c++) out[c] = v(0, c);
}

void sco_distance(sco_handle* h, int k1, int k2, double* dist, int* shift)
{
    const std::pair<double, int> r = h->sc->distanceBtnScanContext(get_mat(h, k1), get_mat(h, k2));
    *dist = r.first; *shift = r.second;
}

void sco_distance_raw(sco_handle* h, const float* d1, const float* d2, double* dist, int* shift)
{
    const std::pair<double, int> r = h->sc->distanceBtnScanContext(to_mat(h, d1), to_mat(h, d2));
    *dist = r.first; *shift = r.second;
}

int sco_fast_align(sco_handle* h, int k1, int k2)
{
    return h->sc->fastAlignUsingVkey(h->sc->makeSectorkeyFromScancontext(get_mat(h, k1)),
                                     h->sc->makeSectorkeyFromScancontext(get_mat(h, k2)));
}

double sco_dist_direct(sco_handle* h, int k1, int k2, int shift)
{
    return h->sc->distDirectSC(get_mat(h, k1), h->sc->circshift(get_mat(h, k2), shift));
}

void sco_detect_intra(sco_handle* h, int cur, int* id, float* second)
{
    const std::pair<int, float> r = h->sc->detectIntraLoopClosureID(cur);
    delete h->sc->kdTree; h->sc->kdTree = NULL; /* the reference leaks one tree per call (descriptor.h:1631) */
    *id = r.first; *second = r.second;
}

void sco_detect_inter(sco_handle* h, int cur, int* id, float* second)
{
    const std::pair<int, float> r = h->sc->detectInterLoopClosureID(cur);
    *id = r.first; *second = r.second;
}

int sco_knn(sco_handle* h, int cur, int n_db, int k, int metric, int32_t* ids, float* d2)
{
    for (int i = 0; i < k; i++) { ids[i] = -1; d2[i] = FLT_MAX; }
    if (n_db <= 0) return 0;
    if (metric == 1) {
        Eigen::MatrixXf keys(h->R, n_db);
        for (int j = 0; j < n_db; j++) for (int r = 0; r < h->R; r++) keys(r, j) = h->sc->polarcontext_invkeys_mat_[j][r];
        std::unique_ptr<Nabo::NNSearchF> t(Nabo::NNSearchF::createKDTreeLinearHeap(keys, h->R));
        Eigen::VectorXf q(h->R);
        for (int r = 0; r < h->R; r++) q[r] = h->sc->polarcontext_invkeys_mat_[cur][r];
        Eigen::VectorXi idx(k); Eigen::VectorXf dd(k);
        t->knn(q, idx, dd, k);
        int found = 0;
        for (int i = 0; i < k; i++) if (idx[i] >= 0) { ids[found] = idx[i]; d2[found] = dd[i]; found++; }
        return found;
    }
    ensure_tree(h, n_db);
    std::vector<size_t> ci(k); std::vector<float> cd(k);
    nanoflann::KNNResultSet<float> rs(k);
    rs.init(&ci[0], &cd[0]);
    h->tree->index->findNeighbors(rs, h->sc->polarcontext_invkeys_mat_[cur].data(), nanoflann::SearchParams(10));
    const int found = (int)rs.size();
    for (int i = 0; i < found; i++) { ids[i] = (int32_t)ci[i]; d2[i] = cd[i]; }
    return found;
}

void sco_query_batch(sco_handle* h, const int32_t* queries, int nq, int n_db, int k, int metric, int nthreads,
                     int32_t* cand_ids, float* cand_d2, double* cand_dist, int32_t* cand_shift,
                     int32_t* best_id, double* best_dist, int32_t* best_shift)
{
    (void)metric; /* nanoflann flavour only: the loop below is descriptor.h:1705-1737 per query */
    ensure_tree(h, n_db);
    scan_context_descriptor* sc = h->sc;
    auto work = [&](int t0, int t1) {
        for (int qi = t0; qi < t1; qi++) {
            const int cur = queries[qi];
            std::vector<size_t> ci(k); std::vector<float> cd(k);
            nanoflann::KNNResultSet<float> rs(k);
            rs.init(&ci[0], &cd[0]);
            h->tree->index->findNeighbors(rs, sc->polarcontext_invkeys_mat_[cur].data(), nanoflann::SearchParams(10));
            const int found = (int)rs.size();
            double min_dist = 10000000; int nn_align = 0, nn_idx = -1;
            for (int i = 0; i < k; i++) {
                int32_t id = -1; float d2 = FLT_MAX; double dist = NAN; int shift = 0;
                if (i < found) {
                    id = (int32_t)ci[i]; d2 = cd[i];
                    const std::pair<double, int> r = sc->distanceBtnScanContext(get_mat(h, cur), get_mat(h, id));
                    dist = r.first; shift = r.second;
                    if (dist < min_dist && id != cur) { min_dist = dist; nn_align = shift; nn_idx = id; }
                }
                cand_ids[(size_t)qi * k + i] = id; cand_d2[(size_t)qi * k + i] = d2;
                cand_dist[(size_t)qi * k + i] = dist; cand_shift[(size_t)qi * k + i] = shift;
            }
            best_id[qi] = nn_idx; best_dist[qi] = min_dist; best_shift[qi] = nn_align;
        }
    };
    if (nthreads <= 1) { work(0, nq); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back(work, (int)((long long)nq * t / nthreads), (int)((long long)nq * (t + 1) / nthreads));
    for (auto& t : th) t.join();
}

float sco_atanf_libm(float x) { return atanf(x); }
float sco_atanf_port(float x) { return atanf(x); }

} // extern "C"
