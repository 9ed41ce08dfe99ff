/*
 * sc_oracle.cpp — CPU ORACLE (test infrastructure, not product) for the Scan Context path.
 *
 * A dependency-free restatement of class scan_context_descriptor,
 * /root/reference/include/descriptor.h:1304-1801. Eigen expressions are restated as plain
 * sequential loops: mean() = (sum in index order)/size, norm() = sqrt(sum of squares in index
 * order), dot() = sum of products in index order. Doubles where the reference uses double,
 * floats where it uses float. Built with -ffp-contract=off so no FMA is ever formed (the
 * reference builds for baseline x86-64, which has none).
 *
 * Documented deviations from the reference as shipped:
 *  Q1  PC_UNIT_SECTORANGLE / tree_making_period_conter are uninitialised members there
 *      (descriptor.h:1332-1334 shadow them); here they are 360/S and 0.
 *  Q2  polarcontext_invkeys_mat_ is never filled there (descriptor.h:1596 is commented out) so
 *      the nanoflann path reads an empty vector; here it is filled with the float ring key on
 *      every save, which is what the commented line and upstream Scan Context do.
 *  Q9  points whose x or y is NaN fall off the end of xy2theta (descriptor.h:1352-1374, UB);
 *      here they are dropped. (0,0) gives theta = NaN and int(ceil(NaN)) which is INT_MIN on
 *      x86-64 and therefore sector 1 after the clamp; that is reproduced.
 *  kNN ties: nanoflann keeps the first-visited of equal distances (traversal dependent,
 *      nanoflann.hpp:175-202); libnabo likewise. Here: lowest index first.
 *  libnabo (un-vendored, ethz-asl/libnabo @ 2cc2650) is restated from its published behaviour:
 *      sequential float sum of squared differences, results sorted ascending, candidates with
 *      d2 <= FLT_EPSILON skipped unless ALLOW_SELF_MATCH (the reference passes no flags,
 *      descriptor.h:1642), missing results reported as index -1.
 */
#include "sc_oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <deque>
#include <thread>
#include <utility>
#include <vector>

namespace {

/* ---- the float atan of this libm, restated (fdlibm s_atanf; glibc 2.39 flt-32/s_atanf.c
 * follows it). tests/test_oracle_atanf.py checks it bit for bit against libm. The CUDA
 * polar-binning kernel carries the same sequence of IEEE operations. */
inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

/* The algorithm and constants below are those of fdlibm's s_atanf.c, whose notice is kept as its licence asks:
 * ====================================================
 * Copyright (C) 1993 by Sun Microsystems, Inc. All rights reserved.
 *
 * Developed at SunPro, a Sun Microsystems, Inc. business.
 * Permission to use, copy, modify, and distribute this
 * software is freely granted, provided that this notice
 * is preserved.
 * ====================================================
 */
float atanf_port(float x)
{
    static const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    static const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    static const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                                 9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                                 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
    const uint32_t hx = f2u(x), ix = hx & 0x7fffffffu;
    int id;
    if (ix >= 0x4c000000u) {            /* |x| >= 2^25 */
        if (ix > 0x7f800000u) return x + x;
        const float r = atanhi[3] + atanlo[3];
        return (hx >> 31) ? -r : r;
    }
    if (ix < 0x3ee00000u) {             /* |x| < 0.4375 */
        if (ix < 0x31000000u) return x; /* |x| < 2^-29 */
        id = -1;
    } else {
        x = std::fabs(x);
        if (ix < 0x3f980000u) {         /* |x| < 1.1875 */
            if (ix < 0x3f300000u) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else                  { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000u) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else                  { id = 3; x = -1.0f / x; }
        }
    }
    float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return (hx >> 31) ? -z : z;
}

} // namespace

struct sco_handle {
    /* descriptor.h:1768-1787 */
    double LIDAR_HEIGHT;
    int PC_NUM_RING, PC_NUM_SECTOR;
    double PC_MAX_RADIUS, PC_UNIT_SECTORANGLE, PC_UNIT_RINGGAP;
    int NUM_EXCLUDE_RECENT, NUM_CANDIDATES_FROM_TREE;
    double SEARCH_RATIO, SC_DIST_THRES;
    int TREE_MAKING_PERIOD_, tree_making_period_conter;

    /* descriptor.h:1791-1798. polarcontexts_ holds the descriptors as their row-major float
     * wire image (lossless: every stored value is a float widened to double, :1422,1440,1580);
     * they are widened back to double on use. Entries are owned, or borrowed by sco_bulk_load. */
    std::vector<const float*> polarcontexts_;
    std::deque<std::vector<float>> owned_;
    std::vector<double> widen(const float* w) const { return std::vector<double>(w, w + (size_t)PC_NUM_RING * PC_NUM_SECTOR); }
    std::vector<std::pair<int8_t, int>> polarcontext_indexs_;
    std::vector<std::vector<float>> polarcontext_invkeys_mat_;
    int n_tree; /* number of keys in polarcontext_invkeys_to_search_ at the last rebuild */

    /* descriptor.h:1352-1374 */
    float xy2theta(float x, float y, bool* defined) const
    {
        *defined = true;
        if ((x >= 0) & (y >= 0)) return (float)((180 / M_PI) * atanf(y / x));
        if ((x < 0) & (y >= 0)) return (float)(180 - ((180 / M_PI) * atanf(y / (-x))));
        if ((x < 0) & (y < 0)) return (float)(180 + ((180 / M_PI) * atanf(y / x)));
        if ((x >= 0) & (y < 0)) return (float)(360 - ((180 / M_PI) * atanf((-y) / x)));
        *defined = false; /* NaN coordinate: the reference falls off the end (UB) */
        return 0.0f;
    }

    static int ceil_to_int(double v)
    {
        /* int(ceil(v)) as x86-64 cvttsd2si does it: NaN and out-of-range give INT_MIN */
        const double c = std::ceil(v);
        if (!(c >= -2147483648.0 && c <= 2147483647.0)) return INT_MIN;
        return (int)c;
    }

    /* descriptor.h:1404-1461 */
    void makeScancontext(const float* pts, int n, int stride, std::vector<double>& desc,
                         float* wire, int* out_ring, int* out_sector) const
    {
        const int R = PC_NUM_RING, S = PC_NUM_SECTOR;
        const int NO_POINT = -1000;
        desc.assign((size_t)R * S, (double)NO_POINT);
        for (int i = 0; i < n; i++) {
            const float* p = pts + (size_t)i * stride;
            const float px = p[0], py = p[1];
            const float pz = (float)(p[2] + LIDAR_HEIGHT);            /* :1422 */
            const float azim_range = std::sqrt(px * px + py * py);     /* :1425 */
            bool defined;
            const float azim_angle = xy2theta(px, py, &defined);       /* :1426 */
            if (out_ring) { out_ring[i] = 0; out_sector[i] = 0; }
            if (!defined) continue;                                     /* deviation Q9 */
            if (azim_range > PC_MAX_RADIUS) continue;                   /* :1429 */
            const int ring_idx = std::max(std::min(R, ceil_to_int((azim_range / PC_MAX_RADIUS) * R)), 1);
            const int sctor_idx = std::max(std::min(S, ceil_to_int((azim_angle / 360.0) * S)), 1);
            if (out_ring) { out_ring[i] = ring_idx; out_sector[i] = sctor_idx; }
            double& bin = desc[(size_t)(ring_idx - 1) * S + (sctor_idx - 1)];
            if (bin < pz) bin = pz;                                     /* :1438-1441 */
        }
        for (int r = 0; r < R; r++)
            for (int c = 0; c < S; c++) {
                double& v = desc[(size_t)r * S + c];
                if (v == NO_POINT) v = 0;                               /* :1450-1453 */
                if (wire) wire[(size_t)r * S + c] = (float)v;           /* :1454 */
            }
    }

    /* descriptor.h:1463-1475 */
    void makeRingkey(const std::vector<double>& desc, std::vector<float>& key) const
    {
        const int R = PC_NUM_RING, S = PC_NUM_SECTOR;
        key.resize(R);
        for (int r = 0; r < R; r++) {
            double s = 0;
            for (int c = 0; c < S; c++) s += desc[(size_t)r * S + c];
            key[r] = (float)(s / S);
        }
    }

    /* descriptor.h:1477-1489 */
    void makeSectorkey(const std::vector<double>& desc, std::vector<double>& key) const
    {
        const int R = PC_NUM_RING, S = PC_NUM_SECTOR;
        key.resize(S);
        for (int c = 0; c < S; c++) {
            double s = 0;
            for (int r = 0; r < R; r++) s += desc[(size_t)r * S + c];
            key[c] = s / R;
        }
    }

    /* descriptor.h:1491-1511 with circshift (:1376-1395): shifted[(c+s)%S] = in[c] */
    int fastAlignUsingVkey(const std::vector<double>& v1, const std::vector<double>& v2) const
    {
        const int S = PC_NUM_SECTOR;
        int argmin_vkey_shift = 0;
        double min_veky_diff_norm = 10000000;
        for (int shift = 0; shift < S; shift++) {
            double ss = 0;
            for (int c = 0; c < S; c++) {
                const double d = v1[c] - v2[(c - shift + S) % S];
                ss += d * d;
            }
            const double cur = std::sqrt(ss);
            if (cur < min_veky_diff_norm) { argmin_vkey_shift = shift; min_veky_diff_norm = cur; }
        }
        return argmin_vkey_shift;
    }

    /* descriptor.h:1513-1536 on (sc1, circshift(sc2, shift)) */
    double distDirectSC(const std::vector<double>& a, const std::vector<double>& b, int shift) const
    {
        const int R = PC_NUM_RING, S = PC_NUM_SECTOR;
        int num_eff_cols = 0;
        double sum_sector_similarity = 0;
        for (int c = 0; c < S; c++) {
            const int cb = (c - shift + S) % S;
            double na = 0, nb = 0, dot = 0;
            for (int r = 0; r < R; r++) { const double x = a[(size_t)r * S + c]; na += x * x; }
            for (int r = 0; r < R; r++) { const double y = b[(size_t)r * S + cb]; nb += y * y; }
            na = std::sqrt(na); nb = std::sqrt(nb);
            if ((na == 0) | (nb == 0)) continue;
            for (int r = 0; r < R; r++) dot += a[(size_t)r * S + c] * b[(size_t)r * S + cb];
            sum_sector_similarity = sum_sector_similarity + dot / (na * nb);
            num_eff_cols = num_eff_cols + 1;
        }
        const double sc_sim = sum_sector_similarity / num_eff_cols; /* 0/0 = NaN when no column counts */
        return 1.0 - sc_sim;
    }

    /* descriptor.h:1538-1569 */
    std::pair<double, int> distanceBtnScanContext(const std::vector<double>& sc1, const std::vector<double>& sc2) const
    {
        const int S = PC_NUM_SECTOR;
        std::vector<double> vkey1, vkey2;
        makeSectorkey(sc1, vkey1);
        makeSectorkey(sc2, vkey2);
        const int argmin_vkey_shift = fastAlignUsingVkey(vkey1, vkey2);
        const int SEARCH_RADIUS = (int)std::round(0.5 * SEARCH_RATIO * S);
        std::vector<int> space{argmin_vkey_shift};
        for (int ii = 1; ii < SEARCH_RADIUS + 1; ii++) {
            space.push_back((argmin_vkey_shift + ii + S) % S);
            space.push_back((argmin_vkey_shift - ii + S) % S);
        }
        std::sort(space.begin(), space.end());
        int argmin_shift = 0;
        double min_sc_dist = 10000000;
        for (int num_shift : space) {
            const double cur = distDirectSC(sc1, sc2, num_shift);
            if (cur < min_sc_dist) { argmin_shift = num_shift; min_sc_dist = cur; }
        }
        return std::make_pair(min_sc_dist, argmin_shift);
    }

    /* descriptor.h:1587-1602 */
    int save(const std::vector<double>& sc, int8_t robot, int index)
    {
        std::vector<float> ringkey;
        makeRingkey(sc, ringkey);
        owned_.emplace_back(sc.begin(), sc.end()); /* double -> float is exact here */
        polarcontexts_.push_back(owned_.back().data());
        polarcontext_invkeys_mat_.push_back(ringkey); /* deviation Q2 */
        polarcontext_indexs_.push_back(std::make_pair(robot, index));
        return (int)polarcontexts_.size() - 1;
    }

    /* nanoflann.hpp:383-408 (L2_Adaptor::evalMetric without the early exit, which only
     * truncates sums that are rejected anyway) */
    static float d2_nanoflann(const float* a, const float* b, int dim)
    {
        float result = 0;
        int d = 0;
        for (; d + 3 < dim; d += 4) {
            const float diff0 = a[d] - b[d], diff1 = a[d + 1] - b[d + 1];
            const float diff2 = a[d + 2] - b[d + 2], diff3 = a[d + 3] - b[d + 3];
            result += diff0 * diff0 + diff1 * diff1 + diff2 * diff2 + diff3 * diff3;
        }
        for (; d < dim; d++) { const float diff0 = a[d] - b[d]; result += diff0 * diff0; }
        return result;
    }
    /* libnabo leaf loop: sequential accumulation */
    static float d2_sequential(const float* a, const float* b, int dim)
    {
        float dist = 0;
        for (int d = 0; d < dim; d++) { const float diff = a[d] - b[d]; dist += diff * diff; }
        return dist;
    }

    int knn(const float* q, int n_db, int k, int metric, int32_t* ids, float* d2) const
    {
        const int R = PC_NUM_RING;
        int count = 0;
        for (int i = 0; i < k; i++) { ids[i] = -1; d2[i] = FLT_MAX; }
        for (int j = 0; j < n_db; j++) {
            const float* key = polarcontext_invkeys_mat_[j].data();
            const float dist = metric == 0 ? d2_nanoflann(q, key, R) : d2_sequential(q, key, R);
            if (metric == 1 && !(dist > FLT_EPSILON)) continue; /* libnabo self-match rule */
            /* both trees accept a point only if dist < current worst; before the set is full the
             * worst is FLT_MAX in nanoflann (nanoflann.hpp:163) and +inf in libnabo */
            const float worst = count == k ? d2[k - 1] : (metric == 0 ? FLT_MAX : INFINITY);
            if (!(dist < worst)) continue;                      /* ties with the worst are rejected */
            int i = count < k ? count : k - 1;
            for (; i > 0 && d2[i - 1] > dist; --i) { d2[i] = d2[i - 1]; ids[i] = ids[i - 1]; }
            d2[i] = dist; ids[i] = j;
            if (count < k) count++;
        }
        return count;
    }

    /* descriptor.h:1613-1674 */
    std::pair<int, float> detectIntra(int curPtr) const
    {
        std::pair<int, float> result{-1, 0.0f};
        const int K = NUM_CANDIDATES_FROM_TREE;
        if (curPtr < NUM_EXCLUDE_RECENT + K + 1) return result;     /* :1620 */
        const int historyIndex = curPtr - NUM_EXCLUDE_RECENT;        /* :1627 */
        std::vector<int32_t> indice(K);
        std::vector<float> distance(K);
        knn(polarcontext_invkeys_mat_[curPtr].data(), historyIndex, K, 1, indice.data(), distance.data());
        float minDis = 10000000.0f; /* a float in the reference (:1637) */
        int minIndex = -1, minBias = 0;
        for (int i = 0; i < K; i++) {
            if (indice[i] < 0) continue; /* libnabo reports missing neighbours as -1 */
            const std::pair<double, int> r = distanceBtnScanContext(widen(polarcontexts_[curPtr]), widen(polarcontexts_[indice[i]]));
            if (r.first < minDis) {          /* double < float compare, then narrowed on store */
                minDis = (float)r.first;
                minIndex = indice[i];
                minBias = r.second;
            }
        }
        if (minDis < SC_DIST_THRES) { result.first = minIndex; result.second = (float)minBias; }
        return result;
    }

    /* descriptor.h:1676-1756 */
    std::pair<int, float> detectInter(int currentPtr)
    {
        const int K = NUM_CANDIDATES_FROM_TREE;
        int loop_id = -1;
        if ((int)polarcontext_invkeys_mat_.size() < NUM_EXCLUDE_RECENT + 1) return {loop_id, 0.0f};
        if (tree_making_period_conter % TREE_MAKING_PERIOD_ == 0)
            n_tree = (int)polarcontext_invkeys_mat_.size() - NUM_EXCLUDE_RECENT;   /* :1696 */
        tree_making_period_conter = tree_making_period_conter + 1;
        double min_dist = 10000000;
        int nn_align = 0, nn_idx = -1;
        std::vector<int32_t> ids(K);
        std::vector<float> d2(K);
        const int found = knn(polarcontext_invkeys_mat_[currentPtr].data(), n_tree, K, 0, ids.data(), d2.data());
        for (int i = found; i < K; i++) ids[i] = 0; /* std::vector<size_t>(K) is zero-filled (:1710) */
        for (int i = 0; i < K; i++) {
            const std::pair<double, int> r = distanceBtnScanContext(widen(polarcontexts_[currentPtr]), widen(polarcontexts_[ids[i]]));
            if (r.first < min_dist) {
                if (ids[i] == currentPtr) continue;                  /* :1731 */
                min_dist = r.first; nn_align = r.second; nn_idx = ids[i];
            }
        }
        if (min_dist < SC_DIST_THRES) loop_id = nn_idx;
        const float yaw_diff_rad = (float)(nn_align * PC_UNIT_SECTORANGLE * M_PI / 180.0);
        return {loop_id, yaw_diff_rad};
    }
};

extern "C" {

sco_handle* sco_create(int R, int S, int K, double thr, double lidar_h, double max_r, int excl, int period, double ratio)
{
    sco_handle* h = new sco_handle();
    h->PC_NUM_RING = R; h->PC_NUM_SECTOR = S; h->NUM_CANDIDATES_FROM_TREE = K;
    h->SC_DIST_THRES = thr; h->LIDAR_HEIGHT = lidar_h; h->PC_MAX_RADIUS = max_r;
    h->NUM_EXCLUDE_RECENT = excl; h->TREE_MAKING_PERIOD_ = period; h->SEARCH_RATIO = ratio;
    h->PC_UNIT_SECTORANGLE = 360.0 / double(S);   /* deviation Q1 */
    h->PC_UNIT_RINGGAP = max_r / double(R);
    h->tree_making_period_conter = 0;
    h->n_tree = 0;
    return h;
}
void sco_destroy(sco_handle* h) { delete h; }

void sco_make_scancontext(sco_handle* h, const float* pts, int n, int stride, float* out_desc, int* out_ring, int* out_sector)
{
    std::vector<double> desc;
    h->makeScancontext(pts, n, stride, desc, out_desc, out_ring, out_sector);
}

int sco_make_and_save(sco_handle* h, const float* pts, int n, int stride, int8_t robot, int index, float* out_desc)
{
    std::vector<double> desc;
    h->makeScancontext(pts, n, stride, desc, out_desc, nullptr, nullptr);
    return h->save(desc, robot, index);
}

int sco_save(sco_handle* h, const float* wire, int8_t robot, int index)
{
    const size_t n = (size_t)h->PC_NUM_RING * h->PC_NUM_SECTOR;
    std::vector<double> sc(n);
    for (size_t i = 0; i < n; i++) sc[i] = wire[i];   /* descriptor.h:1575-1582 */
    return h->save(sc, robot, index);
}

int sco_bulk_load(sco_handle* h, const float* wires, int n, int borrow)
{
    const size_t rs = (size_t)h->PC_NUM_RING * h->PC_NUM_SECTOR;
    const int base = (int)h->polarcontexts_.size();
    h->polarcontexts_.resize(base + n);
    h->polarcontext_invkeys_mat_.resize(base + n);
    h->polarcontext_indexs_.resize(base + n);
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::vector<float>> own(borrow ? 0 : n);
    auto work = [&](int i0, int i1) {
        for (int i = i0; i < i1; i++) {
            const float* w = wires + (size_t)i * rs;
            if (!borrow) { own[i].assign(w, w + rs); w = own[i].data(); }
            h->polarcontexts_[base + i] = w;
            h->makeRingkey(h->widen(w), h->polarcontext_invkeys_mat_[base + i]);
            h->polarcontext_indexs_[base + i] = std::make_pair((int8_t)0, base + i);
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < hw; t++) th.emplace_back(work, (int)((long long)n * t / hw), (int)((long long)n * (t + 1) / hw));
    for (auto& t : th) t.join();
    for (auto& v : own) h->owned_.push_back(std::move(v));
    return base + n;
}

int sco_size(sco_handle* h) { return (int)h->polarcontext_indexs_.size(); }

void sco_get_index(sco_handle* h, int key, int* robot, int* index)
{
    if (key < 0 || key >= (int)h->polarcontext_indexs_.size()) { *robot = -1; *index = -1; return; }
    *robot = h->polarcontext_indexs_[key].first; *index = h->polarcontext_indexs_[key].second;
}

void sco_get_desc(sco_handle* h, int key, float* out)
{
    std::memcpy(out, h->polarcontexts_[key], sizeof(float) * h->PC_NUM_RING * h->PC_NUM_SECTOR);
}

void sco_ring_key(sco_handle* h, int key, float* out)
{
    std::memcpy(out, h->polarcontext_invkeys_mat_[key].data(), sizeof(float) * h->PC_NUM_RING);
}

void sco_sector_key(sco_handle* h, int key, double* out)
{
    std::vector<double> v;
    h->makeSectorkey(h->widen(h->polarcontexts_[key]), v);
    std::memcpy(out, v.data(), sizeof(double) * v.size());
}

void sco_distance(sco_handle* h, int k1, int k2, double* dist, int* shift)
{
    const std::pair<double, int> r = h->distanceBtnScanContext(h->widen(h->polarcontexts_[k1]), h->widen(h->polarcontexts_[k2]));
    *dist = r.first; *shift = r.second;
}

void sco_distance_raw(sco_handle* h, const float* d1, const float* d2, double* dist, int* shift)
{
    const size_t n = (size_t)h->PC_NUM_RING * h->PC_NUM_SECTOR;
    std::vector<double> a(d1, d1 + n), b(d2, d2 + n);
    const std::pair<double, int> r = h->distanceBtnScanContext(a, b);
    *dist = r.first; *shift = r.second;
}

int sco_fast_align(sco_handle* h, int k1, int k2)
{
    std::vector<double> v1, v2;
    h->makeSectorkey(h->widen(h->polarcontexts_[k1]), v1);
    h->makeSectorkey(h->widen(h->polarcontexts_[k2]), v2);
    return h->fastAlignUsingVkey(v1, v2);
}

double sco_dist_direct(sco_handle* h, int k1, int k2, int shift)
{
    return h->distDirectSC(h->widen(h->polarcontexts_[k1]), h->widen(h->polarcontexts_[k2]), shift);
}

void sco_detect_intra(sco_handle* h, int cur, int* id, float* second)
{
    const std::pair<int, float> r = h->detectIntra(cur);
    *id = r.first; *second = r.second;
}

void sco_detect_inter(sco_handle* h, int cur, int* id, float* second)
{
    const std::pair<int, float> r = h->detectInter(cur);
    *id = r.first; *second = r.second;
}

int sco_knn(sco_handle* h, int cur, int n_db, int k, int metric, int32_t* ids, float* d2)
{
    return h->knn(h->polarcontext_invkeys_mat_[cur].data(), n_db, k, metric, ids, d2);
}

void sco_query_batch(sco_handle* h, const int32_t* queries, int nq, int n_db, int k, int metric, int nthreads,
                     int32_t* cand_ids, float* cand_d2, double* cand_dist, int32_t* cand_shift,
                     int32_t* best_id, double* best_dist, int32_t* best_shift)
{
    auto work = [&](int t0, int t1) {
        for (int qi = t0; qi < t1; qi++) {
            const int cur = queries[qi];
            int32_t* ids = cand_ids + (size_t)qi * k;
            float* d2 = cand_d2 + (size_t)qi * k;
            h->knn(h->polarcontext_invkeys_mat_[cur].data(), n_db, k, metric, ids, d2);
            double min_dist = 10000000; int nn_align = 0, nn_idx = -1;
            for (int i = 0; i < k; i++) {
                double dist = NAN; int shift = 0;
                if (ids[i] >= 0) {
                    const std::pair<double, int> r = h->distanceBtnScanContext(h->widen(h->polarcontexts_[cur]), h->widen(h->polarcontexts_[ids[i]]));
                    dist = r.first; shift = r.second;
                    if (dist < min_dist && ids[i] != cur) { min_dist = dist; nn_align = shift; nn_idx = ids[i]; }
                }
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
float sco_atanf_port(float x) { return atanf_port(x); }

} // extern "C"
