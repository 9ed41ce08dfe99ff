/*
 * rowkey_oracle.cpp — CPU restatement of the row-key candidate stage of class lidar_iris_descriptor
 * (/root/reference/include/descriptor.h). TEST INFRASTRUCTURE ONLY, see rowkey_oracle.h. Checked against the reference's
 * own text (oracle/_ref/libiris_ref.so) by tests/test_rowkey_oracle.py and the fixtures it generated (tests/golden).
 */
#include "rowkey_oracle.h"

#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <thread>
#include <utility>
#include <vector>

namespace {

/* libnabo's rules as the stand-in states them (oracle/ref_shim/ref_standins.h, Nabo::NNSearchF::knn): sequential float
 * accumulation over the key, d2 <= epsilon skipped (no ALLOW_SELF_MATCH), strict < against the current worst, ascending
 * results, unfilled slots -1 / +inf; exact ties keep the lower index */
void knn_nabo(const float* keys, int n, int rows, const float* q, int k, int32_t* idx, float* d2)
{
    int count = 0;
    for (int i = 0; i < k; i++) { idx[i] = -1; d2[i] = std::numeric_limits<float>::infinity(); }
    for (int j = 0; j < n; j++) {
        const float* key = keys + (size_t)j * rows;
        float dist = 0;
        for (int d = 0; d < rows; d++) { const float diff = q[d] - key[d]; dist += diff * diff; }
        if (!(dist > std::numeric_limits<float>::epsilon())) continue;
        if (!(dist < (count == k ? d2[k - 1] : std::numeric_limits<float>::infinity()))) continue;
        int i = count < k ? count : k - 1;
        for (; i > 0 && d2[i - 1] > dist; --i) { d2[i] = d2[i - 1]; idx[i] = idx[i - 1]; }
        d2[i] = dist; idx[i] = j;
        if (count < k) count++;
    }
}

struct Entry { float feature; int tag; };

float compare(const Entry& a, const Entry& b, int* bias)
{
    *bias = (7 * a.tag + 13 * b.tag) % 360;
    return std::fabs(a.feature - b.feature);
}

} // namespace

struct sco_iris {
    int rows, exclude, K, robot_num, this_id;
    double thres;
    std::vector<std::vector<float>> keys;                 /* irisFeatureRowKey: per robot [n][rows] */
    std::vector<std::vector<Entry>> features;             /* irisFeatures */
    std::vector<std::vector<int>> local2global;
    std::vector<std::pair<int, int>> index;               /* irisFeatureIndexs */
};

extern "C" {

sco_iris* sco_iris_create(int rows, int num_exclude_recent, int num_candidates, double dist_thres, int robot_num, int this_id)
{
    sco_iris* h = new sco_iris();
    h->rows = rows; h->exclude = num_exclude_recent; h->K = num_candidates; h->thres = dist_thres; h->robot_num = robot_num; h->this_id = this_id;
    h->keys.resize(robot_num); h->features.resize(robot_num); h->local2global.resize(robot_num);   /* descriptor.h:500-510 */
    return h;
}
void sco_iris_destroy(sco_iris* h) { delete h; }

/* descriptor.h:1047-1059 */
void sco_iris_save(sco_iris* h, const float* row_key, int robot, int index, float feature)
{
    h->features[robot].push_back(Entry{feature, (int)h->index.size()});
    h->keys[robot].insert(h->keys[robot].end(), row_key, row_key + h->rows);
    h->local2global[robot].push_back((int)h->index.size());
    h->index.push_back(std::make_pair(robot, index));
}

/* descriptor.h:1087-1148 */
void sco_iris_detect_intra(sco_iris* h, int cur_ptr, int* id, float* bias, int* n_cand, int32_t* cand, float* cand_d2)
{
    *id = -1; *bias = 0.0f; *n_cand = 0;
    const int me = h->this_id;
    const float* cur_key = h->keys[me].data() + (size_t)cur_ptr * h->rows;
    const Entry cur = h->features[me][cur_ptr];
    if (cur_ptr < h->exclude + h->K + 1) return;                                      /* :1094-1097 */
    const int history = cur_ptr - h->exclude;                                         /* :1101 */
    std::vector<int32_t> indice(h->K); std::vector<float> distance(h->K);
    knn_nabo(h->keys[me].data(), history, h->rows, cur_key, h->K, indice.data(), distance.data());   /* :1104-1114 */
    *n_cand = h->K;
    for (int i = 0; i < h->K; i++) { cand[i] = indice[i]; cand_d2[i] = distance[i]; }
    float min_dis = 10000000.0f; int min_index = -1, min_bias = 0;
    for (int i = 0; i < h->K; i++) {
        if ((size_t)indice[i] >= h->local2global[me].size()) continue;               /* :1119: int against size_t, so -1 is skipped too */
        int b;
        const float dis = compare(cur, h->features[me][indice[i]], &b);
        if (dis < min_dis) { min_dis = dis; min_index = indice[i]; min_bias = b; }
    }
    if (min_dis < h->thres) { *id = min_index; *bias = (float)min_bias; }             /* :1138-1142 */
}

/* descriptor.h:1150-1250 */
void sco_iris_detect_inter(sco_iris* h, int cur_ptr, int* id, float* bias, int* n_cand, int32_t* cand, float* cand_d2)
{
    *id = -1; *bias = 0.0f; *n_cand = 0;
    const int cur_robot = h->index[cur_ptr].first, cur_index = h->index[cur_ptr].second;
    const float* cur_key = h->keys[cur_robot].data() + (size_t)cur_index * h->rows;
    const Entry cur = h->features[cur_robot][cur_index];
    std::vector<float> new_keys; std::vector<int> new_l2g; std::vector<Entry> new_feat;
    if (cur_robot == h->this_id) {
        for (int i = 0; i < h->robot_num; i++) {
            if (i != h->this_id && h->local2global[i].size() > 0) {                   /* :1166-1178 */
                const size_t add = h->local2global[i].size();
                new_keys.insert(new_keys.end(), h->keys[i].begin(), h->keys[i].begin() + add * h->rows);
                new_l2g.insert(new_l2g.end(), h->local2global[i].begin(), h->local2global[i].end());
                new_feat.insert(new_feat.end(), h->features[i].begin(), h->features[i].end());
            }
        }
    } else if (h->local2global[h->this_id].size() > 0) {                               /* :1180-1191 */
        const int me = h->this_id; const size_t add = h->local2global[me].size();
        new_keys.insert(new_keys.end(), h->keys[me].begin(), h->keys[me].begin() + add * h->rows);
        new_l2g.insert(new_l2g.end(), h->local2global[me].begin(), h->local2global[me].end());
        new_feat.insert(new_feat.end(), h->features[me].begin(), h->features[me].end());
    }
    if ((int)new_l2g.size() < h->K + 1) return;                                        /* :1194-1197 */
    std::vector<int32_t> indice(h->K); std::vector<float> distance(h->K);
    knn_nabo(new_keys.data(), (int)new_l2g.size(), h->rows, cur_key, h->K, indice.data(), distance.data());
    *n_cand = h->K;
    for (int i = 0; i < h->K; i++) { cand[i] = indice[i]; cand_d2[i] = distance[i]; }
    float min_dis = 10000000.0f; int min_index = -1, min_bias = 0;
    for (int i = 0; i < h->K; i++) {
        if ((size_t)indice[i] >= new_l2g.size()) continue;                            /* :1214-1218 */
        int b;
        const float dis = compare(cur, new_feat[indice[i]], &b);
        if (dis < min_dis) { min_dis = dis; min_index = new_l2g[indice[i]]; min_bias = b; }
    }
    if (min_dis < h->thres) { *id = min_index; *bias = (float)min_bias; }
}

void sco_iris_get_index(sco_iris* h, int key, int* robot, int* index) { *robot = h->index[key].first; *index = h->index[key].second; }
int sco_iris_size(sco_iris* h, int id_in) { return id_in == -1 ? (int)h->index.size() : (int)h->local2global[id_in].size(); }

void sco_iris_knn_batch(sco_iris* h, const float* q_keys, int Q, int robot, int n, int K, int threads, int32_t* idx, float* d2)
{
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([=]() {
            for (int q = t; q < Q; q += threads)
                knn_nabo(h->keys[robot].data(), n, h->rows, q_keys + (size_t)q * h->rows, K, idx + (size_t)q * K, d2 + (size_t)q * K);
        });
    for (auto& th : pool) th.join();
}

} // extern "C"
