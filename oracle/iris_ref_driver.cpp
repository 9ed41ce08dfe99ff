/*
 * iris_ref_driver.cpp — builds oracle/_ref/libiris_ref.so: the reference's OWN text of the row-key candidate stage of
 * class lidar_iris_descriptor, compiled behind the sco_iris_* entry points of oracle/rowkey_oracle.h.
 * TEST INFRASTRUCTURE ONLY; recipe in oracle/Makefile (target _ref/libiris_ref.so).
 *
 * Three pieces of /root/reference/include/descriptor.h are cut at build time into a temporary directory (never into the repo):
 *   iris_save.inc     void save(const cv::Mat1b iris, Eigen::MatrixXf rowKey, ...)                   :1047-1063
 *   iris_detect.inc   detectIntraLoopClosureID, detectInterLoopClosureID, getIndex, getSize          :1087-1267
 *   iris_members.inc  the class's member list, "private:" to its closing "};"                        :1269-1302
 * and included into the shell class below, which supplies what the rest of that class (OpenCV image code, absent here)
 * would: the constructor's member initialisation (:486-510, restated), a featureDesc that is one float + a tag, and
 * compare() / getFeature() stubs (see rowkey_oracle.h). Eigen and libnabo are the stand-ins of ref_shim/ref_standins.h.
 */
#include "ref_standins.h"

#include <climits>

using namespace std; /* descriptor.h:19 */

namespace cv { struct Mat1b {}; }

/* the kNN list of the last Nabo call of this thread, for sco_iris_detect_*'s cand outputs */
static thread_local std::vector<int> g_last_idx;
static thread_local std::vector<float> g_last_d2;

namespace NaboIris {
struct NNSearchF {
    Eigen::MatrixXf cloud; int dim;
    /* the reference never frees its trees (descriptor.h:1104,1199): one per thread is reused here instead */
    static NNSearchF* createKDTreeTreeHeap(const Eigen::MatrixXf& c, int d = INT_MAX)
    {
        static thread_local NNSearchF s;
        s.cloud = c; s.dim = d < c.rows() ? d : c.rows();
        return &s;
    }
    void knn(const Eigen::VectorXf& q, Eigen::VectorXi& idx, Eigen::VectorXf& d2, int k) const
    {
        Nabo::NNSearchF lin; lin.cloud = cloud; lin.dim = dim;
        lin.knn(q, idx, d2, k);
        g_last_idx.assign(k, -1); g_last_d2.assign(k, 0.f);
        for (int i = 0; i < k; i++) { g_last_idx[i] = idx[i]; g_last_d2[i] = d2[i]; }
    }
};
} // namespace NaboIris
#define Nabo NaboIris

class lidar_iris_candidates_ref
{
public:
    struct featureDesc { float f; int tag; };

    lidar_iris_candidates_ref(int rows, int numExcludeRecent, int numCandidates, double distThres, int robotNum, int thisID) :
        _rows(rows), _cols(360), _nscan(64), _distThres(distThres), _nscale(4), _minWaveLength(18), _mult(1.6), _sigmaOnf(0.75),
        _matchNum(2), _robotNum(robotNum), _thisID(thisID), _numExcludeRecent(numExcludeRecent), _numCandidates(numCandidates)
    {
        for (int i = 0; i < _robotNum; i++) {            /* descriptor.h:500-510 */
            irisFeatures.push_back(std::vector<featureDesc>());
            irisFeatureRowKey.push_back(Eigen::MatrixXf());
            local2Global.push_back(std::vector<int>());
        }
    }

    featureDesc pending;                                  /* what getFeature() returns for the entry being saved */
    featureDesc getFeature(const cv::Mat1b&) { return pending; }
    float compare(const featureDesc& a, const featureDesc& b, int* bias)
    {
        *bias = (7 * a.tag + 13 * b.tag) % 360;
        return fabs(a.f - b.f);
    }
    int total() const { return (int)irisFeatureIndexs.size(); }

#include "iris_save.inc"
#include "iris_detect.inc"
#include "iris_members.inc"

#undef Nabo

#include "../rowkey_oracle.h"
#include <thread>

struct sco_iris { lidar_iris_candidates_ref* c; int rows, K; std::vector<std::vector<float>> flat; };

extern "C" {

sco_iris* sco_iris_create(int rows, int num_exclude_recent, int num_candidates, double dist_thres, int robot_num, int this_id)
{
    sco_iris* h = new sco_iris();
    h->c = new lidar_iris_candidates_ref(rows, num_exclude_recent, num_candidates, dist_thres, robot_num, this_id);
    h->rows = rows; h->K = num_candidates; h->flat.resize(robot_num);
    return h;
}
void sco_iris_destroy(sco_iris* h) { delete h->c; delete h; }

void sco_iris_save(sco_iris* h, const float* row_key, int robot, int index, float feature)
{
    Eigen::MatrixXf key(h->rows, 1);
    for (int r = 0; r < h->rows; r++) key(r, 0) = row_key[r];
    h->c->pending.f = feature; h->c->pending.tag = h->c->total();
    h->c->save(cv::Mat1b(), key, (int8_t)robot, index);
    h->flat[robot].insert(h->flat[robot].end(), row_key, row_key + h->rows);
}

static void report(sco_iris* h, int* n_cand, int32_t* cand, float* cand_d2)
{
    *n_cand = (int)g_last_idx.size();
    for (size_t i = 0; i < g_last_idx.size(); i++) { cand[i] = g_last_idx[i]; cand_d2[i] = g_last_d2[i]; }
}

void sco_iris_detect_intra(sco_iris* h, int cur_ptr, int* id, float* bias, int* n_cand, int32_t* cand, float* cand_d2)
{
    g_last_idx.clear(); g_last_d2.clear();
    const std::pair<int, float> r = h->c->detectIntraLoopClosureID(cur_ptr);
    *id = r.first; *bias = r.second;
    report(h, n_cand, cand, cand_d2);
}
void sco_iris_detect_inter(sco_iris* h, int cur_ptr, int* id, float* bias, int* n_cand, int32_t* cand, float* cand_d2)
{
    g_last_idx.clear(); g_last_d2.clear();
    const std::pair<int, float> r = h->c->detectInterLoopClosureID(cur_ptr);
    *id = r.first; *bias = r.second;
    report(h, n_cand, cand, cand_d2);
}
void sco_iris_get_index(sco_iris* h, int key, int* robot, int* index)
{
    const std::pair<int8_t, int> p = h->c->getIndex(key);
    *robot = p.first; *index = p.second;
}
int sco_iris_size(sco_iris* h, int id_in) { return h->c->getSize(id_in); }

void sco_iris_knn_batch(sco_iris* h, const float* q_keys, int Q, int robot, int n, int K, int threads, int32_t* idx, float* d2)
{
    /* the stand-in's linear scan on the same keys (the reference has no batched form) */
    Eigen::MatrixXf cloud(h->rows, n);
    for (int j = 0; j < n; j++) for (int r = 0; r < h->rows; r++) cloud(r, j) = h->flat[robot][(size_t)j * h->rows + r];
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() {
            ::Nabo::NNSearchF lin; lin.cloud = cloud; lin.dim = h->rows;
            for (int q = t; q < Q; q += threads) {
                Eigen::VectorXf qq(h->rows); Eigen::VectorXi ii(K); Eigen::VectorXf dd(K);
                for (int r = 0; r < h->rows; r++) qq[r] = q_keys[(size_t)q * h->rows + r];
                lin.knn(qq, ii, dd, K);
                for (int i = 0; i < K; i++) { idx[(size_t)q * K + i] = ii[i]; d2[(size_t)q * K + i] = dd[i]; }
            }
        });
    for (auto& th : pool) th.join();
}

} // extern "C"
