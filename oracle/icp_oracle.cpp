/*
 * icp_oracle.cpp — CPU ORACLE (test infrastructure, not product) for geometric verification.
 *
 * Restates what /root/reference/include/distributedMapping.h:1108-1132 asks of
 * pcl::IterativeClosestPoint (point-to-point, max correspondence distance 100 m, 50 iterations,
 * transformation epsilon 1e-6, euclidean fitness epsilon 1e-6, no RANSAC) and the
 * pcl::VoxelGrid call of :996-998 / :1183-1184.
 *
 * PARITY UNPINNED: PCL is not vendored in the reference (dependencies.rosinstall:33-36 pins
 * zhongshp5/pcl_catkin @ afe789a, not fetchable here), so this file follows PCL's published
 * algorithm — pcl::IterativeClosestPoint::computeTransformation, DefaultConvergenceCriteria::
 * hasConverged, TransformationEstimationSVD (closed-form rigid fit), getFitnessScore (mean
 * squared nearest-neighbour distance over the source) — and DEFINES the expected values.
 * Statistics are accumulated in double (PCL: float), the closed-form rotation is Horn's
 * quaternion solution (same optimum as PCL's SVD with the reflection fix).
 */
#include "sc_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <random>
#include <unordered_map>
#include <vector>

namespace {

struct P3 { double x, y, z; };

/* exact nearest neighbour on a uniform grid: expanding Chebyshev shells, stop once the best
 * distance cannot be beaten by any unvisited cell; falls back to a full scan for isolated points */
struct GridNN {
    double cell; std::vector<P3> pts;
    std::unordered_map<long long, std::vector<int>> cells;
    static long long key(long long i, long long j, long long k) { return ((i + (1LL << 20)) << 42) | ((j + (1LL << 20)) << 21) | (k + (1LL << 20)); }
    void build(const std::vector<P3>& p, double c)
    {
        pts = p; cell = c;
        for (int i = 0; i < (int)p.size(); i++)
            cells[key((long long)std::floor(p[i].x / c), (long long)std::floor(p[i].y / c), (long long)std::floor(p[i].z / c))].push_back(i);
    }
    int nearest(const P3& q, double* d2out) const
    {
        const long long ci = (long long)std::floor(q.x / cell), cj = (long long)std::floor(q.y / cell), ck = (long long)std::floor(q.z / cell);
        double best = DBL_MAX; int bi = -1;
        const int max_shell = 6;
        for (int r = 0; r <= max_shell; r++) {
            for (long long i = ci - r; i <= ci + r; i++)
                for (long long j = cj - r; j <= cj + r; j++)
                    for (long long k = ck - r; k <= ck + r; k++) {
                        if (std::max(std::max(std::llabs(i - ci), std::llabs(j - cj)), std::llabs(k - ck)) != r) continue;
                        auto it = cells.find(key(i, j, k));
                        if (it == cells.end()) continue;
                        for (int id : it->second) {
                            const double dx = pts[id].x - q.x, dy = pts[id].y - q.y, dz = pts[id].z - q.z;
                            const double d2 = dx * dx + dy * dy + dz * dz;
                            if (d2 < best || (d2 == best && id < bi)) { best = d2; bi = id; }
                        }
                    }
            if (bi >= 0 && best <= (double)r * cell * (double)r * cell) { *d2out = best; return bi; }
        }
        for (int id = 0; id < (int)pts.size(); id++) {
            const double dx = pts[id].x - q.x, dy = pts[id].y - q.y, dz = pts[id].z - q.z;
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best || (d2 == best && id < bi)) { best = d2; bi = id; }
        }
        *d2out = best; return bi;
    }
};

/* largest-eigenvalue eigenvector of a symmetric 4x4 by cyclic Jacobi */
void jacobi4(double A[4][4], double V[4][4])
{
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) V[i][j] = (i == j);
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = 0;
        for (int i = 0; i < 4; i++) for (int j = i + 1; j < 4; j++) off += A[i][j] * A[i][j];
        if (off < 1e-300) break;
        for (int p = 0; p < 4; p++)
            for (int q = p + 1; q < 4; q++) {
                if (std::fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 4; k++) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 4; k++) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 4; k++) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
            }
    }
}

/* closed-form rigid fit src -> tgt over paired points (Horn 1987) */
void rigid_fit(const std::vector<P3>& a, const std::vector<P3>& b, double T[16])
{
    const int n = (int)a.size();
    P3 ca{0, 0, 0}, cb{0, 0, 0};
    for (int i = 0; i < n; i++) { ca.x += a[i].x; ca.y += a[i].y; ca.z += a[i].z; cb.x += b[i].x; cb.y += b[i].y; cb.z += b[i].z; }
    ca.x /= n; ca.y /= n; ca.z /= n; cb.x /= n; cb.y /= n; cb.z /= n;
    double S[3][3] = {{0}};
    for (int i = 0; i < n; i++) {
        const double p[3] = {a[i].x - ca.x, a[i].y - ca.y, a[i].z - ca.z};
        const double q[3] = {b[i].x - cb.x, b[i].y - cb.y, b[i].z - cb.z};
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) S[r][c] += p[r] * q[c];
    }
    double N[4][4] = {
        {S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
        {S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
        {S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
        {S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
    double V[4][4];
    jacobi4(N, V);
    int best = 0;
    for (int i = 1; i < 4; i++) if (N[i][i] > N[best][best]) best = i;
    double qw = V[0][best], qx = V[1][best], qy = V[2][best], qz = V[3][best];
    const double nq = std::sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
    qw /= nq; qx /= nq; qy /= nq; qz /= nq;
    const double R[3][3] = {
        {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)},
        {2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)},
        {2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)}};
    const double t[3] = {cb.x - (R[0][0] * ca.x + R[0][1] * ca.y + R[0][2] * ca.z),
                         cb.y - (R[1][0] * ca.x + R[1][1] * ca.y + R[1][2] * ca.z),
                         cb.z - (R[2][0] * ca.x + R[2][1] * ca.y + R[2][2] * ca.z)};
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T[r * 4 + c] = R[r][c]; T[r * 4 + 3] = t[r]; }
    T[12] = T[13] = T[14] = 0; T[15] = 1;
}

void matmul4(const double A[16], const double B[16], double C[16])
{
    double tmp[16];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) {
        double s = 0; for (int k = 0; k < 4; k++) s += A[r * 4 + k] * B[k * 4 + c];
        tmp[r * 4 + c] = s;
    }
    std::memcpy(C, tmp, sizeof(tmp));
}

std::vector<P3> load(const float* p, int n, int stride)
{
    std::vector<P3> v(n);
    for (int i = 0; i < n; i++) v[i] = P3{p[(size_t)i * stride], p[(size_t)i * stride + 1], p[(size_t)i * stride + 2]};
    return v;
}

} // namespace

extern "C" {

int sco_icp(const float* src, int n_src, const float* tgt, int n_tgt, int stride,
            double max_corr_dist, int max_iter, double trans_eps, double fit_eps,
            float* T_out, float* fitness, int* converged)
{
    std::vector<P3> s = load(src, n_src, stride), t = load(tgt, n_tgt, stride);
    GridNN nn; nn.build(t, 1.0);
    double final_T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    /* icp.hpp (PCL 1.8 - 1.12): setRelativeMSE(euclidean_fitness_epsilon_), setTranslationThreshold(transformation_epsilon_),
     * setRotationThreshold(1 - transformation_epsilon_); absolute-MSE threshold at its default 1e-12 */
    const double rotation_threshold = 1.0 - trans_eps, mse_rel = fit_eps, mse_abs = 1e-12;
    const double max_d2 = max_corr_dist * max_corr_dist;
    double prev_mse = DBL_MAX;
    int iterations = 0; bool conv = false;
    std::vector<P3> cur = s, a, b;
    while (true) {
        a.clear(); b.clear();
        double sum_d2 = 0;
        for (int i = 0; i < n_src; i++) {
            double d2; const int j = nn.nearest(cur[i], &d2);
            if (j < 0 || d2 > max_d2) continue;
            a.push_back(cur[i]); b.push_back(t[j]); sum_d2 += d2;
        }
        if ((int)a.size() < 3) { conv = false; break; }
        double T[16];
        rigid_fit(a, b, T);
        for (auto& p : cur) {
            const P3 q = p;
            p.x = T[0] * q.x + T[1] * q.y + T[2] * q.z + T[3];
            p.y = T[4] * q.x + T[5] * q.y + T[6] * q.z + T[7];
            p.z = T[8] * q.x + T[9] * q.y + T[10] * q.z + T[11];
        }
        matmul4(T, final_T, final_T);
        ++iterations;
        /* DefaultConvergenceCriteria::hasConverged */
        if (iterations >= max_iter) { conv = true; break; }
        const double cos_angle = 0.5 * (T[0] + T[5] + T[10] - 1);
        const double translation_sqr = T[3] * T[3] + T[7] * T[7] + T[11] * T[11];
        if (cos_angle >= rotation_threshold && translation_sqr <= trans_eps) { conv = true; break; }
        const double cur_mse = sum_d2 / (double)a.size();
        if (std::fabs(cur_mse - prev_mse) < mse_abs) { conv = true; break; }
        if (std::fabs(cur_mse - prev_mse) / prev_mse < mse_rel) { conv = true; break; }
        prev_mse = cur_mse;
    }
    /* getFitnessScore(): mean squared NN distance of the aligned source */
    double fs = 0; int nr = 0;
    for (int i = 0; i < n_src; i++) { double d2; if (nn.nearest(cur[i], &d2) >= 0) { fs += d2; nr++; } }
    *fitness = nr > 0 ? (float)(fs / nr) : FLT_MAX;
    for (int i = 0; i < 16; i++) T_out[i] = (float)final_T[i];
    *converged = conv ? 1 : 0;
    return iterations;
}

/* RANSAC + SVD verification as geometricVerificationService drives PCL (distributedMapping.h:1211-1243):
 * CorrespondenceEstimation (nearest target of every source point), CorrespondenceRejectorSampleConsensus
 * (pcl::RandomSampleConsensus over SampleConsensusModelRegistration: 3-point samples, inlier = residual <= threshold,
 * adaptive stop k = log(1-0.99)/log(1-w^3), at most max_iter), TransformationEstimationSVD on the inliers,
 * success iff inliers >= ratio * correspondences. PARITY UNPINNED (PCL not vendored; its sampler is random):
 * the GPU result is compared statistically (same verdict, pose and inlier ratio within tolerance). */
int sco_verify_ransac(const float* src, int n_src, const float* tgt, int n_tgt, int stride, int max_iter, double thr, double ratio,
                      unsigned seed, float* T_out, int* n_corr, int* n_inliers, int* success)
{
    for (int i = 0; i < 16; i++) T_out[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    *n_corr = 0; *n_inliers = 0; *success = 0;
    if (n_src < 3 || n_tgt < 1) return 0;
    std::vector<P3> s = load(src, n_src, stride), t = load(tgt, n_tgt, stride);
    GridNN nn; nn.build(t, 1.0);
    std::vector<P3> b(n_src);
    for (int i = 0; i < n_src; i++) { double d2; b[i] = t[nn.nearest(s[i], &d2)]; }
    std::mt19937 rng(seed);
    std::uniform_int_distribution<int> pick(0, n_src - 1);
    const double thr2 = thr * thr;
    int best_cnt = 0; double bestT[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double k = max_iter; int it = 0;
    while (it < k && it < max_iter) {
        int i0 = pick(rng), i1 = pick(rng), i2 = pick(rng);
        ++it;
        if (i0 == i1 || i0 == i2 || i1 == i2) continue;
        std::vector<P3> pa{s[i0], s[i1], s[i2]}, pb{b[i0], b[i1], b[i2]};
        double T[16];
        rigid_fit(pa, pb, T);
        int cnt = 0;
        for (int i = 0; i < n_src; i++) {
            const double x = T[0] * s[i].x + T[1] * s[i].y + T[2] * s[i].z + T[3] - b[i].x;
            const double y = T[4] * s[i].x + T[5] * s[i].y + T[6] * s[i].z + T[7] - b[i].y;
            const double z = T[8] * s[i].x + T[9] * s[i].y + T[10] * s[i].z + T[11] - b[i].z;
            cnt += (x * x + y * y + z * z <= thr2);
        }
        if (cnt > best_cnt) {
            best_cnt = cnt; std::memcpy(bestT, T, sizeof(T));
            const double w = (double)cnt / n_src, p3 = std::max(1e-12, std::min(1.0 - 1e-12, w * w * w));
            k = std::log(1.0 - 0.99) / std::log(1.0 - p3);
        }
    }
    std::vector<P3> ia, ib;
    for (int i = 0; i < n_src; i++) {
        const double x = bestT[0] * s[i].x + bestT[1] * s[i].y + bestT[2] * s[i].z + bestT[3] - b[i].x;
        const double y = bestT[4] * s[i].x + bestT[5] * s[i].y + bestT[6] * s[i].z + bestT[7] - b[i].y;
        const double z = bestT[8] * s[i].x + bestT[9] * s[i].y + bestT[10] * s[i].z + bestT[11] - b[i].z;
        if (x * x + y * y + z * z <= thr2) { ia.push_back(s[i]); ib.push_back(b[i]); }
    }
    *n_corr = n_src; *n_inliers = (int)ia.size();
    if (ia.size() >= 3) { double T[16]; rigid_fit(ia, ib, T); for (int i = 0; i < 16; i++) T_out[i] = (float)T[i]; }
    *success = ((double)ia.size() >= ratio * n_src) ? 1 : 0;
    return it;
}

void sco_nn(const float* src, int n_src, const float* tgt, int n_tgt, int stride, int32_t* idx, float* d2)
{
    for (int i = 0; i < n_src; i++) {
        const float* p = src + (size_t)i * stride;
        double best = DBL_MAX; int bi = -1;
        for (int j = 0; j < n_tgt; j++) {
            const float* q = tgt + (size_t)j * stride;
            const double dx = (double)q[0] - p[0], dy = (double)q[1] - p[1], dz = (double)q[2] - p[2];
            const double d = dx * dx + dy * dy + dz * dz;
            if (d < best) { best = d; bi = j; }
        }
        idx[i] = bi; d2[i] = (float)best;
    }
}

int sco_voxel_grid(const float* pts, int n, int stride, float leaf, float* out)
{
    /* pcl::VoxelGrid::applyFilter: leaf index = floor(p/leaf) per axis, one centroid per leaf,
     * output ordered by linear leaf index (x fastest). Points with non-finite coordinates are skipped. */
    if (n <= 0) return 0;
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int i = 0; i < n; i++) {
        const float* p = pts + (size_t)i * stride;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], (double)p[a]); mx[a] = std::max(mx[a], (double)p[a]); }
    }
    const double inv = 1.0 / leaf;
    long long mnb[3], dv[3];
    for (int a = 0; a < 3; a++) { mnb[a] = (long long)std::floor(mn[a] * inv); dv[a] = (long long)std::floor(mx[a] * inv) - mnb[a] + 1; }
    std::vector<std::pair<long long, int>> keyed;
    for (int i = 0; i < n; i++) {
        const float* p = pts + (size_t)i * stride;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        const long long ix = (long long)std::floor(p[0] * inv) - mnb[0], iy = (long long)std::floor(p[1] * inv) - mnb[1], iz = (long long)std::floor(p[2] * inv) - mnb[2];
        keyed.emplace_back(ix + iy * dv[0] + iz * dv[0] * dv[1], i);
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<long long, int>& a, const std::pair<long long, int>& b) { return a.first < b.first; });
    int m = 0;
    for (size_t i = 0; i < keyed.size();) {
        size_t j = i; double sx = 0, sy = 0, sz = 0;
        for (; j < keyed.size() && keyed[j].first == keyed[i].first; j++) {
            const float* p = pts + (size_t)keyed[j].second * stride; sx += p[0]; sy += p[1]; sz += p[2];
        }
        const double c = (double)(j - i);
        out[m * 3] = (float)(sx / c); out[m * 3 + 1] = (float)(sy / c); out[m * 3 + 2] = (float)(sz / c);
        m++; i = j;
    }
    return m;
}

} // extern "C"
