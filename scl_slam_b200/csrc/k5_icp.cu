// k5_icp.cu — K5, ICP geometric verification.
//
// Replaces the pcl::IterativeClosestPoint block of performIntraLoopClosure,
// /root/reference/include/distributedMapping.h:1108-1132: point-to-point ICP of the current
// keyframe cloud onto the merged history cloud with max correspondence distance 100 m, 50
// iterations, transformation epsilon 1e-6, euclidean fitness epsilon 1e-6, no RANSAC, followed
// by getFitnessScore() (:1121) and getFinalTransformation() (:1132).
//
// Shape of one iteration (one fused kernel launch):
//   * the target cloud is indexed ONCE by two voxel hash grids (1 m and 8 m cells; open
//     addressing on 64-bit cell keys, points stored cell-sorted as float4 = x,y,z,index);
//   * each thread takes one source point, applies the current transform, and finds its EXACT
//     nearest target point by visiting Chebyshev shells of cells until the best distance can no
//     longer be beaten by an unvisited cell (fine grid shells 0..2, then coarse grid shells
//     0..2, then a full scan for isolated points). Ties: lower target index;
//   * the 6x6 point-to-point normal equations (J = [-[p]x | I]: 21 + 6 sums, plus the squared
//     error and the correspondence count) are reduced with warp shuffles in FP64, one atomicAdd
//     per warp per term.
// The host solves the 6x6 system (Cholesky, FP64), composes the increment (exponential map)
// and applies PCL's DefaultConvergenceCriteria with the thresholds pcl::IterativeClosestPoint hands it (icp.hpp, PCL 1.8 -
// 1.12): iterations >= max, or rotation cos >= 1 - transformation epsilon and translation^2 <= transformation epsilon, or
// |dMSE| < 1e-12, or relative dMSE < euclidean fitness epsilon.
// PARITY UNPINNED (PCL is not vendored in the reference): PCL's default per-iteration step is the closed-form
// TransformationEstimationSVD on the current correspondences, this kernel takes one Gauss-Newton step on the same
// objective (north_star's 6x6 normal-equation build). Both stop at the same fixed point — where the increment is the
// identity — which is what the tests compare (1e-3 m, 1e-3 rad, fitness 1e-3 rel.) against the oracle's closed-form iterates.
//
// Roofline: latency / L2 bound hash probes on a cloud that fits L2 (16 B/point); the kernel is
// reported by achieved GB/s only.
#include "engine_internal.h"
#include "common.cuh"

#include <cmath>
#include <vector>

namespace {

constexpr unsigned long long kEmpty = 0xffffffffffffffffull;

struct Grid {
    const unsigned long long* keys;   // [cap] cell key or kEmpty
    const int* start;                 // [cap] first point of the cell in pts
    const int* count;                 // [cap]
    const float4* pts;                // [n] cell-sorted points (w = original index)
    unsigned mask;                    // cap - 1
    float inv_cell, cell;
};

__host__ __device__ inline unsigned long long cell_key(int i, int j, int k)
{
    return ((unsigned long long)(unsigned)(i + (1 << 20)) << 42) | ((unsigned long long)(unsigned)(j + (1 << 20)) << 21) |
           (unsigned long long)(unsigned)(k + (1 << 20));
}
__device__ inline unsigned hash_key(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}
__device__ inline int clampi(float v) { return (int)fminf(fmaxf(floorf(v), -1000000.0f), 1000000.0f); }

__global__ void pack_xyz_kernel(const unsigned char* __restrict__ raw, int n, int stride, float4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    out[i] = make_float4(p[0], p[1], p[2], __int_as_float(i));
}

__global__ void grid_insert_kernel(const float4* __restrict__ pts, int n, float inv_cell, unsigned mask,
                                   unsigned long long* __restrict__ keys, int* __restrict__ count, int* __restrict__ slot_of)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    const unsigned long long key = cell_key(clampi(p.x * inv_cell), clampi(p.y * inv_cell), clampi(p.z * inv_cell));
    unsigned s = hash_key(key) & mask;
    while (true) {
        const unsigned long long prev = atomicCAS(&keys[s], kEmpty, key);
        if (prev == kEmpty || prev == key) break;
        s = (s + 1) & mask;
    }
    atomicAdd(&count[s], 1);
    slot_of[i] = (int)s;
}

// exclusive scan of count[0..cap) into start[0..cap) by one CTA (cap <= a few million)
__global__ void __launch_bounds__(1024) grid_scan_kernel(const int* __restrict__ count, int cap, int* __restrict__ start)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < cap; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < cap ? count[i] : 0;
        int x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, off); if (lane >= off) w += y; }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int prefix = carry + (warp ? warp_sums[warp - 1] : 0) + x - v;
        if (i < cap) start[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
}

__global__ void grid_scatter_kernel(const float4* __restrict__ pts, int n, const int* __restrict__ slot_of, const int* __restrict__ start,
                                    int* __restrict__ fill, float4* __restrict__ sorted)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slot_of[i];
    sorted[start[s] + atomicAdd(&fill[s], 1)] = pts[i];
}

struct Best { float d2; int idx; };

__device__ __forceinline__ void visit_cell(const Grid& g, int i, int j, int k, float qx, float qy, float qz, Best& b)
{
    const unsigned long long key = cell_key(i, j, k);
    unsigned s = hash_key(key) & g.mask;
    while (true) {
        const unsigned long long cur = __ldg(&g.keys[s]);
        if (cur == kEmpty) return;
        if (cur == key) break;
        s = (s + 1) & g.mask;
    }
    const int st = __ldg(&g.start[s]), cn = __ldg(&g.count[s]);
    for (int t = 0; t < cn; t++) {
        const float4 p = __ldg(&g.pts[st + t]);
        const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        const int id = __float_as_int(p.w);
        if (d2 < b.d2 || (d2 == b.d2 && id < b.idx)) { b.d2 = d2; b.idx = id; }
    }
}

// shells 0..max_shell of grid g; returns true once the best is provably the nearest
__device__ bool search_grid(const Grid& g, float qx, float qy, float qz, int max_shell, Best& b)
{
    const int ci = clampi(qx * g.inv_cell), cj = clampi(qy * g.inv_cell), ck = clampi(qz * g.inv_cell);
    for (int r = 0; r <= max_shell; r++) {
        for (int i = ci - r; i <= ci + r; i++)
            for (int j = cj - r; j <= cj + r; j++) {
                const bool edge = (i == ci - r) || (i == ci + r) || (j == cj - r) || (j == cj + r);
                if (edge) { for (int k = ck - r; k <= ck + r; k++) visit_cell(g, i, j, k, qx, qy, qz, b); }
                else { visit_cell(g, i, j, ck - r, qx, qy, qz, b); if (r > 0) visit_cell(g, i, j, ck + r, qx, qy, qz, b); }
            }
        /* every unvisited cell is at Chebyshev distance >= r+1, hence farther than r*cell */
        const float bound = (float)r * g.cell;
        if (b.idx >= 0 && b.d2 <= bound * bound) return true;
    }
    return false;
}

__device__ Best nearest(const Grid& fine, const Grid& coarse, const float4* __restrict__ tgt, int n_tgt, float qx, float qy, float qz)
{
    Best b{3.0e38f, -1};
    if (search_grid(fine, qx, qy, qz, 2, b)) return b;
    if (search_grid(coarse, qx, qy, qz, 2, b)) return b;
    for (int t = 0; t < n_tgt; t++) {                      /* isolated point: exact full scan */
        const float4 p = __ldg(&tgt[t]);
        const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        if (d2 < b.d2 || (d2 == b.d2 && t < b.idx)) { b.d2 = d2; b.idx = t; }
    }
    return b;
}

constexpr int kAcc = 29;   /* 21 (upper JtJ) + 6 (Jt r) + sum d2 + count */

// One ICP iteration: transform, exact NN, warp-reduced normal equations. T is row-major 3x4.
__global__ void __launch_bounds__(256) icp_iter_kernel(const float4* __restrict__ src, int n_src, const float4* __restrict__ tgt, int n_tgt,
                                                       Grid fine, Grid coarse, const float* __restrict__ T, float max_d2,
                                                       double* __restrict__ acc /* [kAcc] */, int* __restrict__ nn_idx, float* __restrict__ nn_d2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double v[kAcc];
#pragma unroll
    for (int a = 0; a < kAcc; a++) v[a] = 0.0;
    if (i < n_src) {
        const float4 s = src[i];
        const float px = fmaf(T[0], s.x, fmaf(T[1], s.y, fmaf(T[2], s.z, T[3])));
        const float py = fmaf(T[4], s.x, fmaf(T[5], s.y, fmaf(T[6], s.z, T[7])));
        const float pz = fmaf(T[8], s.x, fmaf(T[9], s.y, fmaf(T[10], s.z, T[11])));
        const Best b = nearest(fine, coarse, tgt, n_tgt, px, py, pz);
        if (nn_idx) { nn_idx[i] = b.idx; nn_d2[i] = b.d2; }
        if (b.idx >= 0 && b.d2 <= max_d2) {
            const float4 q = __ldg(&tgt[b.idx]);
            const double x = px, y = py, z = pz;
            const double rx = x - q.x, ry = y - q.y, rz = z - q.z;
            /* J = [ -[p]x | I ], rows: (0, z,-y, 1,0,0), (-z, 0, x, 0,1,0), (y,-x, 0, 0,0,1) */
            const double J[3][6] = {{0, z, -y, 1, 0, 0}, {-z, 0, x, 0, 1, 0}, {y, -x, 0, 0, 0, 1}};
            const double r[3] = {rx, ry, rz};
            int o = 0;
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int c = a; c < 6; c++) v[o++] = J[0][a] * J[0][c] + J[1][a] * J[1][c] + J[2][a] * J[2][c];
#pragma unroll
            for (int a = 0; a < 6; a++) v[21 + a] = J[0][a] * r[0] + J[1][a] * r[1] + J[2][a] * r[2];
            v[27] = (double)b.d2;
            v[28] = 1.0;
        }
    }
#pragma unroll
    for (int a = 0; a < kAcc; a++) {
        double x = v[a];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(&acc[a], x);
    }
}

struct GridBufs { DevBuf* keys; DevBuf* start; DevBuf* count; DevBuf* sorted; DevBuf* slot; };

int build_grid(scl_engine* e, const float4* pts, int n, float cell, GridBufs gb, Grid* out)
{
    unsigned cap = 1024;
    while (cap < 2u * (unsigned)n) cap <<= 1;
    CK(gb.keys->ensure((size_t)cap * 8));
    CK(gb.start->ensure((size_t)cap * 4));
    CK(gb.count->ensure((size_t)cap * 8));          /* count[cap] + fill[cap] */
    CK(gb.sorted->ensure((size_t)n * 16));
    CK(gb.slot->ensure((size_t)n * 4));
    CK(cudaMemsetAsync(gb.keys->p, 0xff, (size_t)cap * 8, e->stream));
    CK(cudaMemsetAsync(gb.count->p, 0, (size_t)cap * 8, e->stream));
    int* count = gb.count->as<int>(); int* fill = count + cap;
    const int blocks = (n + 255) / 256;
    grid_insert_kernel<<<blocks, 256, 0, e->stream>>>(pts, n, 1.0f / cell, cap - 1, gb.keys->as<unsigned long long>(), count, gb.slot->as<int>());
    grid_scan_kernel<<<1, 1024, 0, e->stream>>>(count, (int)cap, gb.start->as<int>());
    grid_scatter_kernel<<<blocks, 256, 0, e->stream>>>(pts, n, gb.slot->as<int>(), gb.start->as<int>(), fill, gb.sorted->as<float4>());
    CK(cudaGetLastError());
    out->keys = gb.keys->as<unsigned long long>(); out->start = gb.start->as<int>(); out->count = count;
    out->pts = gb.sorted->as<float4>(); out->mask = cap - 1; out->inv_cell = 1.0f / cell; out->cell = cell;
    return SCL_OK;
}

bool cholesky6(double A[6][6], double b[6], double x[6])
{
    double L[6][6] = {{0}};
    for (int i = 0; i < 6; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[i][j];
            for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
            if (i == j) { if (!(s > 1e-12)) return false; L[i][i] = std::sqrt(s); }
            else L[i][j] = s / L[j][j];
        }
    double y[6];
    for (int i = 0; i < 6; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
    for (int i = 5; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < 6; k++) s -= L[k][i] * x[k]; x[i] = s / L[i][i]; }
    return true;
}

// rigid increment from a twist (omega, v): R = exp([omega]x), t = v
void twist_to_T(const double x[6], double T[16])
{
    const double wx = x[0], wy = x[1], wz = x[2];
    const double th = std::sqrt(wx * wx + wy * wy + wz * wz);
    double a, b;                     /* R = I + a K + b K^2 */
    if (th < 1e-9) { a = 1.0; b = 0.5; } else { a = std::sin(th) / th; b = (1.0 - std::cos(th)) / (th * th); }
    const double K[3][3] = {{0, -wz, wy}, {wz, 0, -wx}, {-wy, wx, 0}};
    double K2[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { K2[i][j] = 0; for (int k = 0; k < 3; k++) K2[i][j] += K[i][k] * K[k][j]; }
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) T[i * 4 + j] = (i == j) + a * K[i][j] + b * K2[i][j];
        T[i * 4 + 3] = x[3 + i];
    }
    T[12] = T[13] = T[14] = 0; T[15] = 1;
}

void mul4(const double A[16], const double B[16], double C[16])
{
    double t[16];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { double s = 0; for (int k = 0; k < 4; k++) s += A[r * 4 + k] * B[k * 4 + c]; t[r * 4 + c] = s; }
    for (int i = 0; i < 16; i++) C[i] = t[i];
}

// ---- closed-form rigid fit (Horn 1987): rotation = eigenvector of the largest eigenvalue of a symmetric 4x4 built
// from the cross-covariance S = sum (a - ca)(b - cb)^T; cyclic Jacobi. Used on the device for the 3-point RANSAC
// hypotheses and on the host for the final fit on the inliers (pcl::registration::TransformationEstimationSVD,
// distributedMapping.h:1228-1230, has the same optimum).
__host__ __device__ inline void horn_fit(const double S[9], const double ca[3], const double cb[3], double T[12])
{
    double N[4][4] = {
        {S[0] + S[4] + S[8], S[5] - S[7], S[6] - S[2], S[1] - S[3]},
        {S[5] - S[7], S[0] - S[4] - S[8], S[1] + S[3], S[6] + S[2]},
        {S[6] - S[2], S[1] + S[3], -S[0] + S[4] - S[8], S[5] + S[7]},
        {S[1] - S[3], S[6] + S[2], S[5] + S[7], -S[0] - S[4] + S[8]}};
    double V[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 24; sweep++) {
        double off = 0;
        for (int i = 0; i < 4; i++) for (int j = i + 1; j < 4; j++) off += N[i][j] * N[i][j];
        if (off < 1e-280) break;
        for (int p = 0; p < 4; p++)
            for (int q = p + 1; q < 4; q++) {
                if (fabs(N[p][q]) < 1e-300) continue;
                const double theta = (N[q][q] - N[p][p]) / (2 * N[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(t * t + 1), sn = t * c;
                for (int k = 0; k < 4; k++) { const double akp = N[k][p], akq = N[k][q]; N[k][p] = c * akp - sn * akq; N[k][q] = sn * akp + c * akq; }
                for (int k = 0; k < 4; k++) { const double apk = N[p][k], aqk = N[q][k]; N[p][k] = c * apk - sn * aqk; N[q][k] = sn * apk + c * aqk; }
                for (int k = 0; k < 4; k++) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - sn * vkq; V[k][q] = sn * vkp + c * vkq; }
            }
    }
    int best = 0;
    for (int i = 1; i < 4; i++) if (N[i][i] > N[best][best]) best = i;
    double qw = V[0][best], qx = V[1][best], qy = V[2][best], qz = V[3][best];
    const double nq = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
    qw /= nq; qx /= nq; qy /= nq; qz /= nq;
    const double R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw),
                         2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw),
                         2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)};
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) T[r * 4 + c] = R[r * 3 + c];
        T[r * 4 + 3] = cb[r] - (R[r * 3] * ca[0] + R[r * 3 + 1] * ca[1] + R[r * 3 + 2] * ca[2]);
    }
}

__device__ inline unsigned rng_hash(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// One RANSAC hypothesis per CTA (pcl::registration::CorrespondenceRejectorSampleConsensus, distributedMapping.h:1217-1225):
// three distinct correspondences -> rigid transform -> inlier count over all correspondences.
__global__ void __launch_bounds__(128) ransac_hyp_kernel(const float4* __restrict__ src, const float4* __restrict__ tgt, const int* __restrict__ nn,
                                                         int n, unsigned seed, float thr2, float* __restrict__ hyp_T /* [H][12] */,
                                                         int* __restrict__ hyp_inliers)
{
    __shared__ float sT[12];
    __shared__ int s_cnt;
    const int h = blockIdx.x;
    if (threadIdx.x == 0) {
        s_cnt = 0;
        int pick[3];
        unsigned st = rng_hash(seed ^ (0x9e3779b9u * (unsigned)(h + 1)));
        for (int k = 0; k < 3; k++) {
            while (true) {
                st = rng_hash(st + 0x632be5abu);
                const int c = (int)(st % (unsigned)n);
                bool dup = false;
                for (int j = 0; j < k; j++) dup |= pick[j] == c;
                if (!dup) { pick[k] = c; break; }
            }
        }
        double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0}, a[3][3], b[3][3];
        for (int k = 0; k < 3; k++) {
            const float4 p = src[pick[k]], q = tgt[nn[pick[k]]];
            a[k][0] = p.x; a[k][1] = p.y; a[k][2] = p.z; b[k][0] = q.x; b[k][1] = q.y; b[k][2] = q.z;
            for (int d = 0; d < 3; d++) { ca[d] += a[k][d] / 3.0; cb[d] += b[k][d] / 3.0; }
        }
        double S[9] = {0};
        for (int k = 0; k < 3; k++)
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) S[r * 3 + c] += (a[k][r] - ca[r]) * (b[k][c] - cb[c]);
        double T[12];
        horn_fit(S, ca, cb, T);
        for (int i = 0; i < 12; i++) { sT[i] = (float)T[i]; hyp_T[(size_t)h * 12 + i] = (float)T[i]; }
    }
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float4 p = src[i], q = __ldg(&tgt[nn[i]]);
        const float x = fmaf(sT[0], p.x, fmaf(sT[1], p.y, fmaf(sT[2], p.z, sT[3]))) - q.x;
        const float y = fmaf(sT[4], p.x, fmaf(sT[5], p.y, fmaf(sT[6], p.z, sT[7]))) - q.y;
        const float z = fmaf(sT[8], p.x, fmaf(sT[9], p.y, fmaf(sT[10], p.z, sT[11]))) - q.z;
        cnt += (fmaf(x, x, fmaf(y, y, z * z)) <= thr2) ? 1 : 0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) hyp_inliers[h] = s_cnt;
}

// centroids + cross-covariance sums over the inliers of transform T (15 FP64 sums + count), warp-reduced
__global__ void __launch_bounds__(256) inlier_stats_kernel(const float4* __restrict__ src, const float4* __restrict__ tgt, const int* __restrict__ nn,
                                                           int n, const float* __restrict__ T, float thr2, double* __restrict__ acc /* [16] */)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double v[16];
#pragma unroll
    for (int a = 0; a < 16; a++) v[a] = 0.0;
    if (i < n) {
        const float4 p = src[i], q = __ldg(&tgt[nn[i]]);
        const float x = fmaf(T[0], p.x, fmaf(T[1], p.y, fmaf(T[2], p.z, T[3]))) - q.x;
        const float y = fmaf(T[4], p.x, fmaf(T[5], p.y, fmaf(T[6], p.z, T[7]))) - q.y;
        const float z = fmaf(T[8], p.x, fmaf(T[9], p.y, fmaf(T[10], p.z, T[11]))) - q.z;
        if (fmaf(x, x, fmaf(y, y, z * z)) <= thr2) {
            const double a[3] = {p.x, p.y, p.z}, b[3] = {q.x, q.y, q.z};
            for (int d = 0; d < 3; d++) { v[d] = a[d]; v[3 + d] = b[d]; }
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) v[6 + r * 3 + c] = a[r] * b[c];
            v[15] = 1.0;
        }
    }
#pragma unroll
    for (int a = 0; a < 16; a++) {
        double x = v[a];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(&acc[a], x);
    }
}

} // namespace

extern "C" int scl_icp(scl_engine* e, const void* src, int n_src, const void* tgt, int n_tgt, int stride_bytes,
                       const scl_icp_params* prm, float* T_out, float* fitness, int* converged, int* iterations)
{
    LOCK();
    if (!prm || !T_out || !fitness || !converged) FAIL(SCL_ERR_INVALID, "null argument");
    if (n_src < 0 || n_tgt < 0 || stride_bytes < 12 || (stride_bytes & 3)) FAIL(SCL_ERR_INVALID, "bad cloud arguments");
    if (n_src == 0 || n_tgt == 0) return scl_icp_device(e, n_src, n_tgt, prm, T_out, fitness, converged, iterations);
    if (!src || !tgt) FAIL(SCL_ERR_INVALID, "null cloud");

    /* upload + pack to float4 */
    const size_t sb = (size_t)(n_src - 1) * stride_bytes + 12, tb = (size_t)(n_tgt - 1) * stride_bytes + 12;
    CK(e->icp_raw.ensure((sb > tb ? sb : tb) + 16));
    CK(e->icp_src.ensure((size_t)n_src * 16));
    CK(e->icp_tgt.ensure((size_t)n_tgt * 16));
    CK(cudaMemcpyAsync(e->icp_raw.p, src, sb, cudaMemcpyHostToDevice, e->stream));
    pack_xyz_kernel<<<(n_src + 255) / 256, 256, 0, e->stream>>>(e->icp_raw.as<unsigned char>(), n_src, stride_bytes, e->icp_src.as<float4>());
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpyAsync(e->icp_raw.p, tgt, tb, cudaMemcpyHostToDevice, e->stream));
    pack_xyz_kernel<<<(n_tgt + 255) / 256, 256, 0, e->stream>>>(e->icp_raw.as<unsigned char>(), n_tgt, stride_bytes, e->icp_tgt.as<float4>());
    CK(cudaGetLastError());
    return scl_icp_device(e, n_src, n_tgt, prm, T_out, fitness, converged, iterations);
}

// A cloud that is already in device memory (packed 16-byte x, y, z, * records) becomes the ICP source (which = 0) or target
// (1): same layout as the host path produces (w = the point's index). The caller holds the engine lock.
int scl_icp_set_cloud_dev(scl_engine* e, int which, const void* xyzw_dev, int n)
{
    DevBuf& dst = which == 0 ? e->icp_src : e->icp_tgt;
    if (n <= 0) return SCL_OK;
    CK(dst.ensure((size_t)n * 16));
    pack_xyz_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>(static_cast<const unsigned char*>(xyzw_dev), n, 16, dst.as<float4>());
    CK(cudaGetLastError());
    return SCL_OK;
}

// ICP of the float4 clouds already in e->icp_src / e->icp_tgt (device memory); the caller holds the engine lock.
int scl_icp_device(scl_engine* e, int n_src, int n_tgt, const scl_icp_params* prm, float* T_out, float* fitness, int* converged, int* iterations)
{
    double Tf[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    for (int i = 0; i < 16; i++) T_out[i] = (float)Tf[i];
    *fitness = 3.402823466e+38f; *converged = 0;
    if (iterations) *iterations = 0;
    if (n_src == 0 || n_tgt == 0) return SCL_OK;       /* PCL: nothing to align, not converged */

    Grid fine, coarse;
    int rc = build_grid(e, e->icp_tgt.as<float4>(), n_tgt, 1.0f,
                        GridBufs{&e->icp_grid[0][0], &e->icp_grid[0][1], &e->icp_grid[0][2], &e->icp_grid[0][3], &e->icp_grid[0][4]}, &fine);
    if (rc) return rc;
    rc = build_grid(e, e->icp_tgt.as<float4>(), n_tgt, 8.0f,
                    GridBufs{&e->icp_grid[1][0], &e->icp_grid[1][1], &e->icp_grid[1][2], &e->icp_grid[1][3], &e->icp_grid[1][4]}, &coarse);
    if (rc) return rc;

    CK(e->icp_acc.ensure(kAcc * 8 + 12 * 4));
    double* d_acc = e->icp_acc.as<double>();
    float* d_T = reinterpret_cast<float*>(d_acc + kAcc);
    const float max_d2 = (float)(prm->max_corr_dist * prm->max_corr_dist);
    const int blocks = (n_src + 255) / 256;
    /* pcl::IterativeClosestPoint::computeTransformation hands its settings to DefaultConvergenceCriteria like this (icp.hpp,
     * PCL 1.8 - 1.12): setRelativeMSE(euclidean_fitness_epsilon_), setTranslationThreshold(transformation_epsilon_),
     * setRotationThreshold(1 - transformation_epsilon_); the absolute-MSE threshold keeps its default 1e-12 and
     * max_iterations_similar_transforms_ its default 0 (the first similar iteration ends the loop). */
    const double rotation_threshold = 1.0 - prm->trans_eps, mse_rel = prm->fitness_eps, mse_abs = 1e-12;
    double prev_mse = 1.7976931348623157e308;
    int it = 0; bool conv = false;
    double acc[kAcc];

    auto run_iter = [&](const double T[16]) -> int {
        float Tf32[12];
        for (int i = 0; i < 12; i++) Tf32[i] = (float)T[i];
        CK(cudaMemcpyAsync(d_T, Tf32, sizeof(Tf32), cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemsetAsync(d_acc, 0, kAcc * 8, e->stream));
        icp_iter_kernel<<<blocks, 256, 0, e->stream>>>(e->icp_src.as<float4>(), n_src, e->icp_tgt.as<float4>(), n_tgt, fine, coarse, d_T,
                                                       max_d2, d_acc, nullptr, nullptr);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(acc, d_acc, kAcc * 8, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        return SCL_OK;
    };

    while (true) {
        rc = run_iter(Tf); if (rc) return rc;
        const double ncorr = acc[28];
        if (ncorr < 3.0) { conv = false; break; }                 /* min_number_correspondences_ */
        double A[6][6], b[6], x[6];
        int o = 0;
        for (int a = 0; a < 6; a++) for (int c = a; c < 6; c++) { A[a][c] = acc[o]; A[c][a] = acc[o]; o++; }
        for (int a = 0; a < 6; a++) b[a] = -acc[21 + a];
        if (!cholesky6(A, b, x)) { conv = false; break; }
        double dT[16];
        twist_to_T(x, dT);
        mul4(dT, Tf, Tf);
        ++it;
        if (it >= prm->max_iterations) { conv = true; break; }
        const double cos_angle = 0.5 * (dT[0] + dT[5] + dT[10] - 1.0);
        const double translation_sqr = dT[3] * dT[3] + dT[7] * dT[7] + dT[11] * dT[11];
        if (cos_angle >= rotation_threshold && translation_sqr <= prm->trans_eps) { conv = true; break; }
        const double cur_mse = acc[27] / ncorr;
        if (std::fabs(cur_mse - prev_mse) < mse_abs) { conv = true; break; }
        if (std::fabs(cur_mse - prev_mse) / prev_mse < mse_rel) { conv = true; break; }
        prev_mse = cur_mse;
    }
    /* getFitnessScore(): mean squared NN distance of the aligned source, no distance gate */
    const float save = max_d2; (void)save;
    {
        float Tf32[12];
        for (int i = 0; i < 12; i++) Tf32[i] = (float)Tf[i];
        CK(cudaMemcpyAsync(d_T, Tf32, sizeof(Tf32), cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemsetAsync(d_acc, 0, kAcc * 8, e->stream));
        icp_iter_kernel<<<blocks, 256, 0, e->stream>>>(e->icp_src.as<float4>(), n_src, e->icp_tgt.as<float4>(), n_tgt, fine, coarse, d_T,
                                                       3.0e38f, d_acc, nullptr, nullptr);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(acc, d_acc, kAcc * 8, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    if (acc[28] > 0) *fitness = (float)(acc[27] / acc[28]);
    for (int i = 0; i < 16; i++) T_out[i] = (float)Tf[i];
    *converged = conv ? 1 : 0;
    if (iterations) *iterations = it;
    return SCL_OK;
}

extern "C" void scl_default_ransac_params(scl_ransac_params* p)
{
    p->max_iterations = 1000; p->inlier_threshold = 0.25; p->min_inlier_ratio = 0.45; p->seed = 1u;   /* distributedMapping.h:187-189 */
}

extern "C" int scl_verify_ransac(scl_engine* e, const void* src, int n_src, const void* tgt, int n_tgt, int stride_bytes,
                                 const scl_ransac_params* prm, float* T_out, int* n_corr, int* n_inliers, int* success)
{
    LOCK();
    if (!prm || !T_out || !n_corr || !n_inliers || !success) FAIL(SCL_ERR_INVALID, "null argument");
    if (n_src < 0 || n_tgt < 0 || stride_bytes < 12 || (stride_bytes & 3) || prm->max_iterations < 1) FAIL(SCL_ERR_INVALID, "bad arguments");
    for (int i = 0; i < 16; i++) T_out[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    *n_corr = 0; *n_inliers = 0; *success = 0;
    if (n_src < 3 || n_tgt < 1) return SCL_OK;
    if (!src || !tgt) FAIL(SCL_ERR_INVALID, "null cloud");
    const size_t sb = (size_t)(n_src - 1) * stride_bytes + 12, tb = (size_t)(n_tgt - 1) * stride_bytes + 12;
    CK(e->icp_raw.ensure((sb > tb ? sb : tb) + 16));
    CK(e->icp_src.ensure((size_t)n_src * 16));
    CK(e->icp_tgt.ensure((size_t)n_tgt * 16));
    CK(cudaMemcpyAsync(e->icp_raw.p, src, sb, cudaMemcpyHostToDevice, e->stream));
    pack_xyz_kernel<<<(n_src + 255) / 256, 256, 0, e->stream>>>(e->icp_raw.as<unsigned char>(), n_src, stride_bytes, e->icp_src.as<float4>());
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpyAsync(e->icp_raw.p, tgt, tb, cudaMemcpyHostToDevice, e->stream));
    pack_xyz_kernel<<<(n_tgt + 255) / 256, 256, 0, e->stream>>>(e->icp_raw.as<unsigned char>(), n_tgt, stride_bytes, e->icp_tgt.as<float4>());
    CK(cudaGetLastError());
    Grid fine, coarse;
    int rc = build_grid(e, e->icp_tgt.as<float4>(), n_tgt, 1.0f,
                        GridBufs{&e->icp_grid[0][0], &e->icp_grid[0][1], &e->icp_grid[0][2], &e->icp_grid[0][3], &e->icp_grid[0][4]}, &fine);
    if (rc) return rc;
    rc = build_grid(e, e->icp_tgt.as<float4>(), n_tgt, 8.0f,
                    GridBufs{&e->icp_grid[1][0], &e->icp_grid[1][1], &e->icp_grid[1][2], &e->icp_grid[1][3], &e->icp_grid[1][4]}, &coarse);
    if (rc) return rc;
    /* initial matching (:1211-1215): nearest target point of every source point, no distance gate */
    const int H = prm->max_iterations;
    CK(e->icp_nn.ensure((size_t)n_src * 8 + (size_t)H * 13 * 4));
    int* d_nn = e->icp_nn.as<int>();
    float* d_nnd2 = reinterpret_cast<float*>(d_nn + n_src);
    float* d_hypT = d_nnd2 + n_src;
    int* d_hypI = reinterpret_cast<int*>(d_hypT + (size_t)H * 12);
    CK(e->icp_acc.ensure(kAcc * 8 + 12 * 4));
    double* d_acc = e->icp_acc.as<double>();
    float* d_T = reinterpret_cast<float*>(d_acc + kAcc);
    const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    CK(cudaMemcpyAsync(d_T, ident, sizeof(ident), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(d_acc, 0, kAcc * 8, e->stream));
    icp_iter_kernel<<<(n_src + 255) / 256, 256, 0, e->stream>>>(e->icp_src.as<float4>(), n_src, e->icp_tgt.as<float4>(), n_tgt, fine, coarse, d_T,
                                                                 3.0e38f, d_acc, d_nn, d_nnd2);
    CK(cudaGetLastError());
    /* RANSAC (:1217-1225): H three-point hypotheses, inlier threshold on the residual after the hypothesis */
    const float thr2 = (float)(prm->inlier_threshold * prm->inlier_threshold);
    ransac_hyp_kernel<<<H, 128, 0, e->stream>>>(e->icp_src.as<float4>(), e->icp_tgt.as<float4>(), d_nn, n_src, prm->seed, thr2, d_hypT, d_hypI);
    CK(cudaGetLastError());
    std::vector<int> inl(H);
    CK(cudaMemcpyAsync(inl.data(), d_hypI, (size_t)H * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    int best = 0;
    for (int h = 1; h < H; h++) if (inl[h] > inl[best]) best = h;
    *n_corr = n_src; *n_inliers = inl[best];
    /* SVD on the inliers of the best hypothesis (:1228-1230) */
    CK(cudaMemcpyAsync(d_T, d_hypT + (size_t)best * 12, 12 * 4, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemsetAsync(d_acc, 0, kAcc * 8, e->stream));
    inlier_stats_kernel<<<(n_src + 255) / 256, 256, 0, e->stream>>>(e->icp_src.as<float4>(), e->icp_tgt.as<float4>(), d_nn, n_src, d_T, thr2, d_acc);
    CK(cudaGetLastError());
    double acc[16];
    CK(cudaMemcpyAsync(acc, d_acc, sizeof(acc), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (acc[15] >= 3.0) {
        const double m = acc[15];
        const double ca[3] = {acc[0] / m, acc[1] / m, acc[2] / m}, cb[3] = {acc[3] / m, acc[4] / m, acc[5] / m};
        double S[9];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) S[r * 3 + c] = acc[6 + r * 3 + c] - m * ca[r] * cb[c];
        double T[12];
        horn_fit(S, ca, cb, T);
        for (int i = 0; i < 12; i++) T_out[i] = (float)T[i];
    }
    *success = ((double)*n_inliers >= prm->min_inlier_ratio * (double)*n_corr) ? 1 : 0;   /* :1238 */
    return SCL_OK;
}
