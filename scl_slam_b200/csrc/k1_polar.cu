// k1_polar.cu — K1 polar binning (+ K2 ring keys in its epilogue) and the stand-alone K2 kernel.
//
// Replaces makeScancontext (/root/reference/include/descriptor.h:1404-1461), xy2theta
// (:1352-1374), makeRingkeyFromScancontext (:1463-1475) and the database append of save()
// (:1587-1599).
//
// Layout: clouds arrive as the caller's array-of-structs (pcl::PointXYZI, 32 B/point, x y z at
// byte 0 4 8); with 16-byte aligned points one LDG.128 fetches x,y,z,pad per point. A batch of
// scans is one launch: grid = (chunks per scan, scans), two full waves of resident CTAs; the
// offsets of up to 256 scans travel as kernel parameters. Each CTA bins its chunks into a
// shared-memory R x S array of order-preserving uint keys with a warp-aggregated atomicMax
// (lanes that hit the same bin are first reduced with __match_any_sync/__reduce_max_sync, so
// one ATOMS per distinct bin per warp), merges that array into the scan's global bin array with
// atomicMax, and the LAST CTA of the scan (ticket counter) decodes the bins, applies the
// "-1000 -> 0" rule, writes the R*S float wire image straight into the database slot (or the
// caller's buffer), reduces the ring key and writes the per-entry column statistics K4 reads.
// Because max is order-free and heights are plain floats, bin contents are bit-exact whatever
// the arrival order.
//
// Per point the kernel does one float division and two fixed-depth table searches: no atanf, no
// square root, no double-precision index arithmetic (see "exact bin tables" below); 130
// instructions per point in the production instantiation.
//
// Roofline: HBM streaming, 32 B/point read in place (16 B/point algorithmic for packed input),
// R*S*4 + R*4 B written per scan. Measured (profiles/r02i_*): 62 us per 64 scans of 113 k
// points = 3.7 TB/s in place (57 % of the HBM peak); bound by instruction issue at five CTAs/SM.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <vector>

namespace {

constexpr float kNoPoint = -1000.0f;   /* descriptor.h:1411 */

// ---- exact bin tables -------------------------------------------------------------------------
// The reference computes ring = clamp(ceil(double(r)/max_radius*R), 1, R) and
// sector = clamp(ceil(double(theta)/360*S), 1, S) with theta = float(180/pi * atanf(y/x)) folded by quadrant
// (descriptor.h:1352-1374,1434-1435). Both are MONOTONE step functions of one float — r, and within a
// quadrant the ratio t = |y|/|x| (libm's atanf is monotone on [0, inf]: checked exhaustively).
// So the host evaluates the reference formula (same double arithmetic, same atanf algorithm) only to find,
// by bisection over float bit patterns, the exact floats at which each function steps; the kernel then needs
// one float division and two binary searches per point — no atanf, no FP64 division — and is bit-exact by
// construction, including at the sector boundaries.
// The ring table goes one step further back: r = sqrtf(s) with s = fl(fl(x*x) + fl(y*y)) (:1425), and a correctly
// rounded square root is monotone, so the ring is a monotone step function of s as well. The thresholds are
// bisected in s (the host takes the correctly rounded sqrtf of every probe), and the kernel has no square root.
struct BinTables {
    int n_ring;            /* ring thresholds: ring = 1 + #(thr <= s), s = fl(fl(x*x) + fl(y*y)) */
    int n_sec[4];          /* per quadrant: sector = base + dir * #(thr <= t) */
    int sec_base[4], sec_dir[4], sec_off[4];
    float s_max;           /* largest float s with double(sqrtf(s)) <= max_radius */
};

inline uint32_t h_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float h_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// libm's float atan restated (fdlibm s_atanf, bit-identical to glibc 2.39 on all 2^32 inputs — see oracle/sc_oracle.cpp)
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
float h_atanf(float x)
{
    static const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    static const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    static const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f,
                                 -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f,
                                 1.6285819933e-02f};
    const uint32_t hx = h_f2u(x), ix = hx & 0x7fffffffu;
    int id;
    volatile float t1, t2;   /* volatile temporaries: no host-side contraction or reassociation */
    if (ix >= 0x4c000000u) { if (ix > 0x7f800000u) return x + x; const float r = atanhi[3] + atanlo[3]; return (hx >> 31) ? -r : r; }
    if (ix < 0x3ee00000u) { if (ix < 0x31000000u) return x; id = -1; }
    else {
        x = fabsf(x);
        if (ix < 0x3f980000u) {
            if (ix < 0x3f300000u) { id = 0; t1 = 2.0f * x; t1 = t1 - 1.0f; t2 = 2.0f + x; x = t1 / t2; }
            else { id = 1; t1 = x - 1.0f; t2 = x + 1.0f; x = t1 / t2; }
        } else {
            if (ix < 0x401c0000u) { id = 2; t1 = x - 1.5f; t2 = 1.5f * x; t2 = 1.0f + t2; x = t1 / t2; }
            else { id = 3; x = -1.0f / x; }
        }
    }
    volatile float z = x * x, w = z * z, p;
    p = w * aT[10]; p = aT[8] + p; p = w * p; p = aT[6] + p; p = w * p; p = aT[4] + p; p = w * p; p = aT[2] + p; p = w * p; p = aT[0] + p;
    volatile float s1 = z * p;
    p = w * aT[9]; p = aT[7] + p; p = w * p; p = aT[5] + p; p = w * p; p = aT[3] + p; p = w * p; p = aT[1] + p;
    volatile float s2 = w * p;
    volatile float sum = s1 + s2, xs = x * sum;
    if (id < 0) return x - xs;
    volatile float u = xs - atanlo[id]; u = u - x;
    const float zz = atanhi[id] - u;
    return (hx >> 31) ? -zz : zz;
}

// correctly rounded float square root (IEEE 754; the reference's sqrt of a float sum, descriptor.h:1425, rounds the same
// way whether it goes through sqrtf or through the double sqrt: sqrt is immune to double rounding from 53 to 24 bits)
float h_sqrtf(float s) { volatile float v = s; volatile float r = std::sqrt(v); return r; }

int h_ceil_to_int(double v) { const double c = std::ceil(v); if (!(c >= -2147483648.0 && c <= 2147483647.0)) return (int)0x80000000; return (int)c; }

int h_ring_of(float r, int R, double max_radius)
{
    volatile double q = (double)r / max_radius; q = q * (double)R;
    return std::max(std::min(R, h_ceil_to_int(q)), 1);
}
// sector for quadrant qd and ratio t = |y|/|x| (the argument xy2theta hands to atan)
int h_sector_of(int qd, float t, int S)
{
    volatile double a = (180 / M_PI) * (double)h_atanf(t);
    volatile double th = qd == 0 ? a : qd == 1 ? 180 - a : qd == 2 ? 180 + a : 360 - a;
    const float theta = (float)th;
    volatile double q = (double)theta / 360.0; q = q * (double)S;
    return std::max(std::min(S, h_ceil_to_int(q)), 1);
}
// smallest non-negative float (by bit pattern, 0..+inf) at which pred becomes true; pred must be monotone false->true
template <typename Pred> bool h_first_true(Pred pred, float* out, float upper = std::numeric_limits<float>::infinity())
{
    uint32_t lo = 0, hi = h_f2u(upper);
    if (!pred(h_u2f(hi))) return false;
    while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (pred(h_u2f(mid))) hi = mid; else lo = mid + 1; }
    *out = h_u2f(lo);
    return true;
}

void build_tables(int R, int S, double max_radius, BinTables& bt, std::vector<float>& tab)
{
    tab.clear();
    float sm = 0.0f;
    h_first_true([&](float s) { return (double)h_sqrtf(s) > max_radius; }, &sm);   /* smallest s whose root lies beyond the radius (:1429) ... */
    bt.s_max = std::nextafterf(sm, 0.0f);                                          /* ... so this is the largest one inside */
    /* ring thresholds are searched inside [0, s_max] only (beyond it int(ceil()) overflows, and the point is dropped anyway) */
    for (int i = 1; i < R; i++) { float f; if (h_first_true([&](float s) { return h_ring_of(h_sqrtf(s), R, max_radius) > i; }, &f, bt.s_max)) tab.push_back(f); }
    bt.n_ring = (int)tab.size();
    for (int qd = 0; qd < 4; qd++) {
        const int v0 = h_sector_of(qd, 0.0f, S), vinf = h_sector_of(qd, std::numeric_limits<float>::infinity(), S);
        bt.sec_base[qd] = v0; bt.sec_dir[qd] = vinf >= v0 ? 1 : -1; bt.sec_off[qd] = (int)tab.size();
        if (vinf >= v0) {
            for (int v = v0 + 1; v <= vinf; v++) { float f; if (h_first_true([&](float t) { return h_sector_of(qd, t, S) >= v; }, &f)) tab.push_back(f); }
        } else {
            for (int v = v0 - 1; v >= vinf; v--) { float f; if (h_first_true([&](float t) { return h_sector_of(qd, t, S) <= v; }, &f)) tab.push_back(f); }
        }
        bt.n_sec[qd] = (int)tab.size() - bt.sec_off[qd];
    }
}

struct PolarParams {
    int R, S;
    double lidar_height;
    BinTables bt;
    const float* tab;      /* device copy of the threshold table */
    int n_tab;
};

// Shared-memory image of the tables: every table is padded with NaN (never <= anything) to a power of two minus one, so
// that the search is a fixed number of steps without a branch.
constexpr int kRingPad = 64, kSecPad = 32;
struct PolarSmem {
    float ring[kRingPad];
    float sec[4][kSecPad];
    int4 qd[4];            /* per quadrant: sector base, direction, -, - */
};

template <int TOP>       /* TOP = half the padded size: 32 -> 63 entries, 16 -> 31, 8 -> 15 */
__device__ __forceinline__ int count_le_fixed(const float* __restrict__ t, float v)
{
    int pos = 0;
#pragma unroll
    for (int step = TOP; step > 0; step >>= 1) pos += (t[pos + step - 1] <= v) ? step : 0;
    return pos;
}

// Per-point bin: one float division, two fixed-depth table searches, no square root, no divergent branch.
// s is the float sum the reference hands to sqrt (:1425); the ratio handed to the sector table is the one xy2theta
// hands to atan in each quadrant (y/x, y/-x, y/x, -y/x). RING_TOP / SEC_TOP: search depths (2 TOP - 1 table entries).
template <int RING_TOP, int SEC_TOP>
__device__ __forceinline__ bool polar_bin(const PolarSmem& ps, float s_max, float x, float y, float z, double lidar_height,
                                          int& ring, int& sector, float& zf)
{
    zf = __double2float_rn(__dadd_rn((double)z, lidar_height));                    /* :1422 */
    const float s = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));                   /* :1425, before the root */
    const bool xn = x < 0, yn = y < 0;
    const bool valid = (xn | (x >= 0)) & (yn | (y >= 0));      /* a NaN coordinate is UB in the reference: dropped (Q9) */
    const int qd = xn ? (yn ? 2 : 1) : (yn ? 3 : 0);
    const float num = qd == 3 ? -y : y, den = qd == 1 ? -x : x;
    const float t = __fdiv_rn(num, den);
    ring = 1 + count_le_fixed<RING_TOP>(ps.ring, s);
    const int4 q = ps.qd[qd];
    const int cnt = count_le_fixed<SEC_TOP>(ps.sec[qd], t);
    sector = (t != t) ? 1 : q.x + q.y * cnt;    /* (0,0): theta = NaN -> int(ceil(NaN)) = INT_MIN -> clamped to 1 */
    return valid & !(s > s_max);                /* double(sqrtf(s)) > max_radius (:1429) */
}

template <bool VEC4>
__device__ __forceinline__ void load_xyz(const unsigned char* base, size_t i, int stride, bool vec4, float& x, float& y, float& z)
{
    const unsigned char* q = base + i * (size_t)stride;
    if (VEC4 || vec4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(q));
        x = v.x; y = v.y; z = v.z;
    } else {
        const float* f = reinterpret_cast<const float*>(q);
        x = __ldg(f); y = __ldg(f + 1); z = __ldg(f + 2);
    }
}

// Per-entry cache of what distanceBtnScanContext recomputes for both operands of every pair (descriptor.h:1541-1542 ->
// makeSectorkeyFromScancontext :1477-1489; distDirectSC's column norms :1521-1522): cstat[0..S) = column means (the
// sector key), cstat[S..2S) = column Euclidean norms, both in double and in the reference's order (sequential over the
// rows). v * v is exact in double for a float v, so the fused multiply-add rounds exactly like mul-then-add.
// Worker `tid` of `nworkers` takes every nworkers-th column.
__device__ __forceinline__ void column_stats(const float* __restrict__ tile, int R, int S, int pitch, int tid, int nworkers, double* __restrict__ out)
{
    for (int c = tid; c < S; c += nworkers) {
        double s = 0.0, ss = 0.0;
        for (int r = 0; r < R; r++) {
            const double v = (double)tile[r * pitch + c];
            s = __dadd_rn(s, v);
            ss = __fma_rn(v, v, ss);
        }
        out[c] = __ddiv_rn(s, (double)R);
        out[S + c] = __dsqrt_rn(ss);
    }
}

// grid = (chunks, scans); block = kPolarThreads.
// FAST: 16-byte aligned points (one LDG.128 each) and no per-point bin output: the production path. The generic
// instantiation (any 4-byte aligned stride, optional per-point (ring, sector) for the parity tests) shares polar_bin.
constexpr int kPolarThreads = 256;
constexpr int kInlineScans = 256;          /* batches up to this many scans carry their offsets in the kernel parameters */
struct ScanOffsets { int v[kInlineScans + 1]; };
template <int kPointsPerThread, bool FAST, int RING_TOP, int SEC_TOP>
// (48 registers: five CTAs per SM. Forcing six with __launch_bounds__(256, 6) costs 12 bytes of spill and measured slower, 64.6 against 60.6 us.)
__global__ void __launch_bounds__(kPolarThreads) polar_bin_kernel(
    const unsigned char* __restrict__ pts, const int* __restrict__ offsets /* null: inl holds them */, const ScanOffsets inl, int stride, int vec4,
    PolarParams p, uint32_t* __restrict__ gbins /* [scans][R*S], zero between launches */,
    int* __restrict__ tickets /* [scans], zero between launches */,
    float* __restrict__ out_desc /* [scans][R*S] */, float* __restrict__ out_keys /* [scans][R] */,
    float* __restrict__ out_knorm /* [scans] */, float* __restrict__ kn2max, double* __restrict__ out_cstat /* [scans][2*S] or null */,
    int* __restrict__ out_ring, int* __restrict__ out_sector)
{
    extern __shared__ uint32_t sbins[];   /* R*S keys, then R*S + R floats for the epilogue */
    __shared__ int s_last;
    __shared__ PolarSmem ps;
    const int RS = p.R * p.S;
    {
        const float nan = __int_as_float(0x7fc00000);
        for (int i = threadIdx.x; i < kRingPad; i += kPolarThreads) ps.ring[i] = i < p.bt.n_ring ? __ldg(p.tab + i) : nan;
        for (int i = threadIdx.x; i < 4 * kSecPad; i += kPolarThreads) {
            const int qd = i / kSecPad, k = i % kSecPad;
            const int n = qd == 0 ? p.bt.n_sec[0] : qd == 1 ? p.bt.n_sec[1] : qd == 2 ? p.bt.n_sec[2] : p.bt.n_sec[3];
            const int off = qd == 0 ? p.bt.sec_off[0] : qd == 1 ? p.bt.sec_off[1] : qd == 2 ? p.bt.sec_off[2] : p.bt.sec_off[3];
            ps.sec[qd][k] = k < n ? __ldg(p.tab + off + k) : nan;
        }
        if (threadIdx.x < 4) {
            const int qd = threadIdx.x;
            const int base = qd == 0 ? p.bt.sec_base[0] : qd == 1 ? p.bt.sec_base[1] : qd == 2 ? p.bt.sec_base[2] : p.bt.sec_base[3];
            const int dir = qd == 0 ? p.bt.sec_dir[0] : qd == 1 ? p.bt.sec_dir[1] : qd == 2 ? p.bt.sec_dir[2] : p.bt.sec_dir[3];
            ps.qd[qd] = make_int4(base, dir, 0, 0);
        }
    }
    const double lidar_height = p.lidar_height;
    const float s_max = p.bt.s_max;
    const int S = p.S;
    const int scan = blockIdx.y;
    const int p0 = offsets ? offsets[scan] : inl.v[scan], p1 = offsets ? offsets[scan + 1] : inl.v[scan + 1];
    const uint32_t key_none = scl_float_key(kNoPoint);

    for (int i = threadIdx.x; i < RS; i += kPolarThreads) sbins[i] = 0u;
    __syncthreads();

    /* Software pipeline: the loads of the next batch of kPointsPerThread points per thread are in flight while the current
     * batch is binned (40 % of the stall samples of the unpipelined loop were the first use of a freshly loaded point). */
    constexpr int chunk = kPolarThreads * kPointsPerThread;
    const int step = gridDim.x * chunk;
    const float qnan = __int_as_float(0x7fc00000);
    const int lane = threadIdx.x & 31;
    float x[kPointsPerThread], y[kPointsPerThread], z[kPointsPerThread];
    float nx[kPointsPerThread], ny[kPointsPerThread], nz[kPointsPerThread];
    int base = p0 + blockIdx.x * chunk;              /* uniform over the CTA: the loop below holds warp-wide collectives */
#pragma unroll
    for (int j = 0; j < kPointsPerThread; j++) {
        const int i = base + j * kPolarThreads + threadIdx.x;
        if (i < p1) load_xyz<FAST>(pts, (size_t)i, stride, vec4 != 0, x[j], y[j], z[j]);
        else { x[j] = y[j] = z[j] = qnan; }
    }
    for (; base < p1; base += step) {
        const int nbase = base + step;
#pragma unroll
        for (int j = 0; j < kPointsPerThread; j++) {
            const int i = nbase + j * kPolarThreads + threadIdx.x;
            if (i < p1) load_xyz<FAST>(pts, (size_t)i, stride, vec4 != 0, nx[j], ny[j], nz[j]);
            else { nx[j] = ny[j] = nz[j] = qnan; }
        }
#pragma unroll
        for (int j = 0; j < kPointsPerThread; j++) {
            const int i = base + j * kPolarThreads + threadIdx.x;
            int ring = 0, sector = 0; float zf;
            /* a slot past the end of the scan holds NaN coordinates: not valid, never binned */
            bool ok = polar_bin<RING_TOP, SEC_TOP>(ps, s_max, x[j], y[j], z[j], lidar_height, ring, sector, zf);
            if (!FAST) {
                if (out_ring != nullptr && i < p1) { out_ring[i] = ok ? ring : 0; out_sector[i] = ok ? sector : 0; }
            }
            ok = ok && !(zf != zf);                     /* desc < NaN is false: NaN heights never win (:1438) */
            const int bin = ok ? (ring - 1) * S + (sector - 1) : -1;
            const uint32_t key = ok ? scl_float_key(zf) : 0u;
            /* warp-aggregated max: one shared atomic per distinct bin per warp */
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            const uint32_t m = __reduce_max_sync(peers, key);
            if (ok && (__ffs(peers) - 1) == lane) atomicMax(&sbins[bin], m);
        }
#pragma unroll
        for (int j = 0; j < kPointsPerThread; j++) { x[j] = nx[j]; y[j] = ny[j]; z[j] = nz[j]; }
    }
    __syncthreads();
    uint32_t* g = gbins + (size_t)scan * RS;
    for (int i = threadIdx.x; i < RS; i += kPolarThreads) {
        const uint32_t k = sbins[i];
        if (k > key_none) atomicMax(&g[i], k);        /* values <= -1000 can never be the bin result */
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&tickets[scan], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    /* ---- epilogue by the last CTA of this scan: decode, zero the empties, wire image, ring key, column statistics */
    float* sdesc = reinterpret_cast<float*>(sbins + RS);
    float* od = out_desc + (size_t)scan * RS;
    for (int i = threadIdx.x; i < RS; i += kPolarThreads) {
        const uint32_t k = __ldcg(&g[i]);
        float v = (k > key_none) ? scl_key_float(k) : 0.0f;  /* max(-1000, ..) == -1000 -> 0 (:1450-1453) */
        sdesc[i] = v;
        od[i] = v;
        g[i] = 0u;                                           /* leave the scratch clean for the next launch */
    }
    if (threadIdx.x == 0) tickets[scan] = 0;
    __syncthreads();
    /* K2: ring key = float(mean over the row in double, index order) (:1463-1475); the upper warps take the per-entry
     * cache K4 reads (sector key + column norms), so an insert needs no second launch over the new descriptors */
    if (threadIdx.x < 32) {
        for (int r = threadIdx.x; r < p.R; r += 32) {
            double s = 0.0;
            for (int c = 0; c < S; c++) s = __dadd_rn(s, (double)sdesc[r * S + c]);
            const float kf = __double2float_rn(__ddiv_rn(s, (double)S));
            out_keys[(size_t)scan * p.R + r] = kf;
            sdesc[RS + r] = kf;
        }
    } else if (out_cstat) {
        column_stats(sdesc, p.R, S, S, threadIdx.x - 32, kPolarThreads - 32, out_cstat + (size_t)scan * 2 * S);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float n2 = 0.0f;
        for (int r = 0; r < p.R; r++) n2 = fmaf(sdesc[RS + r], sdesc[RS + r], n2);
        out_knorm[scan] = n2;
        if (kn2max) atomicMax(reinterpret_cast<int*>(kn2max), __float_as_int(n2));   /* n2 >= 0: int order == float order */
    }
}


// K2 stand-alone: ring keys of n descriptors already in device memory (the insert path,
// descriptor.h:1572-1599). One warp per descriptor; the descriptor is staged in shared memory
// with a padded row pitch so the per-row sequential sums are bank-conflict free.
__global__ void __launch_bounds__(256) ring_key_kernel(const float* __restrict__ desc, int n, int R, int S,
                                                       float* __restrict__ keys, float* __restrict__ knorm, float* __restrict__ kn2max,
                                                       double* __restrict__ cstat)
{
    extern __shared__ float sm[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pitch = S | 1;
    float* my = sm + (size_t)warp * (R * pitch + R);
    float* mykey = my + R * pitch;
    for (int d = blockIdx.x * warps + warp; d < n; d += gridDim.x * warps) {
        const float* src = desc + (size_t)d * R * S;
        for (int i = lane; i < R * S; i += 32) my[(i / S) * pitch + (i % S)] = __ldg(src + i);
        __syncwarp();
        if (keys) {
            for (int r = lane; r < R; r += 32) {
                double s = 0.0;
                for (int c = 0; c < S; c++) s = __dadd_rn(s, (double)my[r * pitch + c]);
                const float kf = __double2float_rn(__ddiv_rn(s, (double)S));
                keys[(size_t)d * R + r] = kf;
                mykey[r] = kf;
            }
            __syncwarp();
            if (lane == 0) {
                float n2 = 0.0f;
                for (int r = 0; r < R; r++) n2 = fmaf(mykey[r], mykey[r], n2);
                knorm[d] = n2;
                if (kn2max) atomicMax(reinterpret_cast<int*>(kn2max), __float_as_int(n2));
            }
        }
        if (cstat) column_stats(my, R, S, pitch, lane, 32, cstat + (size_t)d * 2 * S);
        __syncwarp();
    }
}

// The same with compile-time geometry (20x60, 40x120): the descriptor is fetched with 16-byte loads that are ALL in
// flight before the first one is used (one DRAM round trip per descriptor instead of ten), the index arithmetic folds,
// and two warps share a descriptor's rows when R > 32. The sums keep the reference's order (sequential, FP64).
template <int R, int S, int kWarps>
__global__ void __launch_bounds__(32 * kWarps) ring_key_fixed_kernel(const float* __restrict__ desc, int n, float* __restrict__ keys,
                                                             float* __restrict__ knorm, float* __restrict__ kn2max, double* __restrict__ cstat)
{
    constexpr int kPitch = S | 1, kV4 = R * S / 4, kPerLane = (kV4 + 31) / 32;
    static_assert((R * S) % 4 == 0 && S % 4 == 0, "rows are whole float4s");
    __shared__ float tile[kWarps][R * kPitch];
    __shared__ float skey[kWarps][R];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.x * kWarps + warp;
    if (d >= n) return;
    const float4* src = reinterpret_cast<const float4*>(desc + (size_t)d * R * S);
    float4 v[kPerLane];
#pragma unroll
    for (int i = 0; i < kPerLane; i++) { const int k = lane + 32 * i; v[i] = k < kV4 ? __ldg(src + k) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
    for (int i = 0; i < kPerLane; i++) {
        const int k = lane + 32 * i;
        if (k < kV4) {
            const int e = 4 * k, r = e / S, c = e % S;       /* a float4 never straddles a row: S % 4 == 0 */
            float* t = &tile[warp][r * kPitch + c];
            t[0] = v[i].x; t[1] = v[i].y; t[2] = v[i].z; t[3] = v[i].w;
        }
    }
    __syncwarp();
    if (cstat) column_stats(tile[warp], R, S, kPitch, lane, 32, cstat + (size_t)d * 2 * S);
    if (!keys) return;
    for (int r = lane; r < R; r += 32) {
        double s = 0.0;
#pragma unroll 4
        for (int c = 0; c < S; c++) s = __dadd_rn(s, (double)tile[warp][r * kPitch + c]);
        const float kf = __double2float_rn(__ddiv_rn(s, (double)S));
        keys[(size_t)d * R + r] = kf;
        skey[warp][r] = kf;
    }
    __syncwarp();
    if (lane == 0) {
        float n2 = 0.0f;
#pragma unroll
        for (int r = 0; r < R; r++) n2 = fmaf(skey[warp][r], skey[warp][r], n2);
        knorm[d] = n2;
        if (kn2max) atomicMax(reinterpret_cast<int*>(kn2max), __float_as_int(n2));
    }
}

} // namespace

cudaError_t scl_launch_polar(const void* pts_dev, const int* offsets_dev, const int* offsets_host, int n_scans, int max_points, int stride_bytes,
                             int R, int S, double lidar_height, double max_radius, uint32_t* gbins, int* tickets,
                             float* out_desc, float* out_keys, float* out_knorm, float* kn2max, double* out_cstat, int* out_ring, int* out_sector,
                             cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    /* bin tables: built once per geometry on the host, kept in device memory */
    struct Cached { int R, S; double max_radius; int device; BinTables bt; float* dev; int n; };
    static std::vector<Cached> cache;
    static std::mutex cache_mu;
    PolarParams p;
    int p_device = 0;
    {
        std::lock_guard<std::mutex> lk(cache_mu);
        int device = 0; cudaGetDevice(&device);
        p_device = device;
        const Cached* hit = nullptr;
        for (const Cached& c : cache) if (c.R == R && c.S == S && c.max_radius == max_radius && c.device == device) { hit = &c; break; }
        if (!hit) {
            Cached c; c.R = R; c.S = S; c.max_radius = max_radius; c.device = device;
            std::vector<float> tab;
            build_tables(R, S, max_radius, c.bt, tab);
            c.n = (int)tab.size();
            cudaError_t e = cudaMalloc(&c.dev, (tab.size() + 1) * sizeof(float));
            if (e != cudaSuccess) return e;
            e = cudaMemcpy(c.dev, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return e;
            cache.push_back(c);
            hit = &cache.back();
        }
        p.R = R; p.S = S; p.lidar_height = lidar_height; p.bt = hit->bt; p.tab = hit->dev; p.n_tab = hit->n;
    }
    constexpr int kPPT = 4;        /* per batch; two batches per thread are live (software pipeline) */
    const int chunk = kPolarThreads * kPPT;
    int chunks = (max_points + chunk - 1) / chunk;
    if (chunks < 1) chunks = 1;
    const int vec4 = (stride_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(pts_dev) & 15) == 0);
    int max_sec = 0;
    for (int q = 0; q < 4; q++) if (p.bt.n_sec[q] > max_sec) max_sec = p.bt.n_sec[q];
    if (p.bt.n_ring > kRingPad - 1 || max_sec > kSecPad - 1) return cudaErrorNotSupported;       /* up to 64 rings x 124 sectors */
    const size_t smem = (size_t)R * S * 8 + (size_t)R * 4;
    if (smem > 200 * 1024) return cudaErrorNotSupported;
    /* The production path (16-byte aligned points, no per-point bin output) with the search depths the geometry needs:
     * 20x60 has 19 ring and 15 sector thresholds per quadrant (5 + 4 steps), 40x120 has 39 and 30 (6 + 5). Everything else
     * (packed 12-byte points, the parity tests' per-point bins) takes the generic instantiation. */
    const bool fast = vec4 && out_ring == nullptr;
    const int variant = !fast ? 0 : (p.bt.n_ring <= 31 && max_sec <= 15) ? 1 : 2;
    /* Offsets: a batch of up to kInlineScans scans carries them in the kernel parameters (no copy-engine operation in front
     * of the launch: a stream-ordered switch between the copy and the compute engine costs several microseconds). */
    ScanOffsets inl;
    const bool use_inl = offsets_host != nullptr && n_scans <= kInlineScans;
    if (use_inl) for (int i = 0; i <= n_scans; i++) inl.v[i] = offsets_host[i];
    else if (!offsets_dev) return cudaErrorInvalidValue;
    /* Grid: two full waves of resident CTAs (ncu on 64 scans x 113k points: 1216 CTAs at 5 per SM were 1.64 waves and the
     * SMs idled 19 % of the launch), but never more chunks than a scan has */
    auto waves_grid = [&](int per_sm) {
        const int want = (2 * per_sm * SCL_NUM_SMS) / n_scans;        /* measured on the same batch: 1 wave 61.4 us, 2 waves 60.6, 3 waves 62.9, 4 waves 64.0, 6 waves 65.5 */
        if (want >= 1 && chunks > want) chunks = want;
    };
    /* gridDim.y is limited to 65535: larger batches go in slices (the per-scan scratch rows move with the slice) */
    for (int s0 = 0; s0 < n_scans; s0 += 65535) {
        const int ns = n_scans - s0 < 65535 ? n_scans - s0 : 65535;
#define SCL_POLAR_LAUNCH(FAST, RT, ST)                                                                                              \
        do {                                                                                                                       \
            if (smem > 48 * 1024) {   /* geometries above ~6100 bins need the opt-in (per device: the attribute belongs to the current device's context) */ \
                static SclOncePerDevice once;                                                                                      \
                if (once.first()) {                                                                                                \
                    cudaError_t ea = cudaFuncSetAttribute(polar_bin_kernel<kPPT, FAST, RT, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
                    if (ea != cudaSuccess) return ea;                                                                              \
                }                                                                                                                  \
            }                                                                                                                      \
            if (s0 == 0) {                                                                                                         \
                /* resident CTAs per SM: asked once per (instantiation, device, shared-memory size), not per launch */               \
                static std::atomic<long long> occ_cache[16];                                                                       \
                const int dslot = p_device & 15;                                                                                   \
                const long long tag = ((long long)smem << 8) | 1;                                                                  \
                long long c = occ_cache[dslot].load(std::memory_order_relaxed);                                                    \
                int per_sm = (int)(c & 0xff) - 1;                                                                                  \
                if ((c >> 8) != (tag >> 8) || per_sm < 1) {                                                                        \
                    per_sm = 0;                                                                                                    \
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, polar_bin_kernel<kPPT, FAST, RT, ST>, kPolarThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 4; \
                    if (per_sm > 200) per_sm = 200;                                                                                \
                    occ_cache[dslot].store(((long long)smem << 8) | (long long)(per_sm + 1), std::memory_order_relaxed);          \
                }                                                                                                                  \
                waves_grid(per_sm);                                                                                                \
            }                                                                                                                      \
            dim3 grid(chunks, ns);                                                                                                 \
            polar_bin_kernel<kPPT, FAST, RT, ST><<<grid, kPolarThreads, smem, stream>>>(                                            \
                static_cast<const unsigned char*>(pts_dev), use_inl ? nullptr : offsets_dev + s0, inl, stride_bytes, vec4, p, gbins + (size_t)s0 * R * S, tickets + s0, \
                out_desc + (size_t)s0 * R * S, out_keys + (size_t)s0 * R, out_knorm + s0, kn2max,                                   \
                out_cstat ? out_cstat + (size_t)s0 * 2 * S : nullptr, out_ring, out_sector);                                        \
        } while (0)
        if (variant == 1) SCL_POLAR_LAUNCH(true, 16, 8);
        else if (variant == 2) SCL_POLAR_LAUNCH(true, 32, 16);
        else SCL_POLAR_LAUNCH(false, 32, 16);
#undef SCL_POLAR_LAUNCH
        cudaError_t el = cudaGetLastError();
        if (el != cudaSuccess) return el;
    }
    return cudaSuccess;
}

cudaError_t scl_launch_ring_keys(const float* desc_dev, int n, int R, int S, float* keys, float* knorm, float* kn2max, double* cstat,
                                 cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(desc_dev) & 15) == 0) {
        if (R == 20 && S == 60) { SCL_PREFER_SMEM((ring_key_fixed_kernel<20, 60, 4>)); ring_key_fixed_kernel<20, 60, 4><<<(n + 3) / 4, 128, 0, stream>>>(desc_dev, n, keys, knorm, kn2max, cstat); return cudaGetLastError(); }
        if (R == 40 && S == 120) { SCL_PREFER_SMEM((ring_key_fixed_kernel<40, 120, 2>)); ring_key_fixed_kernel<40, 120, 2><<<(n + 1) / 2, 64, 0, stream>>>(desc_dev, n, keys, knorm, kn2max, cstat); return cudaGetLastError(); }
    }
    int warps = 8;
    while (warps > 1 && (size_t)warps * (R * (S | 1) + R) * sizeof(float) > 200 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * (R * (S | 1) + R) * sizeof(float);
    /* per device, not per process: a function attribute belongs to the current device's context */
    cudaError_t ea = cudaFuncSetAttribute(ring_key_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (ea != cudaSuccess) return ea;
    int blocks = (n + warps - 1) / warps;
    if (blocks > 8 * SCL_NUM_SMS) blocks = 8 * SCL_NUM_SMS;
    ring_key_kernel<<<blocks, warps * 32, smem, stream>>>(desc_dev, n, R, S, keys, knorm, kn2max, cstat);
    return cudaGetLastError();
}

// The bin tables of a geometry, for inspection (include/scl_engine.h: scl_polar_tables). Pure host code: this is what the
// kernel keeps in shared memory, so a CPU test can check the tables against the oracle's per-point bins without a GPU.
int scl_polar_tables_host(int R, int S, double max_radius, float* ring_thr, int* n_ring, float* s_max,
                          float* sec_thr, int* n_sec, int* sec_base, int* sec_dir)
{
    if (R < 1 || S < 1 || !(max_radius > 0.0)) return 1;
    BinTables bt; std::vector<float> tab;
    build_tables(R, S, max_radius, bt, tab);
    int max_sec = 0;
    for (int q = 0; q < 4; q++) if (bt.n_sec[q] > max_sec) max_sec = bt.n_sec[q];
    if (bt.n_ring > kRingPad - 1 || max_sec > kSecPad - 1) return 2;
    if (n_ring) *n_ring = bt.n_ring;
    if (s_max) *s_max = bt.s_max;
    if (ring_thr) for (int i = 0; i < bt.n_ring; i++) ring_thr[i] = tab[i];
    for (int q = 0; q < 4; q++) {
        if (n_sec) n_sec[q] = bt.n_sec[q];
        if (sec_base) sec_base[q] = bt.sec_base[q];
        if (sec_dir) sec_dir[q] = bt.sec_dir[q];
        if (sec_thr) for (int k = 0; k < bt.n_sec[q]; k++) sec_thr[q * (kSecPad - 1) + k] = tab[bt.sec_off[q] + k];
    }
    return 0;
}

int scl_polar_inline_scans() { return kInlineScans; }

void scl_preload_k1()
{
    SCL_TOUCH((polar_bin_kernel<4, true, 16, 8>)); SCL_TOUCH((polar_bin_kernel<4, true, 32, 16>)); SCL_TOUCH((polar_bin_kernel<4, false, 32, 16>)); SCL_TOUCH(ring_key_kernel);
    SCL_TOUCH((ring_key_fixed_kernel<20, 60, 4>)); SCL_TOUCH((ring_key_fixed_kernel<40, 120, 2>));
}
