// k1_polar.cu — K1 polar binning (+ K2 ring keys in its epilogue) and the stand-alone K2 kernel.
//
// Replaces makeScancontext (/root/reference/include/descriptor.h:1404-1461), xy2theta
// (:1352-1374), makeRingkeyFromScancontext (:1463-1475) and the database append of save()
// (:1587-1599).
//
// Layout: clouds arrive as the caller's array-of-structs (pcl::PointXYZI, 32 B/point, x y z at
// byte 0 4 8); with 16-byte aligned points one LDG.128 fetches x,y,z,pad per point. A batch of
// scans is one launch: grid = (chunks per scan, scans). Each CTA bins its chunk into a
// shared-memory R x S array of order-preserving uint keys with a warp-aggregated atomicMax
// (lanes that hit the same bin are first reduced with __match_any_sync/__reduce_max_sync, so
// one ATOMS per distinct bin per warp), merges that array into the scan's global bin array with
// atomicMax, and the LAST CTA of the scan (ticket counter) decodes the bins, applies the
// "-1000 -> 0" rule, writes the R*S float wire image straight into the database slot and
// reduces the ring key. Because max is order-free and heights are plain floats, bin contents
// are bit-exact whatever the arrival order.
//
// Roofline: HBM streaming, 32 B/point read in place (16 B/point algorithmic for packed input),
// R*S*4 + R*4 B written per scan; arithmetic per point is ~60 FP32/FP64 ops (the bit-exact
// atanf and the double-precision index math).
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr float kNoPoint = -1000.0f;   /* descriptor.h:1411 */

struct PolarParams {
    int R, S;
    double lidar_height, max_radius;
};

// Per-point bin computation, operation for operation what the oracle does (sc_oracle.cpp
// makeScancontext), which in turn follows descriptor.h:1420-1435. Returns false if dropped.
__device__ __forceinline__ bool polar_bin(const PolarParams& p, float x, float y, float z, int& ring, int& sector, float& zf)
{
    zf = __double2float_rn(__dadd_rn((double)z, p.lidar_height));                  /* :1422 */
    const float azim_range = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))); /* :1425 */
    const double k = 180.0 / 3.14159265358979323846;                                 /* (180/M_PI) */
    double theta;
    if ((x >= 0) & (y >= 0))      theta = __dmul_rn(k, (double)scl_atanf(__fdiv_rn(y, x)));
    else if ((x < 0) & (y >= 0))  theta = __dsub_rn(180.0, __dmul_rn(k, (double)scl_atanf(__fdiv_rn(y, -x))));
    else if ((x < 0) & (y < 0))   theta = __dadd_rn(180.0, __dmul_rn(k, (double)scl_atanf(__fdiv_rn(y, x))));
    else if ((x >= 0) & (y < 0))  theta = __dsub_rn(360.0, __dmul_rn(k, (double)scl_atanf(__fdiv_rn(-y, x))));
    else return false;                          /* NaN coordinate: UB in the reference, dropped (Q9) */
    const float azim_angle = __double2float_rn(theta);                               /* returned as float */
    if ((double)azim_range > p.max_radius) return false;                             /* :1429 */
    ring = max(min(p.R, scl_ceil_to_int(__dmul_rn(__ddiv_rn((double)azim_range, p.max_radius), (double)p.R))), 1);
    sector = max(min(p.S, scl_ceil_to_int(__dmul_rn(__ddiv_rn((double)azim_angle, 360.0), (double)p.S))), 1);
    return true;
}

__device__ __forceinline__ void load_xyz(const unsigned char* base, size_t i, int stride, bool vec4, float& x, float& y, float& z)
{
    const unsigned char* q = base + i * (size_t)stride;
    if (vec4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(q));
        x = v.x; y = v.y; z = v.z;
    } else {
        const float* f = reinterpret_cast<const float*>(q);
        x = __ldg(f); y = __ldg(f + 1); z = __ldg(f + 2);
    }
}

// grid = (chunks, scans); block = 256
template <int kPointsPerThread>
__global__ void __launch_bounds__(256) polar_bin_kernel(
    const unsigned char* __restrict__ pts, const int* __restrict__ offsets, int stride, int vec4,
    PolarParams p, uint32_t* __restrict__ gbins /* [scans][R*S], zero between launches */,
    int* __restrict__ tickets /* [scans], zero between launches */,
    float* __restrict__ out_desc /* [scans][R*S] */, float* __restrict__ out_keys /* [scans][R] */,
    float* __restrict__ out_knorm /* [scans] */, float* __restrict__ kn2max, int* __restrict__ out_ring, int* __restrict__ out_sector)
{
    extern __shared__ uint32_t sbins[];   /* R*S keys, then R*S floats for the epilogue */
    __shared__ int s_last;
    const int RS = p.R * p.S;
    const int scan = blockIdx.y;
    const int p0 = offsets[scan], p1 = offsets[scan + 1];
    const uint32_t key_none = scl_float_key(kNoPoint);

    for (int i = threadIdx.x; i < RS; i += blockDim.x) sbins[i] = 0u;
    __syncthreads();

    const int chunk = blockDim.x * kPointsPerThread;
    for (int base = p0 + blockIdx.x * chunk; base < p1; base += gridDim.x * chunk) {
        float x[kPointsPerThread], y[kPointsPerThread], z[kPointsPerThread];
#pragma unroll
        for (int j = 0; j < kPointsPerThread; j++) {   /* all loads first: kPointsPerThread LDG.128 in flight */
            const int i = base + j * blockDim.x + threadIdx.x;
            if (i < p1) load_xyz(pts, (size_t)i, stride, vec4 != 0, x[j], y[j], z[j]);
            else { x[j] = y[j] = z[j] = __int_as_float(0x7fc00000); }
        }
#pragma unroll
        for (int j = 0; j < kPointsPerThread; j++) {
            const int i = base + j * blockDim.x + threadIdx.x;
            int ring = 0, sector = 0; float zf;
            bool ok = (i < p1) && polar_bin(p, x[j], y[j], z[j], ring, sector, zf);
            if (out_ring != nullptr && i < p1) { out_ring[i] = ok ? ring : 0; out_sector[i] = ok ? sector : 0; }
            ok = ok && !(zf != zf);                     /* desc < NaN is false: NaN heights never win (:1438) */
            const int bin = ok ? (ring - 1) * p.S + (sector - 1) : -1;
            const uint32_t key = ok ? scl_float_key(zf) : 0u;
            /* warp-aggregated max: one shared atomic per distinct bin per warp */
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            const uint32_t m = __reduce_max_sync(peers, key);
            if (ok && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicMax(&sbins[bin], m);
        }
    }
    __syncthreads();
    uint32_t* g = gbins + (size_t)scan * RS;
    for (int i = threadIdx.x; i < RS; i += blockDim.x) {
        const uint32_t k = sbins[i];
        if (k > key_none) atomicMax(&g[i], k);        /* values <= -1000 can never be the bin result */
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&tickets[scan], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    /* ---- epilogue by the last CTA of this scan: decode, zero the empties, wire image, ring key */
    float* sdesc = reinterpret_cast<float*>(sbins + RS);
    float* od = out_desc + (size_t)scan * RS;
    for (int i = threadIdx.x; i < RS; i += blockDim.x) {
        const uint32_t k = __ldcg(&g[i]);
        float v = (k > key_none) ? scl_key_float(k) : 0.0f;  /* max(-1000, ..) == -1000 -> 0 (:1450-1453) */
        sdesc[i] = v;
        od[i] = v;
        g[i] = 0u;                                           /* leave the scratch clean for the next launch */
    }
    if (threadIdx.x == 0) tickets[scan] = 0;
    __syncthreads();
    /* K2: ring key = float(mean over the row in double, index order) (:1463-1475) */
    for (int r = threadIdx.x; r < p.R; r += blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < p.S; c++) s = __dadd_rn(s, (double)sdesc[r * p.S + c]);
        const float kf = __double2float_rn(__ddiv_rn(s, (double)p.S));
        out_keys[(size_t)scan * p.R + r] = kf;
        sdesc[RS + r] = kf;
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
                                                       float* __restrict__ keys, float* __restrict__ knorm, float* __restrict__ kn2max)
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
        __syncwarp();
    }
}

} // namespace

cudaError_t scl_launch_polar(const void* pts_dev, const int* offsets_dev, int n_scans, int max_points, int stride_bytes,
                             int R, int S, double lidar_height, double max_radius, uint32_t* gbins, int* tickets,
                             float* out_desc, float* out_keys, float* out_knorm, float* kn2max, int* out_ring, int* out_sector,
                             cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    PolarParams p{R, S, lidar_height, max_radius};
    constexpr int kPPT = 8;
    const int chunk = 256 * kPPT;
    int chunks = (max_points + chunk - 1) / chunk;
    if (chunks < 1) chunks = 1;
    /* enough CTAs to cover the machine about 4x, but never more chunks than a scan has */
    const int want = (4 * SCL_NUM_SMS + n_scans - 1) / n_scans;
    if (chunks > want) chunks = want;
    const int vec4 = (stride_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(pts_dev) & 15) == 0);
    const size_t smem = (size_t)R * S * 8 + (size_t)R * 4;
    dim3 grid(chunks, n_scans);
    polar_bin_kernel<kPPT><<<grid, 256, smem, stream>>>(static_cast<const unsigned char*>(pts_dev), offsets_dev, stride_bytes, vec4, p,
                                                       gbins, tickets, out_desc, out_keys, out_knorm, kn2max, out_ring, out_sector);
    return cudaGetLastError();
}

cudaError_t scl_launch_ring_keys(const float* desc_dev, int n, int R, int S, float* keys, float* knorm, float* kn2max, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const int warps = 8;
    const size_t smem = (size_t)warps * (R * (S | 1) + R) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) { cudaFuncSetAttribute(ring_key_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_done = true; }
    int blocks = (n + warps - 1) / warps;
    if (blocks > 8 * SCL_NUM_SMS) blocks = 8 * SCL_NUM_SMS;
    ring_key_kernel<<<blocks, warps * 32, smem, stream>>>(desc_dev, n, R, S, keys, knorm, kn2max);
    return cudaGetLastError();
}
