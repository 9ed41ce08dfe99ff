// k6_cloud.cu — K6, the cloud preparation either side of the descriptor and ICP kernels:
//
//   * scl_voxel_grid       replaces pcl::VoxelGrid<PointXYZI>::filter as used by downSizeFilterDes in front of the
//                          descriptor (/root/reference/include/distributedMapping.h:996-998) and by downSizeFilterICP
//                          (:1181-1185, :1200-1201): one centroid (x, y, z, intensity) per occupied leaf, output ordered
//                          by the linear leaf index.
//   * scl_assemble_submap  replaces loopFindNearKeyframes (:1163-1186): every keyframe cloud moved into the world frame
//                          with transformPointCloud (:234-253), concatenated, then the VoxelGrid above.
//
// PCL is not vendored in /root/reference, so the arithmetic follows PCL's published algorithm (voxel_grid.hpp
// applyFilter, centroid.h CentroidPoint): float min/max of the finite points, inverse leaf = 1.0f / leaf,
// min_b = int(floor(min * inv)), leaf index ijk = int(floor(p * inv) - float(min_b)), linear index
// ijk.x + ijk.y * div.x + ijk.z * div.x * div.y, points sorted by that index, float sums divided by float(n).
// PCL sorts with std::sort (unstable), so the summation order inside a leaf is not defined upstream; here (and in
// the oracle) it is the input order: the sort is stable and one thread adds a leaf's points in sequence, which makes
// the result bit-reproducible ("parity unpinned" against PCL itself, bit-exact against oracle/cloud_oracle.cpp).
//
// The 6-DoF pose -> 3x4 matrix step (pcl::getTransformation: libm sinf/cosf) is done on the host like the reference
// does, so the kernel only carries the per-point expression of :246-248 with explicit round-to-nearest mul/add
// (the reference builds for baseline x86-64: no FMA).
//
// Roofline: HBM streaming — 32 B/point read (PCL layout in place), 16 B/point written, plus the sort's passes
// (cub::DeviceRadixSort, library code: 8 B/point per pass). Sorting is the only library call of the engine.
#include "common.cuh"
#include "kernels.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace {

__device__ __forceinline__ bool finite3(float x, float y, float z)
{
    return (fabsf(x) <= 3.402823466e+38f) & (fabsf(y) <= 3.402823466e+38f) & (fabsf(z) <= 3.402823466e+38f);   /* false for NaN and inf */
}
// x, y, z from the first 12 bytes; intensity at byte 12 of a packed 16-byte record, at byte 16 of a pcl::PointXYZI (stride >= 32)
__device__ __forceinline__ float4 load_xyzi(const unsigned char* p, int stride)
{
    float4 v = *reinterpret_cast<const float4*>(p);
    if (stride >= 32) v.w = *reinterpret_cast<const float*>(p + 16);
    return v;
}
__device__ __forceinline__ int ordered_i(float f) { const int b = __float_as_int(f); return b ^ ((b >> 31) & 0x7fffffff); }

// out[offsets[c] + i] = T_c * p (x, y, z), intensity kept (distributedMapping.h:243-249); 16 B per output point
__global__ void __launch_bounds__(256) transform_concat_kernel(const unsigned char* __restrict__ pts, const int* __restrict__ offsets, int n_clouds,
                                                                int stride, const float* __restrict__ T /* [n_clouds][12] row-major 3x4 */,
                                                                float4* __restrict__ out)
{
    const int c = blockIdx.y;
    const int p0 = offsets[c], p1 = offsets[c + 1];
    __shared__ float t[12];
    if (threadIdx.x < 12) t[threadIdx.x] = T[c * 12 + threadIdx.x];
    __syncthreads();
    for (int i = p0 + blockIdx.x * blockDim.x + threadIdx.x; i < p1; i += gridDim.x * blockDim.x) {
        const float4 v = load_xyzi(pts + (size_t)i * stride, stride);
        float4 o;
        o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[0], v.x), __fmul_rn(t[1], v.y)), __fmul_rn(t[2], v.z)), t[3]);
        o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[4], v.x), __fmul_rn(t[5], v.y)), __fmul_rn(t[6], v.z)), t[7]);
        o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[8], v.x), __fmul_rn(t[9], v.y)), __fmul_rn(t[10], v.z)), t[11]);
        o.w = v.w;
        out[i] = o;
    }
}

// min / max of the finite points: warp reduction, then atomics on the order-preserving int image. bounds[0..2] = min, [3..5] = max
__global__ void __launch_bounds__(256) bounds_kernel(const unsigned char* __restrict__ pts, int n, int stride, int* __restrict__ bounds)
{
    int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 v = *reinterpret_cast<const float4*>(pts + (size_t)i * stride);
        if (finite3(v.x, v.y, v.z)) {
            /* -0.0 and +0.0 compare equal as floats; their int images differ by one: harmless for floor(min * inv) */
            const int a[3] = {ordered_i(v.x), ordered_i(v.y), ordered_i(v.z)};
#pragma unroll
            for (int k = 0; k < 3; k++) { mn[k] = min(mn[k], a[k]); mx[k] = max(mx[k], a[k]); }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        mn[k] = __reduce_min_sync(0xffffffffu, mn[k]); mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(bounds + k, mn[k]); atomicMax(bounds + 3 + k, mx[k]); }
    }
}

struct VoxelGeom { float inv; int min_b[3]; int div0, div01; };

// leaf index per point (voxel_grid.hpp: ijk = int(floor(p * inv) - float(min_b))); non-finite points sort last
__global__ void __launch_bounds__(256) voxel_key_kernel(const unsigned char* __restrict__ pts, int n, int stride, VoxelGeom g,
                                                         uint32_t* __restrict__ keys, int* __restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = *reinterpret_cast<const float4*>(pts + (size_t)i * stride);
    uint32_t key = 0xffffffffu;
    if (finite3(v.x, v.y, v.z)) {
        const int i0 = (int)__fsub_rn(floorf(__fmul_rn(v.x, g.inv)), (float)g.min_b[0]);
        const int i1 = (int)__fsub_rn(floorf(__fmul_rn(v.y, g.inv)), (float)g.min_b[1]);
        const int i2 = (int)__fsub_rn(floorf(__fmul_rn(v.z, g.inv)), (float)g.min_b[2]);
        key = (uint32_t)(i0 + i1 * g.div0 + i2 * g.div01);
    }
    keys[i] = key;
    vals[i] = i;
}

__global__ void __launch_bounds__(256) voxel_head_kernel(const uint32_t* __restrict__ keys, int n, int* __restrict__ head)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    head[i] = (k != 0xffffffffu && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
}

// one thread per leaf: its points are added in sorted (= input) order, then divided by float(n) (centroid.h)
__global__ void __launch_bounds__(128) voxel_centroid_kernel(const unsigned char* __restrict__ pts, int stride, const uint32_t* __restrict__ keys,
                                                              const int* __restrict__ vals, const int* __restrict__ head, const int* __restrict__ ord,
                                                              int n, float4* __restrict__ out, int* __restrict__ n_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == n - 1) *n_out = ord[i] + head[i];
    if (!head[i]) return;
    const uint32_t k = keys[i];
    float sx = 0.0f, sy = 0.0f, sz = 0.0f, si = 0.0f;
    int cnt = 0;
    for (int j = i; j < n && keys[j] == k; j++) {
        const float4 v = load_xyzi(pts + (size_t)vals[j] * stride, stride);
        sx = __fadd_rn(sx, v.x); sy = __fadd_rn(sy, v.y); sz = __fadd_rn(sz, v.z); si = __fadd_rn(si, v.w);
        cnt++;
    }
    const float c = (float)cnt;
    out[ord[i]] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
}

__global__ void __launch_bounds__(256) pack_xyzi_kernel(const unsigned char* __restrict__ pts, int n, int stride, float4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = load_xyzi(pts + (size_t)i * stride, stride);
}

} // namespace

cudaError_t scl_launch_transform_concat(const void* pts, const int* offsets_dev, int n_clouds, int max_points, int stride_bytes,
                                        const float* T_dev, void* out_xyzi, cudaStream_t stream)
{
    if (n_clouds <= 0 || max_points <= 0) return cudaSuccess;
    int bx = (max_points + 255) / 256;
    if (bx > 4 * SCL_NUM_SMS) bx = 4 * SCL_NUM_SMS;
    transform_concat_kernel<<<dim3(bx, n_clouds), 256, 0, stream>>>(static_cast<const unsigned char*>(pts), offsets_dev, n_clouds, stride_bytes,
                                                                  T_dev, static_cast<float4*>(out_xyzi));
    return cudaGetLastError();
}

cudaError_t scl_launch_cloud_bounds(const void* pts, int n, int stride_bytes, int* bounds6, cudaStream_t stream)
{
    static const int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    cudaError_t e = cudaMemcpyAsync(bounds6, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess || n <= 0) return e;
    int blocks = (n + 255) / 256;
    if (blocks > 4 * SCL_NUM_SMS) blocks = 4 * SCL_NUM_SMS;
    bounds_kernel<<<blocks, 256, 0, stream>>>(static_cast<const unsigned char*>(pts), n, stride_bytes, bounds6);
    return cudaGetLastError();
}

size_t scl_voxel_temp_bytes(int n)
{
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int*)nullptr, (int*)nullptr, n);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, n);
    return (a > b ? a : b) + 256;
}

// keys_a/keys_b, vals_a/vals_b: n entries each; head/ord: n ints; out: n float4 capacity; n_out: device int
cudaError_t scl_launch_voxel_grid(const void* pts, int n, int stride_bytes, float inv_leaf, const int* min_b, int div0, int div01, int key_bits,
                                  uint32_t* keys_a, uint32_t* keys_b, int* vals_a, int* vals_b, int* head, int* ord,
                                  void* temp, size_t temp_bytes, void* out_xyzi, int* n_out, cudaStream_t stream)
{
    if (n <= 0) return cudaMemsetAsync(n_out, 0, sizeof(int), stream);
    VoxelGeom g; g.inv = inv_leaf; g.min_b[0] = min_b[0]; g.min_b[1] = min_b[1]; g.min_b[2] = min_b[2]; g.div0 = div0; g.div01 = div01;
    const int blocks = (n + 255) / 256;
    voxel_key_kernel<<<blocks, 256, 0, stream>>>(static_cast<const unsigned char*>(pts), n, stride_bytes, g, keys_a, vals_a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    size_t tb = temp_bytes;
    (void)key_bits;                                       /* all 32 bits: the sentinel of non-finite points is 0xffffffff */
    e = cub::DeviceRadixSort::SortPairs(temp, tb, keys_a, keys_b, vals_a, vals_b, n, 0, 32, stream);
    if (e != cudaSuccess) return e;
    voxel_head_kernel<<<blocks, 256, 0, stream>>>(keys_b, n, head);
    tb = temp_bytes;
    e = cub::DeviceScan::ExclusiveSum(temp, tb, head, ord, n, stream);
    if (e != cudaSuccess) return e;
    voxel_centroid_kernel<<<(n + 127) / 128, 128, 0, stream>>>(static_cast<const unsigned char*>(pts), stride_bytes, keys_b, vals_b, head, ord, n,
                                                             static_cast<float4*>(out_xyzi), n_out);
    return cudaGetLastError();
}

cudaError_t scl_launch_pack_xyzi(const void* pts, int n, int stride_bytes, void* out_xyzi, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    pack_xyzi_kernel<<<(n + 255) / 256, 256, 0, stream>>>(static_cast<const unsigned char*>(pts), n, stride_bytes, static_cast<float4*>(out_xyzi));
    return cudaGetLastError();
}
