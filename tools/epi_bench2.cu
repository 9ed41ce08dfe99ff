// epi_bench2.cu — which software pipeline of tcgen05.ld + min-tree reaches the TMEM read bandwidth? 8 warps per SM (two per scheduler),
// no MMA, no barriers; cycles per 256 x 128 scores per SM (= one key tile of the kNN kernel).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define LD32(r, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]), \
          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(addr))
#define WAIT32(r) asm volatile("tcgen05.wait::ld.sync.aligned;" \
        : "+r"(r[0]),"+r"(r[1]),"+r"(r[2]),"+r"(r[3]),"+r"(r[4]),"+r"(r[5]),"+r"(r[6]),"+r"(r[7]),"+r"(r[8]),"+r"(r[9]),"+r"(r[10]),"+r"(r[11]),"+r"(r[12]),"+r"(r[13]),"+r"(r[14]),"+r"(r[15]), \
          "+r"(r[16]),"+r"(r[17]),"+r"(r[18]),"+r"(r[19]),"+r"(r[20]),"+r"(r[21]),"+r"(r[22]),"+r"(r[23]),"+r"(r[24]),"+r"(r[25]),"+r"(r[26]),"+r"(r[27]),"+r"(r[28]),"+r"(r[29]),"+r"(r[30]),"+r"(r[31]) :: "memory")
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ int examine(uint32_t (&r)[32], float thr)
{
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float x0 = __uint_as_float(r[8*j]), x1 = __uint_as_float(r[8*j+1]), x2 = __uint_as_float(r[8*j+2]), x3 = __uint_as_float(r[8*j+3]),
                    x4 = __uint_as_float(r[8*j+4]), x5 = __uint_as_float(r[8*j+5]), x6 = __uint_as_float(r[8*j+6]), x7 = __uint_as_float(r[8*j+7]);
        g[j] = fminf(fmin3(fmin3(x0, x1, x2), fmin3(x3, x4, x5), x6), x7);
    }
    const float m = fminf(fmin3(g[0], g[1], g[2]), g[3]);
    return __any_sync(0xffffffffu, m < thr) ? 1 : 0;
}
template <int V>
__global__ void __launch_bounds__(384, 1) bench(int iters, int warps, float thr, long long* out, int* sink)
{
    __shared__ uint32_t tmem_slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const int warp = threadIdx.x >> 5;
    int hits = 0;
    if (warp < warps) {
        uint32_t va[32], vb[32], vc[32], vd[32];
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128 % 512);
        const long long t0 = clock64();
        if (V == 0) {            /* the kNN kernel today: one load in flight while the previous chunk is examined */
            LD32(va, base);
#pragma unroll 1
            for (int it = 0; it < iters; it++) {
                WAIT32(va); LD32(vb, base + 32); hits += examine(va, thr);
                WAIT32(vb); LD32(va, base + 64); hits += examine(vb, thr);
                WAIT32(va); LD32(vb, base + 96); hits += examine(va, thr);
                WAIT32(vb); LD32(va, base); hits += examine(vb, thr);
            }
            WAIT32(va);
        } else if (V == 1) {     /* pairs: two loads back to back, then both examined */
#pragma unroll 1
            for (int it = 0; it < iters; it++) {
                LD32(va, base); LD32(vb, base + 32); WAIT32(va); WAIT32(vb); hits += examine(va, thr); hits += examine(vb, thr);
                LD32(va, base + 64); LD32(vb, base + 96); WAIT32(va); WAIT32(vb); hits += examine(va, thr); hits += examine(vb, thr);
            }
        } else if (V == 2) {     /* quads: the whole 128-column slot requested at once */
#pragma unroll 1
            for (int it = 0; it < iters; it++) {
                LD32(va, base); LD32(vb, base + 32); LD32(vc, base + 64); LD32(vd, base + 96);
                WAIT32(va); hits += examine(va, thr); WAIT32(vb); hits += examine(vb, thr);
                WAIT32(vc); hits += examine(vc, thr); WAIT32(vd); hits += examine(vd, thr);
            }
        } else if (V == 3) {     /* pairs, software pipelined: the next pair is requested before the current pair is examined */
            LD32(va, base); LD32(vb, base + 32);
#pragma unroll 1
            for (int it = 0; it < iters; it++) {
                WAIT32(va); WAIT32(vb); LD32(vc, base + 64); LD32(vd, base + 96); hits += examine(va, thr); hits += examine(vb, thr);
                WAIT32(vc); WAIT32(vd); LD32(va, base); LD32(vb, base + 32); hits += examine(vc, thr); hits += examine(vd, thr);
            }
            WAIT32(va); WAIT32(vb);
        }
        const long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    sink[blockIdx.x * 384 + threadIdx.x] = hits;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
template <int V> void run(const char* name, int warps)
{
    long long* d; int* sink; cudaMalloc(&d, 16); cudaMalloc(&sink, 148 * 384 * 4); cudaMemset(d, 0, 16);
    const int iters = 2000;
    for (int rep = 0; rep < 2; rep++) bench<V><<<148, 384>>>(iters, warps, -1.0e30f, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double per_chunk = (double)h / iters / 4;
    printf("%-42s %2d warps/SM: %6.1f cycles per 32-column chunk per warp = %6.1f cycles per 256x128 scores per SM (%s)\n", name, warps,
           per_chunk, per_chunk * 32.0 / warps, cudaGetErrorString(e));
    cudaFree(d); cudaFree(sink);
}
int main()
{
    for (int w : {4, 8}) {
        run<0>("one load ahead (double buffer)", w);
        run<1>("pairs, not pipelined", w);
        run<2>("quads, not pipelined", w);
        run<3>("pairs, pipelined (4 buffers)", w);
    }
    return 0;
}
