// epi_bench.cu — microbenchmark of the kNN epilogue inner loop in isolation: tcgen05.ld 32x32b.x32 (double buffered) +
// FMNMX3 tree + vote, W epilogue warps per SM, no MMA and no barriers. Cycles per 32-column chunk per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(taddr));
}
__device__ __forceinline__ void wait32(uint32_t (&r)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]),"+r"(r[1]),"+r"(r[2]),"+r"(r[3]),"+r"(r[4]),"+r"(r[5]),"+r"(r[6]),"+r"(r[7]),"+r"(r[8]),"+r"(r[9]),"+r"(r[10]),"+r"(r[11]),"+r"(r[12]),"+r"(r[13]),"+r"(r[14]),"+r"(r[15]),
          "+r"(r[16]),"+r"(r[17]),"+r"(r[18]),"+r"(r[19]),"+r"(r[20]),"+r"(r[21]),"+r"(r[22]),"+r"(r[23]),"+r"(r[24]),"+r"(r[25]),"+r"(r[26]),"+r"(r[27]),"+r"(r[28]),"+r"(r[29]),"+r"(r[30]),"+r"(r[31]) :: "memory");
}
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
template <int MODE> __device__ __forceinline__ bool examine(uint32_t (&r)[32], float thr)
{
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float x0 = __uint_as_float(r[8*j]), x1 = __uint_as_float(r[8*j+1]), x2 = __uint_as_float(r[8*j+2]), x3 = __uint_as_float(r[8*j+3]),
                    x4 = __uint_as_float(r[8*j+4]), x5 = __uint_as_float(r[8*j+5]), x6 = __uint_as_float(r[8*j+6]), x7 = __uint_as_float(r[8*j+7]);
        if (MODE == 0) g[j] = fminf(fmin3(fmin3(x0, x1, x2), fmin3(x3, x4, x5), x6), x7);
        else g[j] = fminf(fminf(fminf(x0, x1), fminf(x2, x3)), fminf(fminf(x4, x5), fminf(x6, x7)));
    }
    const float m = fminf(fminf(g[0], g[1]), fminf(g[2], g[3]));
    return __any_sync(0xffffffffu, m < thr);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
template <int MODE>
__global__ void __launch_bounds__(416, 1) bench(int iters, int warps, float thr, int mma_n, long long* out, int* sink)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    __shared__ volatile int stop;
    if (mma_n > 0) for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 416) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { stop = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
        uint32_t va[32], vb[32];
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128 % 512);
        const long long t0 = clock64();
        ld32(base, va);
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
            wait32(va); ld32(base + 32, vb); hits += examine<MODE>(va, thr);
            wait32(vb); ld32(base + 64, va); hits += examine<MODE>(vb, thr);
            wait32(va); ld32(base + 96, vb); hits += examine<MODE>(va, thr);
            wait32(vb); ld32(base, va); hits += examine<MODE>(vb, thr);
        }
        wait32(va);
        const long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
        __syncwarp();
        if (threadIdx.x == 0) stop = 1;
    } else if (mma_n > 0 && threadIdx.x == 384) {
        /* a stream of bf16 MMAs (M = 128, N = mma_n, K = 16) into the upper TMEM columns, 4 per accumulator like the kNN kernel */
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(mma_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 16 * 1024;
        int n = 0;
        while (!stop && n < 400000) {
            for (int k = 0; k < 4; k++)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + 256 + (uint32_t)((n & 4) ? 128 : 0)), "l"(make_desc(a0 + k * 2 * 2048, 2048, 128)), "l"(make_desc(b0 + k * 2 * 2048, 2048, 128)), "r"(idesc), "r"(k > 0 ? 1u : 0u) : "memory");
            n += 4;
            if ((n & 31) == 0) {   /* bound the queue depth */
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t done = 0; const uint32_t par = ((n >> 5) - 1) & 1;
                while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(par) : "memory");
            }
        }
        if (blockIdx.x == 0) out[1] = n;
    }
    if (threadIdx.x < 384) sink[blockIdx.x * 384 + threadIdx.x] = hits;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
template <int MODE> void run(const char* name, int warps, int mma_n, int dyn_smem = 48 * 1024)
{
    long long* d; int* sink; cudaMalloc(&d, 16); cudaMalloc(&sink, 148 * 384 * 4); cudaMemset(d, 0, 16);
    const int iters = 2000;
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rep = 0; rep < 2; rep++) bench<MODE><<<148, 416, dyn_smem>>>(iters, warps, -1.0e30f, mma_n, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double per_chunk = (double)h[0] / iters / 4;
    printf("%-8s smem %3d KB %2d warps/SM, concurrent MMA N=%3d: %6.1f cycles per 32-column chunk per warp = %6.1f cycles per 256x128 scores per SM; %6.1f cycles per MMA (%s)\n", name, dyn_smem / 1024, warps, mma_n,
           per_chunk, per_chunk * 32.0 / warps, h[1] ? (double)h[0] / h[1] : 0.0, cudaGetErrorString(e));
    cudaFree(d); cudaFree(sink);
}
int main()
{
    for (int sm : {0, 48 * 1024, 190 * 1024}) for (int w : {4, 8}) run<0>("FMNMX3", w, 0, sm);
    for (int w : {4, 8}) run<0>("FMNMX3", w, 128, 48 * 1024);
    return 0;
}
