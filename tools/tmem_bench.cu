// tmem_bench.cu — microbenchmark: cycles per tcgen05.ld (32x32b) per warp, alone and under a concurrent tcgen05.mma stream.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
template <int W> __device__ __forceinline__ uint32_t ld(uint32_t taddr);
template <> __device__ __forceinline__ uint32_t ld<32>(uint32_t taddr)
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) x ^= r[i];
    return x;
}
template <> __device__ __forceinline__ uint32_t ld<8>(uint32_t taddr)
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return r[0]^r[1]^r[2]^r[3]^r[4]^r[5]^r[6]^r[7];
}
template <int W>
__global__ void __launch_bounds__(160, 1) bench(int iters, int with_mma, long long* out, uint32_t* sink)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 160) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { stop = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp < 4) {
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) acc ^= ld<W>(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)((it * W) & 255));
        const long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
        sink[blockIdx.x * 128 + threadIdx.x] = acc;
        __syncwarp();
        if (threadIdx.x == 0) stop = 1;
    } else if (with_mma && threadIdx.x == 128) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 16 * 1024;
        int n = 0;
        while (!stop && n < 200000) {
            for (int k = 0; k < 3; k++)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + 256), "l"(make_desc(a0 + k * 2 * 2048, 2048, 128)), "l"(make_desc(b0 + k * 2 * 4096, 4096, 128)), "r"(idesc), "r"(1u) : "memory");
            n += 3;
            if ((n & 63) == 0) {   /* bound the queue depth */
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t done = 0; const uint32_t par = ((n >> 6) - 1) & 1;
                while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(par) : "memory");
            }
        }
        if (blockIdx.x == 0) out[1] = n;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
template <int W> void run(const char* name, int with_mma)
{
    long long* d; uint32_t* sink; cudaMalloc(&d, 16); cudaMalloc(&sink, 148 * 128 * 4); cudaMemset(d, 0, 16);
    const int iters = 4096;
    cudaFuncSetAttribute(bench<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 2; rep++) bench<W><<<148, 160, 64 * 1024>>>(iters, with_mma, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-28s mma=%d : %7.1f cycles per ld per warp = %6.1f B/clk/SM (4 warps)   mmas issued %lld (%s)\n", name, with_mma,
           (double)h[0] / iters, 4.0 * W * 32 * 4 / ((double)h[0] / iters), h[1], cudaGetErrorString(e));
    cudaFree(d); cudaFree(sink);
}
int main()
{
    run<32>("ld 32x32b.x32 + wait", 0);
    run<8>("ld 32x32b.x8 + wait", 0);
    run<32>("ld 32x32b.x32 + wait", 1);
    run<8>("ld 32x32b.x8 + wait", 1);
    return 0;
}
