// mma_bench.cu — microbenchmark: cycles per tcgen05.mma (cta_group::1, M=128) for operand layouts / kinds.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>  // 0 tf32, 1 f16(bf16)
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// variant: KIND, swizzle (0 none, 1 = 128B), N
template <int KIND, int SW, int N, int TS = 0>
__global__ void __launch_bounds__(128, 1) bench(int iters, int ksteps, long long* out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t fmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        for (int it = 0; it < iters; it++) {
            for (int k = 0; k < ksteps; k++) {
                uint64_t da, db;
                if (SW == 0) {   // K-major no swizzle: chunk stride (LBO) = rows*16, SBO = 128; K step = 2 chunks
                    da = make_desc(a0 + k * 2 * 128 * 16, 128 * 16, 128, 0);
                    db = make_desc(b0 + k * 2 * N * 16, N * 16, 128, 0);
                } else {         // K-major 128B swizzle: row = 128 B, 8-row atom = 1024 B; K step = +32 B
                    da = make_desc(a0 + k * 32, 16, 1024, 2);
                    db = make_desc(b0 + k * 32, 16, 1024, 2);
                }
                if (TS) mma_ts((N <= 128 ? tmem + (it % 3) * 128 : tmem + (it & 1) * 256 * 0), tmem + 448 + k * 8, db, idesc, (k > 0) ? 1u : 0u);
                else mma<KIND>((N <= 128 ? tmem + (it & 3) * 128 : tmem + (it & 1) * 256), da, db, idesc, (k > 0) ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int KIND, int SW, int N, int TS = 0>
void run(const char* name, int ksteps)
{
    long long* d; cudaMalloc(&d, 8);
    const int iters = 512;
    cudaFuncSetAttribute(bench<KIND, SW, N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; rep++) bench<KIND, SW, N, TS><<<148, 128, 96 * 1024>>>(iters, ksteps, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-34s ksteps=%d : %8.1f cycles / mma   (%s)\n", name, ksteps, (double)h / (iters * ksteps), cudaGetErrorString(e));
    cudaFree(d);
}

int main()
{
    run<1, 0, 128>("bf16 SS no-swizzle N=128", 4);
    run<1, 1, 128>("bf16 SS 128B-swizzle N=128", 4);
    run<1, 0, 256>("bf16 SS no-swizzle N=256", 4);
    run<1, 1, 256>("bf16 SS 128B-swizzle N=256", 4);
    run<1, 0, 128, 1>("bf16 TS no-swizzle(B) N=128", 4);
    run<1, 1, 128, 1>("bf16 TS 128B-swizzle(B) N=128", 4);
    run<1, 0, 256, 1>("bf16 TS no-swizzle(B) N=256", 4);
    run<1, 0, 64>("bf16 SS no-swizzle N=64", 4);
    run<0, 0, 128>("tf32 SS no-swizzle N=128", 3);
    return 0;
}
