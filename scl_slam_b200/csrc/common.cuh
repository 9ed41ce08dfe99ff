// common.cuh — shared device helpers for the sm_100a Scan Context kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define SCL_NUM_SMS 148 /* B200: 2 dies x 74 SMs; grids are sized in multiples of this */

// Kernels of different query lanes run on the same SM at the same time only if the SM's L1 / shared-memory split suits all of
// them: every kernel of the query path asks for the largest shared-memory carveout once per device, so a small-footprint kernel
// never forces the SM to drain and re-partition between lanes.
#include <atomic>
struct SclOncePerDevice {
    std::atomic<unsigned long long> seen{0};
    bool first()
    {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) return false;
        const unsigned long long bit = 1ull << (d & 63);
        return !(seen.fetch_or(bit) & bit);
    }
};
// Load a kernel's code on the current device now. With CUDA's lazy module loading the FIRST launch of a kernel may need to
// synchronise the context; a host thread that has just enqueued an exchange kernel (which waits for a peer) and then
// launches a not-yet-loaded kernel on the same device would block before it can give the peer its work. scl_create loads
// every kernel of the query path up front (scl_preload_*), so no launch on that path ever synchronises.
#define SCL_TOUCH(kernel) do { cudaFuncAttributes _a; (void)cudaFuncGetAttributes(&_a, kernel); } while (0)
#define SCL_PREFER_SMEM(kernel)                                                                                              \
    do {                                                                                                                     \
        static SclOncePerDevice _once;                                                                                       \
        if (_once.first()) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
    } while (0)

// ---- order-preserving float <-> uint key (for atomicMax on bins) ---------------------------
// key(a) < key(b)  <=>  a < b for all non-NaN floats (with -0.0 < +0.0).
__device__ __forceinline__ uint32_t scl_float_key(float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float scl_key_float(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// int(ceil(v)) as x86-64 cvttsd2si does it: NaN / out of range -> INT_MIN (oracle: ceil_to_int)
__device__ __forceinline__ int scl_ceil_to_int(double v)
{
    const double c = ceil(v);
    if (!(c >= -2147483648.0 && c <= 2147483647.0)) return (int)0x80000000;
    return (int)c;
}

// ---- mbarrier + 1-D bulk copy (TMA engine, SASS UBLKCP) -------------------------------------
__device__ __forceinline__ uint32_t scl_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void scl_mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(scl_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void scl_mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void scl_mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(scl_smem_u32(bar)), "r"(bytes) : "memory");
}
// Wait for an mbarrier phase. The try_wait carries a suspend-time hint, so a waiting warp SLEEPS in
// hardware until the phase flips instead of spinning: spinning warps steal issue slots from the working
// warps of the same scheduler (measured: a 4x slowdown of the epilogue warps in knn_tc_kernel).
// Bounded: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
__device__ __forceinline__ void scl_mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = scl_smem_u32(bar);
    unsigned long long t0 = 0;
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity), "r"(1000000u) : "memory");
        if (done) return;
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ull) __trap();              /* 4 s */
    }
}
// global -> shared bulk copy; bytes must be a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void scl_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(scl_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(scl_smem_u32(bar)) : "memory");
}
