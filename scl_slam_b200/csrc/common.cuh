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

// ---- float atan, bit-exact with glibc 2.39 atanf (fdlibm s_atanf) --------------------------
// The reference's xy2theta calls atan(float) (descriptor.h:1357 with `using namespace std`,
// :19). The oracle checks this sequence of IEEE operations against libm on all 2^32 inputs
// (oracle/sc_oracle.cpp atanf_port, tests/test_oracle_ref.py). Every operation is an explicit
// round-to-nearest intrinsic so the compiler can neither contract to FMA nor reassociate.
__device__ __forceinline__ float scl_atanf(float x)
{
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const uint32_t hx = __float_as_uint(x), ix = hx & 0x7fffffffu;
    int id;
    if (ix >= 0x4c000000u) {
        if (ix > 0x7f800000u) return __fadd_rn(x, x);
        const float r = __fadd_rn(atanhi[3], atanlo[3]);
        return (hx >> 31) ? -r : r;
    }
    if (ix < 0x3ee00000u) {
        if (ix < 0x31000000u) return x;
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000u) {
            if (ix < 0x3f300000u) { id = 0; x = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, x), 1.0f), __fadd_rn(2.0f, x)); }
            else                  { id = 1; x = __fdiv_rn(__fsub_rn(x, 1.0f), __fadd_rn(x, 1.0f)); }
        } else {
            if (ix < 0x401c0000u) { id = 2; x = __fdiv_rn(__fsub_rn(x, 1.5f), __fadd_rn(1.0f, __fmul_rn(1.5f, x))); }
            else                  { id = 3; x = __fdiv_rn(-1.0f, x); }
        }
    }
    float z = __fmul_rn(x, x);
    const float w = __fmul_rn(z, z);
#define SCL_MA(a, b, c) __fadd_rn((a), __fmul_rn((b), (c))) /* a + b*c, two roundings */
    const float s1 = __fmul_rn(z, SCL_MA(3.3333334327e-01f, w, SCL_MA(1.4285714924e-01f, w, SCL_MA(9.0908870101e-02f, w,
                                  SCL_MA(6.6610731184e-02f, w, SCL_MA(4.9768779427e-02f, w, 1.6285819933e-02f))))));
    const float s2 = __fmul_rn(w, SCL_MA(-2.0000000298e-01f, w, SCL_MA(-1.1111110449e-01f, w, SCL_MA(-7.6918758452e-02f, w,
                                  SCL_MA(-5.8335702866e-02f, w, -3.6531571299e-02f)))));
#undef SCL_MA
    if (id < 0) return __fsub_rn(x, __fmul_rn(x, __fadd_rn(s1, s2)));
    z = __fsub_rn(atanhi[id], __fsub_rn(__fsub_rn(__fmul_rn(x, __fadd_rn(s1, s2)), atanlo[id]), x));
    return (hx >> 31) ? -z : z;
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
