// k3_knn_tc.cu — K3 on the 5th-generation tensor cores: a tcgen05 prefilter for the ring-key kNN,
// followed by an exact re-rank that keeps the result bit-identical to k3_knn.cu.
//
// Replaces (together with k3_knn.cu) nanoflann findNeighbors, /root/reference/include/descriptor.h:1714-1716,
// and libnabo knn, descriptor.h:1642.
//
// Phase A (knn_tc_kernel): the score S(q,k) = |k|^2 - 2 q.k of every (query, key) pair is one
// augmented GEMM on tcgen05.mma kind::tf32 with FP32 accumulators in TMEM. To get ~FP32 accuracy out
// of TF32 inputs every value is split x = hi + lo (hi = top 11 significant bits, lo = next 11):
//      S = [-2q_hi | 1 1 1] . [k_hi | n_hi n_mid n_lo]   (A1 x B_hi)
//        + [-2q_lo | 0 0 0] . [k_hi | ...]               (A2 x B_hi)
//        + [-2q_hi | 1 1 1] . [k_lo | 0 0 0]             (A1 x B_lo)
//   with |k|^2 = n_hi + n_mid + n_lo carried through the same contraction (three more columns).
//   One CTA per SM: it owns a 128-query tile (M = 128 = TMEM lanes) and one contiguous range of the
//   key matrix, and is warp-specialised:
//     warps 4-11 epilogue  : tcgen05.ld 32 columns at a time (thread = query = TMEM lane), compare
//                            against the thread's running threshold, push hits to a staging buffer,
//                            fold the staging buffers into per-thread sorted top-K' lists
//     warps 1-3  producers : stream raw keys (80 B each) from HBM, split hi/lo in registers, write
//                            the UMMA K-major no-swizzle core-matrix layout into shared memory
//     warp  0    MMA issuer: one thread issues the 9 tcgen05.mma per key tile and commits to
//                            mbarriers (smem stage free / accumulator ready)
//   Shared-memory stages and the two TMEM accumulator stages are handed around with mbarriers only.
//
// Phase B (knn_rerank_kernel): one warp per query recomputes the EXACT float distance (the
//   reference's accumulation order, k3_knn.cu) of every proposed key, selects the top-K by
//   (d2, id), and CERTIFIES the result: with T the smallest per-range cut-off score and eps the
//   prefilter's error bound, every key the prefilter dropped has exact d2 > T + |q|^2 - eps; if the
//   K-th selected distance is below that, no dropped key can belong to (or tie with) the top-K.
//   Queries that fail the certificate are appended to a list and redone by the exact kernel.
//
// Roofline: 2*R*Q*N flops against the tensor pipe, 4*R*N bytes against HBM; in practice bound by
// the TMEM read-out + compare of Q*N accumulators in the epilogue warps.
#include "common.cuh"
#include "kernels.h"

#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {

constexpr int kKPrime = 16;        /* proposals kept per (query, sub-range): a sorted list in REGISTERS */
constexpr int kStageCap = 16;      /* staging entries per thread: one 8-column group can add 8 */
constexpr int kEpiThreads = 256, kProdThreads = 96;   /* 8 epilogue warps (two per TMEM lane quadrant, half of the columns each); 384 threads: 168 registers each */
constexpr int kThreads = 32 + kProdThreads + kEpiThreads;
constexpr float kPadNorm = 1.0e30f, kThrInit = 1.0e29f;

// order-preserving float <-> signed int image (for atomicMin on scores that may be negative)
__device__ __forceinline__ int ordered_int(float f) { const int b = __float_as_int(f); return b ^ ((b >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// ---- tcgen05 / mbarrier PTX wrappers ---------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(scl_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(scl_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B stored contiguously; SBO = distance between 8-row
// groups, LBO = distance between the two 16-byte K chunks of one instruction (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;                      /* descriptor version for sm_100 */
    return d;                                    /* base offset 0, layout type 0 = SWIZZLE_NONE */
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}


// 64 consecutive fp32 columns of this warp's 32 TMEM lanes; asynchronous until tmem_wait64 on the same registers
#define SCL_R8(b) "%" #b
__device__ __forceinline__ void tmem_ld64_issue(uint32_t taddr, uint32_t (&r)[64])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
          "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
          "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
}
// wait for the outstanding tcgen05.ld; the registers are in/out operands so no use can be scheduled above the wait
__device__ __forceinline__ void tmem_wait64(uint32_t (&r)[64])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
          "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]),
          "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
          "+r"(r[30]), "+r"(r[31]), "+r"(r[32]), "+r"(r[33]), "+r"(r[34]), "+r"(r[35]), "+r"(r[36]), "+r"(r[37]), "+r"(r[38]), "+r"(r[39]),
          "+r"(r[40]), "+r"(r[41]), "+r"(r[42]), "+r"(r[43]), "+r"(r[44]), "+r"(r[45]), "+r"(r[46]), "+r"(r[47]), "+r"(r[48]), "+r"(r[49]),
          "+r"(r[50]), "+r"(r[51]), "+r"(r[52]), "+r"(r[53]), "+r"(r[54]), "+r"(r[55]), "+r"(r[56]), "+r"(r[57]), "+r"(r[58]), "+r"(r[59]),
          "+r"(r[60]), "+r"(r[61]), "+r"(r[62]), "+r"(r[63])
        :: "memory");
}
#undef SCL_R8
// 8 columns, synchronous (used only on the rare "some score beats the threshold" path)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   /* FMNMX3 */
    return r;
}

template <int R, int NT> struct TcCfg {
    static constexpr int G = ((R / 4 + 1) + 1) / 2 * 2;       /* 16-byte K chunks per row (incl. the norm chunk), even */
    static constexpr int KSTEPS = G / 2;                      /* tcgen05.mma instructions per product (K = 8 tf32 each) */
    static constexpr uint32_t A_LBO = 128 * 16, B_LBO = NT * 16, SBO = 128;
    static constexpr uint32_t A_BLOCK = 128 * G * 16, B_BLOCK = NT * G * 16;
    static constexpr uint32_t OFF_BAR = 0;                                    /* 8 mbarriers + tmem slot */
    static constexpr uint32_t OFF_A = 128;                                    /* A1, A2 */
    static constexpr uint32_t OFF_B = OFF_A + 2 * A_BLOCK;                    /* 2 stages x (B_hi, B_lo) */
    static constexpr uint32_t OFF_STG = OFF_B + 4 * B_BLOCK;                  /* staging [cap][256] val, idx */
    static constexpr uint32_t TOTAL = OFF_STG + 2 * kStageCap * kEpiThreads * 4;
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};

template <int R, int NT>
__global__ void __launch_bounds__(kThreads, 1) knn_tc_kernel(
    const float* __restrict__ qkeys, int Q, const float* __restrict__ keys, const float* __restrict__ knorm, int key_lo, int key_hi,
    int range_len, int n_ranges, int kprime, int sub_base, int n_sub_total, long long* __restrict__ times /* null, or [grid][8] role timers (SCL_TC_TIMES=1) */,
    int* __restrict__ g_thr /* [Q] shared thresholds (ordered-int image) */,
    float* __restrict__ prop_s /* [Q][n_ranges][K'] */, int32_t* __restrict__ prop_idx, float* __restrict__ prop_cut /* [Q][n_ranges] */)
{
    using C = TcCfg<R, NT>;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t *full = bars, *empty = bars + 2, *tfull = bars + 4, *tempty = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qtile = blockIdx.x / n_ranges, range = blockIdx.x % n_ranges;
    const int k_begin = key_lo + range * range_len;       /* this launch covers keys [key_lo, key_hi) */
    const int k_end = min(key_hi, k_begin + range_len);
    const int n_tiles = k_end > k_begin ? (k_end - k_begin + NT - 1) / NT : 0;

    // ---- one-time setup -----------------------------------------------------------------------
    if (threadIdx.x == 0) {
        scl_mbar_init(&full[0], kProdThreads / 32); scl_mbar_init(&full[1], kProdThreads / 32);
        scl_mbar_init(&empty[0], 1); scl_mbar_init(&empty[1], 1);
        scl_mbar_init(&tfull[0], 1); scl_mbar_init(&tfull[1], 1);
        scl_mbar_init(&tempty[0], kEpiThreads / 32); scl_mbar_init(&tempty[1], kEpiThreads / 32);
        scl_mbar_fence_init();
    }
    if (warp == 0) {   /* TMEM: 2 accumulator stages of NT fp32 columns */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(scl_smem_u32(tmem_slot)), "r"(2 * NT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   /* zero both B stages once (the norm chunk of B_lo and the padding chunk stay zero for ever) */
        uint4* z = reinterpret_cast<uint4*>(smem + C::OFF_B);
        for (uint32_t i = threadIdx.x; i < 4 * C::B_BLOCK / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    {   /* A operand: A1 = [-2 q_hi | 1 1 1 0 ...], A2 = [-2 q_lo | 0 ...], rows = the tile's 128 queries */
        float* A1 = reinterpret_cast<float*>(smem + C::OFF_A);
        float* A2 = reinterpret_cast<float*>(smem + C::OFF_A + C::A_BLOCK);
        for (int i = threadIdx.x; i < 128 * C::G; i += kThreads) {
            const int m = i % 128, g = i / 128;
            const int qi = qtile * 128 + m;
            float4 h = make_float4(0, 0, 0, 0), l = make_float4(0, 0, 0, 0);
            if (g < R / 4) {
                if (qi < Q) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(qkeys + (size_t)qi * R + 4 * g));
                    const float hx = tf32_trunc(x.x), hy = tf32_trunc(x.y), hz = tf32_trunc(x.z), hw = tf32_trunc(x.w);
                    h = make_float4(-2.0f * hx, -2.0f * hy, -2.0f * hz, -2.0f * hw);
                    l = make_float4(-2.0f * tf32_trunc(x.x - hx), -2.0f * tf32_trunc(x.y - hy), -2.0f * tf32_trunc(x.z - hz), -2.0f * tf32_trunc(x.w - hw));
                }
            } else if (g == R / 4) {
                h = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            }
            const uint32_t off = (uint32_t)g * C::A_LBO + (uint32_t)(m >> 3) * C::SBO + (uint32_t)(m & 7) * 16;
            *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(A1) + off) = h;
            *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(A2) + off) = l;
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Role -> warp mapping: the scheduler favours higher warp ids, so the epilogue (the busiest role) gets
    // the highest warps (4-11), the producers 1-3 and the single MMA-issuing thread warp 0. tcgen05.ld lets warp w touch
    // TMEM lanes 32*(w%4).., so epilogue warp w serves queries 32*(w%4)..32*(w%4)+31 of the tile.
    if (warp >= 4) {
        // ===== epilogue: thread = (query = TMEM lane, half of the columns) ================================================
        // A key is kept only if its score is below the thread's threshold. The threshold is the K'-th
        // smallest score seen so far for this query — by this CTA, or (through g_thr) by ANY CTA working
        // on the same query tile: each published value is backed by K' keys at or below it, so it bounds
        // the global K'-th smallest score from above and nothing in the true top-K' is ever dropped.
        constexpr int E = kEpiThreads;
        const int half = (warp - 4) >> 2;          /* which half of every tile's columns this warp examines */
        const int row = (warp & 3) * 32 + lane;    /* row of the tile = TMEM lane, 0..127 */
        const int t = half * 128 + row;            /* slot of this thread in the shared-memory lists */
        const int qi = qtile * 128 + row;
        const int sub = sub_base + 2 * range + half;       /* slot of this thread's proposal list among all sub-ranges */
        float* sv = reinterpret_cast<float*>(smem + C::OFF_STG);
        int* si = reinterpret_cast<int*>(sv + kStageCap * E);
        // The thread's K' best (score, key) so far live in registers, kept sorted by a branch-free bubble-through
        // insert (the shared-memory insertion sort this replaces cost ~10x more: a chain of dependent LDS/STS).
        float lv[kKPrime]; int li[kKPrime];
#pragma unroll
        for (int i = 0; i < kKPrime; i++) { lv[i] = kThrInit; li[i] = -1; }
        int cnt = 0;
        int n_slow = 0, n_push = 0, n_fold = 0;      /* developer counters (SCL_TC_TIMES) */
        float thr = kThrInit;
        int* my_gthr = g_thr + (qi < Q ? qi : 0);
        long long t_fold = 0;
        auto fold = [&]() {
            n_fold++; n_push += cnt;
            long long f0 = 0;
            if (times) f0 = clock64();
            const float before = thr;
            for (int s = 0; s < cnt; s++) {
                float val = sv[s * E + t];
                int id = si[s * E + t];
                if (!(val < thr)) continue;
#pragma unroll
                for (int i = 0; i < kKPrime; i++) {
                    const bool lt = val < lv[i];
                    const float tv = lt ? lv[i] : val; const int ti = lt ? li[i] : id;
                    lv[i] = lt ? val : lv[i]; li[i] = lt ? id : li[i];
                    val = tv; id = ti;
                }
                thr = fminf(thr, lv[kKPrime - 1]);     /* never loosen a threshold learned from other CTAs */
            }
            cnt = 0;
            if (thr < before && qi < Q) atomicMin(my_gthr, ordered_int(thr));
            if (times) t_fold += clock64() - f0;
        };
        // One 64-column TMEM load is always in flight while the previous 64 columns are examined. The common
        // case is "nothing below the threshold": a min-tree (FMNMX3) over the 64 scores and one compare. Only
        // the 8-column groups whose minimum beats the threshold are examined element by element.
        uint32_t va[64];
        long long t_ld = 0, t_slow = 0;                      /* developer probes (SCL_TC_TIMES) */
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        auto examine = [&](uint32_t (&r)[64], uint32_t col_first, int key_first) {
            unsigned mask = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float x0 = __uint_as_float(r[8 * j]), x1 = __uint_as_float(r[8 * j + 1]), x2 = __uint_as_float(r[8 * j + 2]),
                            x3 = __uint_as_float(r[8 * j + 3]), x4 = __uint_as_float(r[8 * j + 4]), x5 = __uint_as_float(r[8 * j + 5]),
                            x6 = __uint_as_float(r[8 * j + 6]), x7 = __uint_as_float(r[8 * j + 7]);
                const float gj = fminf(fmin3(fmin3(x0, x1, x2), fmin3(x3, x4, x5), x6), x7);
                mask |= (gj < thr ? 1u : 0u) << j;
            }
            unsigned wm = __reduce_or_sync(0xffffffffu, mask);
            if (wm) n_slow++;
            long long e0 = 0;
            if (times && wm) e0 = clock64();
            const bool was_slow = wm != 0;
#pragma unroll 1
            while (wm) {
                const int j = __ffs(wm) - 1;                 /* warp-uniform: the switch below does not diverge */
                wm &= wm - 1;
                float v[8];
#define SCL_GROUP(J) case J: _Pragma("unroll") for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[8 * J + i]); break;
                switch (j) { SCL_GROUP(0) SCL_GROUP(1) SCL_GROUP(2) SCL_GROUP(3) SCL_GROUP(4) SCL_GROUP(5) SCL_GROUP(6) default: SCL_GROUP(7) }
#undef SCL_GROUP
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (v[i] < thr) { sv[cnt * E + t] = v[i]; si[cnt * E + t] = key_first + 8 * j + i; cnt++; }
                if (__any_sync(0xffffffffu, cnt > kStageCap - 8)) fold();     /* all lanes fold together: amortised */
            }
            if (times && was_slow) t_slow += clock64() - e0;
        };
        long long tw = 0, tp = 0, c0 = clock64();
        int shared_thr = __ldcg(my_gthr);                      /* then fetched one tile ahead: its L2 latency is never exposed */
        for (int tile = 0; tile < n_tiles; tile++) {
            const int a = tile & 1; const uint32_t ph = (tile >> 1) & 1;
            scl_mbar_wait(&tfull[a], ph);
            tc_fence_after();
            if (times) { const long long c1 = clock64(); tw += c1 - c0; c0 = c1; }
            thr = fminf(thr, ordered_float(shared_thr));
            shared_thr = __ldcg(my_gthr);
            const int key0 = k_begin + tile * NT;
            const uint32_t col0 = lane_base + (uint32_t)(a * NT);
#pragma unroll 1
            for (int c = 0; c < NT / 128; c++) {               /* this warp's half of the tile, 64 columns at a time */
                const int cc = half * (NT / 128) + c;
                long long d0 = 0;
                if (times) d0 = clock64();
                tmem_ld64_issue(col0 + cc * 64, va);
                tmem_wait64(va);
                if (times) { const long long d1 = clock64(); t_ld += d1 - d0; }
                examine(va, col0 + cc * 64, key0 + cc * 64);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[a]);
            if (times) { const long long c1 = clock64(); tp += c1 - c0; c0 = c1; }
        }
        if (times) {
            const int ws = __reduce_add_sync(0xffffffffu, n_slow), wp = __reduce_add_sync(0xffffffffu, n_push + cnt);
            const unsigned any_slow_chunks = 0;
            (void)any_slow_chunks;
            if (t == 0) { times[blockIdx.x * 16 + 0] = tw; times[blockIdx.x * 16 + 1] = tp; times[blockIdx.x * 16 + 8] = ws; times[blockIdx.x * 16 + 9] = wp; times[blockIdx.x * 16 + 10] = n_fold; times[blockIdx.x * 16 + 11] = t_ld; times[blockIdx.x * 16 + 12] = t_slow; times[blockIdx.x * 16 + 13] = t_fold; }
        }
        fold();
        if (qi < Q) {
            const size_t o = ((size_t)qi * n_sub_total + sub) * kKPrime;
#pragma unroll
            for (int i = 0; i < kKPrime; i++) {
                prop_s[o + i] = li[i] >= 0 ? lv[i] : __int_as_float(0x7f800000);
                prop_idx[o + i] = li[i];
            }
            /* cut-off of this sub-range: every key NOT proposed had S >= the threshold in force when it was
             * examined >= the final threshold (thresholds only fall); inf if nothing was ever dropped */
            prop_cut[(size_t)qi * n_sub_total + sub] = thr < kThrInit ? thr : __int_as_float(0x7f800000);
        }
    } else if (warp >= 1) {
        // ===== producers: raw keys -> hi/lo split -> UMMA core-matrix layout =======================
        // The raw keys of tile t+1 are already in flight (registers) while tile t is split and stored.
        const int p = threadIdx.x - 32;            /* 0..95 */
        constexpr int KPT = (NT + kProdThreads - 1) / kProdThreads;     /* keys per thread per tile */
        float4 xa[KPT][R / 4];
        float na[KPT];
        auto load_tile = [&](int tile, float4 (&x)[KPT][R / 4], float (&n)[KPT]) {
#pragma unroll
            for (int mm = 0; mm < KPT; mm++) {
                const int key = k_begin + tile * NT + p + mm * kProdThreads;
                if (tile < n_tiles && key < k_end && p + mm * kProdThreads < NT) {
                    const float4* src = reinterpret_cast<const float4*>(keys + (size_t)key * R);
#pragma unroll
                    for (int g = 0; g < R / 4; g++) x[mm][g] = __ldg(src + g);
                    n[mm] = __ldg(knorm + key);
                } else {
#pragma unroll
                    for (int g = 0; g < R / 4; g++) x[mm][g] = make_float4(0, 0, 0, 0);
                    n[mm] = kPadNorm;                     /* padded rows can never be proposed */
                }
            }
        };
        auto store_tile = [&](int s, const float4 (&x)[KPT][R / 4], const float (&n)[KPT]) {
            unsigned char* Bhi = smem + C::OFF_B + (uint32_t)s * 2 * C::B_BLOCK;
            unsigned char* Blo = Bhi + C::B_BLOCK;
#pragma unroll
            for (int mm = 0; mm < KPT; mm++) {
                const int m = p + mm * kProdThreads;
                if (m >= NT) continue;
                const uint32_t row_off = (uint32_t)(m >> 3) * C::SBO + (uint32_t)(m & 7) * 16;
#pragma unroll
                for (int g = 0; g < R / 4; g++) {
                    const float hx = tf32_trunc(x[mm][g].x), hy = tf32_trunc(x[mm][g].y), hz = tf32_trunc(x[mm][g].z), hw = tf32_trunc(x[mm][g].w);
                    *reinterpret_cast<float4*>(Bhi + (uint32_t)g * C::B_LBO + row_off) = make_float4(hx, hy, hz, hw);
                    *reinterpret_cast<float4*>(Blo + (uint32_t)g * C::B_LBO + row_off) =
                        make_float4(tf32_trunc(x[mm][g].x - hx), tf32_trunc(x[mm][g].y - hy), tf32_trunc(x[mm][g].z - hz), tf32_trunc(x[mm][g].w - hw));
                }
                const float n_hi = tf32_trunc(n[mm]), r1 = n[mm] - n_hi, n_mid = tf32_trunc(r1), n_lo = tf32_trunc(r1 - n_mid);
                *reinterpret_cast<float4*>(Bhi + (uint32_t)(R / 4) * C::B_LBO + row_off) = make_float4(n_hi, n_mid, n_lo, 0.0f);
            }
        };
        long long tw = 0, tp = 0, c0 = clock64();
        load_tile(0, xa, na);
#pragma unroll 1
        for (int tile = 0; tile < n_tiles; tile++) {
            const int s = tile & 1; const uint32_t ph = (tile >> 1) & 1;
            scl_mbar_wait(&empty[s], ph ^ 1u);
            if (times) { const long long c1 = clock64(); tw += c1 - c0; c0 = c1; }
            store_tile(s, xa, na);
            fence_async_smem();                    /* generic-proxy writes -> visible to the tensor core (async proxy) */
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            load_tile(tile + 1, xa, na);           /* in flight while we wait for the next free stage */
            if (times) { const long long c1 = clock64(); tp += c1 - c0; c0 = c1; }
        }
        if (times && p == 0) { times[blockIdx.x * 16 + 2] = tw; times[blockIdx.x * 16 + 3] = tp; }
    } else {
        // ===== MMA issuer: one thread ==============================================================
        if (lane == 0) {
            const uint32_t a1 = scl_smem_u32(smem + C::OFF_A), a2 = a1 + C::A_BLOCK;
            long long t_te = 0, t_fu = 0, t_is = 0, c0 = clock64();
            for (int tile = 0; tile < n_tiles; tile++) {
                const int s = tile & 1; const uint32_t ph = (tile >> 1) & 1;
                scl_mbar_wait(&tempty[s], ph ^ 1u);        /* accumulator stage drained by the epilogue */
                if (times) { const long long c1 = clock64(); t_te += c1 - c0; c0 = c1; }
                scl_mbar_wait(&full[s], ph);               /* operands written */
                if (times) { const long long c1 = clock64(); t_fu += c1 - c0; c0 = c1; }
                tc_fence_after();
                const uint32_t bhi = scl_smem_u32(smem + C::OFF_B) + (uint32_t)s * 2 * C::B_BLOCK, blo = bhi + C::B_BLOCK;
                const uint32_t d = tmem_base + (uint32_t)(s * NT);
                uint32_t acc = 0;
                {
#pragma unroll
                for (int k = 0; k < C::KSTEPS; k++) {
                    tc_mma_tf32(d, make_desc(a1 + 2 * k * C::A_LBO, C::A_LBO, C::SBO), make_desc(bhi + 2 * k * C::B_LBO, C::B_LBO, C::SBO), C::IDESC, acc);
                    acc = 1;
                }
#pragma unroll
                for (int k = 0; k < C::KSTEPS; k++)
                    tc_mma_tf32(d, make_desc(a2 + 2 * k * C::A_LBO, C::A_LBO, C::SBO), make_desc(bhi + 2 * k * C::B_LBO, C::B_LBO, C::SBO), C::IDESC, 1);
#pragma unroll
                for (int k = 0; k < C::KSTEPS; k++)
                    tc_mma_tf32(d, make_desc(a1 + 2 * k * C::A_LBO, C::A_LBO, C::SBO), make_desc(blo + 2 * k * C::B_LBO, C::B_LBO, C::SBO), C::IDESC, 1);
                }
                tc_commit(&empty[s]);                      /* smem stage reusable once these MMAs retire */
                tc_commit(&tfull[s]);                      /* accumulator ready for the epilogue */
                if (times) { const long long c1 = clock64(); t_is += c1 - c0; c0 = c1; }
            }
            if (times) { times[blockIdx.x * 16 + 4] = t_te; times[blockIdx.x * 16 + 5] = t_fu; times[blockIdx.x * 16 + 6] = t_is; times[blockIdx.x * 16 + 7] = n_tiles; }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * NT) : "memory");
    }
}

// Bootstrap: before any tensor-core work, every query gets a valid starting threshold from a strided sample of
// kBootKeys keys scored on the CUDA cores. Each lane keeps the 4 smallest scores of its share; the K'-th smallest of
// those 128 scores is >= the K'-th smallest of the sample, hence of the whole database. That bound (+ the
// prefilter's error bound, so that it also holds for the tensor-core scores) removes most of the start-up
// transient of the streaming top-K' in the sample pass.
constexpr int kBootKeys = 4096;
template <int R>
__global__ void __launch_bounds__(256) knn_bootstrap_kernel(const float* __restrict__ qkeys, int Q, const float* __restrict__ keys,
                                                            const float* __restrict__ knorm, const float* __restrict__ kn2max, int n_db,
                                                            int kprime, int* __restrict__ g_thr)
{
    __shared__ __align__(16) float sk[256 * R];          /* rows of R floats: LDS.128 reads at an 80-byte lane stride are conflict free */
    __shared__ float sn[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * 8 + warp;
    float q[R];
#pragma unroll
    for (int d = 0; d < R; d++) q[d] = qi < Q ? __ldg(qkeys + (size_t)qi * R + d) : 0.0f;
    const int n_s = n_db < kBootKeys ? n_db : kBootKeys;
    const long long stride = n_db / n_s;                 /* sample key j is database key j*stride */
    float b0 = kThrInit, b1 = kThrInit, b2 = kThrInit, b3 = kThrInit;
    for (int base = 0; base < n_s; base += 256) {
        __syncthreads();
        {
            const int j = base + threadIdx.x;
            if (j < n_s) {
                const size_t key = (size_t)j * stride;
#pragma unroll
                for (int g = 0; g < R / 4; g++)
                    reinterpret_cast<float4*>(sk + threadIdx.x * R)[g] = __ldg(reinterpret_cast<const float4*>(keys + key * R) + g);
                sn[threadIdx.x] = __ldg(knorm + key);
            }
        }
        __syncthreads();
        const int nk = min(256, n_s - base);
        for (int j = lane; j < nk; j += 32) {
            float dot = 0.0f;
#pragma unroll
            for (int g = 0; g < R / 4; g++) {
                const float4 k4 = reinterpret_cast<const float4*>(sk + j * R)[g];
                dot = fmaf(q[4 * g], k4.x, dot); dot = fmaf(q[4 * g + 1], k4.y, dot);
                dot = fmaf(q[4 * g + 2], k4.z, dot); dot = fmaf(q[4 * g + 3], k4.w, dot);
            }
            float x = fmaf(-2.0f, dot, sn[j]);
            float y;
            y = fminf(b0, x); x = fmaxf(b0, x); b0 = y;
            y = fminf(b1, x); x = fmaxf(b1, x); b1 = y;
            y = fminf(b2, x); x = fmaxf(b2, x); b2 = y;
            b3 = fminf(b3, x);
        }
    }
    /* K'-th smallest among the 32 x 4 kept scores (each lane's four are sorted): a K'-step k-way merge. If a lane
     * held more than four of the true top-K', the value found is only larger, so it stays a valid upper bound. */
    float picked = kThrInit;
    for (int r = 0; r < kprime; r++) {
        float w = b0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) w = fminf(w, __shfl_xor_sync(0xffffffffu, w, off));
        picked = w;
        const unsigned who = __ballot_sync(0xffffffffu, b0 == w);
        if (lane == __ffs(who) - 1) { b0 = b1; b1 = b2; b2 = b3; b3 = kThrInit; }   /* pop this lane's head */
    }
    if (lane == 0 && qi < Q && picked < kThrInit) {
        float qn = 0.0f;
#pragma unroll
        for (int d = 0; d < R; d++) qn = fmaf(q[d], q[d], qn);
        const float sn2 = sqrtf(qn) + sqrtf(__ldg(kn2max));
        atomicMin(g_thr + qi, ordered_int(picked + 3.0517578125e-05f * sn2 * sn2));   /* + 2^-15 (|q|+|k|max)^2 */
    }
}

// Between the sample pass and the main pass: the K'-th smallest score over the sample keys (= over the union of
// the sample pass' proposal lists) becomes every CTA's starting threshold for that query. One warp per query.
__global__ void __launch_bounds__(128) knn_sample_thr_kernel(const float* __restrict__ prop_s, int Q, int n_sub_total, int n_sub_sample,
                                                             int kprime, int* __restrict__ g_thr)
{
    /* every proposal list is sorted ascending: a K'-step k-way merge, each lane holding the heads of its lists */
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= Q) return;
    const float* ps = prop_s + (size_t)qi * n_sub_total * kprime;
    const float inf = __int_as_float(0x7f800000);
    constexpr int kHeads = 10;                          /* up to 320 lists (148 ranges x 2 halves) */
    int head[kHeads];
    float hv[kHeads];
#pragma unroll
    for (int h = 0; h < kHeads; h++) {
        const int l = lane + 32 * h;
        head[h] = 0;
        hv[h] = l < n_sub_sample ? ps[(size_t)l * kprime] : inf;
    }
    float kth = inf;
    for (int r = 0; r < kprime; r++) {
        float bv = inf; int bh = -1;
#pragma unroll
        for (int h = 0; h < kHeads; h++) if (hv[h] < bv) { bv = hv[h]; bh = h; }
        float wv = bv;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) wv = fminf(wv, __shfl_xor_sync(0xffffffffu, wv, off));
        if (!(wv < inf)) { kth = inf; break; }          /* fewer than K' sample keys: no threshold */
        kth = wv;
        const unsigned who = __ballot_sync(0xffffffffu, bh >= 0 && bv == wv);
        if (lane == __ffs(who) - 1) {
#pragma unroll
            for (int h = 0; h < kHeads; h++)
                if (h == bh) {
                    head[h]++;
                    hv[h] = head[h] < kprime ? ps[(size_t)(lane + 32 * h) * kprime + head[h]] : inf;
                }
        }
    }
    if (lane == 0 && kth < inf) atomicMin(g_thr + qi, ordered_int(kth));
}

template <int METRIC>
__device__ __forceinline__ float exact_d2(const float* __restrict__ q, const float* __restrict__ k, int R)
{
    float result = 0.0f;
    if (METRIC == 0) {
        int d = 0;
        for (; d + 3 < R; d += 4) {
            const float d0 = __fsub_rn(q[d], k[d]), d1 = __fsub_rn(q[d + 1], k[d + 1]), d2 = __fsub_rn(q[d + 2], k[d + 2]), d3 = __fsub_rn(q[d + 3], k[d + 3]);
            const float g = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
            result = __fadd_rn(result, g);
        }
        for (; d < R; d++) { const float d0 = __fsub_rn(q[d], k[d]); result = __fadd_rn(result, __fmul_rn(d0, d0)); }
    } else {
        for (int d = 0; d < R; d++) { const float d0 = __fsub_rn(q[d], k[d]); result = __fadd_rn(result, __fmul_rn(d0, d0)); }
    }
    return result;
}

// Phase B: exact re-rank + certificate. One warp per query; n_cand = n_ranges * K' proposals, of which only those
// scoring at or below the cut (a few dozen) can matter — see below — and are compacted into shared memory.
constexpr int kMaxSurvivors = 256;
template <int METRIC>
__global__ void __launch_bounds__(128) knn_rerank_kernel(const float* __restrict__ qkeys, int Q, const float* __restrict__ keys, int R, int K,
                                                         int n_ranges, int kprime, const float* __restrict__ prop_s,
                                                         const int32_t* __restrict__ prop_idx, const float* __restrict__ prop_cut,
                                                         const float* __restrict__ kn2max, int id_mul, int id_add,
                                                         int32_t* __restrict__ out_ids, float* __restrict__ out_d2,
                                                         int32_t* __restrict__ fail_list, int* __restrict__ fail_count)
{
    __shared__ int s_id[4][kMaxSurvivors];
    __shared__ float s_d[4][kMaxSurvivors];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int qi = blockIdx.x * (blockDim.x >> 5) + w;
    if (qi >= Q) return;
    const int n_cand = n_ranges * kprime;
    const float* q = qkeys + (size_t)qi * R;
    const int32_t* pidx = prop_idx + (size_t)qi * n_cand;
    const float* ps = prop_s + (size_t)qi * n_cand;
    const float inf = __int_as_float(0x7f800000);
    float cut = inf;
    for (int r = lane; r < n_ranges; r += 32) cut = fminf(cut, prop_cut[(size_t)qi * n_ranges + r]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cut = fminf(cut, __shfl_xor_sync(0xffffffffu, cut, off));
    /* T = the K'-th smallest score among ALL proposals (a K'-step k-way merge over the sorted lists). Everything that
     * is not evaluated exactly below — keys the prefilter dropped (score >= cut) and proposals scoring above T — has
     * score >= min(cut, T), hence exact d2 > min(cut, T) + |q|^2 - eps; the certificate demands that this exceeds
     * the K-th selected distance. If it fails, the query is redone exactly. So only ~K' survivors are evaluated. */
    {
        constexpr int kHeads = 10;                      /* up to 320 lists */
        int head[kHeads]; float hv[kHeads];
#pragma unroll
        for (int h = 0; h < kHeads; h++) { const int l = lane + 32 * h; head[h] = 0; hv[h] = l < n_ranges ? ps[(size_t)l * kprime] : inf; }
        float T = inf;
        for (int r = 0; r < kprime; r++) {
            float bv = inf; int bh = -1;
#pragma unroll
            for (int h = 0; h < kHeads; h++) if (hv[h] < bv) { bv = hv[h]; bh = h; }
            float wv = bv;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) wv = fminf(wv, __shfl_xor_sync(0xffffffffu, wv, off));
            if (!(wv < inf)) { T = inf; break; }        /* fewer than K' proposals in all: keep them all */
            T = wv;
            const unsigned who = __ballot_sync(0xffffffffu, bh >= 0 && bv == wv);
            if (lane == __ffs(who) - 1) {
#pragma unroll
                for (int h = 0; h < kHeads; h++)
                    if (h == bh) { head[h]++; hv[h] = head[h] < kprime ? ps[(size_t)(lane + 32 * h) * kprime + head[h]] : inf; }
            }
        }
        cut = fminf(cut, T);
    }
    /* the lists are sorted ascending: a list is read only while its entries are at or below the cut */
    __shared__ int s_count[4];
    if (lane == 0) s_count[w] = 0;
    __syncwarp();
    for (int l = lane; l < n_ranges; l += 32) {
        for (int i = 0; i < kprime; i++) {
            const float sc = ps[(size_t)l * kprime + i];
            if (sc > cut || !(sc < inf)) break;
            const int id = pidx[(size_t)l * kprime + i];
            if (id < 0) break;
            const int pos = atomicAdd(&s_count[w], 1);
            if (pos < kMaxSurvivors) s_id[w][pos] = id;
        }
    }
    __syncwarp();
    int n_surv = s_count[w]; bool overflow = false;
    if (n_surv > kMaxSurvivors) { overflow = true; n_surv = kMaxSurvivors; }
    __syncwarp();
    for (int c = lane; c < n_surv; c += 32) {
        float d = exact_d2<METRIC>(q, keys + (size_t)s_id[w][c] * R, R);
        if (METRIC == 1 && !(d > FLT_EPSILON)) d = inf;          /* libnabo self-match rule */
        if (!(d < (METRIC == 0 ? FLT_MAX : inf))) d = inf;       /* never accepted by the trees */
        s_d[w][c] = d;
    }
    __syncwarp();
    /* K rounds: smallest (d2, id) strictly after the previous pick */
    float pd = -1.0f; int pi = -1; float dK = 0.0f; int found = 0;
    for (int r = 0; r < K; r++) {
        float bd = inf; int bi = 0x7fffffff;
        for (int c = lane; c < n_surv; c += 32) {
            const float d = s_d[w][c];
            if (!(d < inf)) continue;
            const int id = s_id[w][c] * id_mul + id_add;
            if (d < pd || (d == pd && id <= pi)) continue;
            if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, off); const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        const bool ok = bi != 0x7fffffff;
        if (lane == 0) { out_ids[(size_t)qi * K + r] = ok ? bi : -1; out_d2[(size_t)qi * K + r] = ok ? bd : FLT_MAX; }
        if (ok) { pd = bd; pi = bi; dK = bd; found++; }
        else break;
    }
    if (lane == 0) {
        for (int r = found; r < K; r++) { out_ids[(size_t)qi * K + r] = -1; out_d2[(size_t)qi * K + r] = FLT_MAX; }
        bool certified = !overflow;
        if (cut < inf) {
            /* dropped keys have S >= cut, i.e. exact d2 > cut + |q|^2 - eps */
            float qn = 0.0f;
            for (int d = 0; d < R; d++) qn = fmaf(q[d], q[d], qn);
            const float sn = sqrtf(qn) + sqrtf(__ldg(kn2max));
            const float eps = 1.52587890625e-05f * sn * sn + 2.0e-6f * dK;       /* 2^-16 (|q|+|k|max)^2 + exact-side rounding */
            certified = certified && (found == K) && (dK + eps < cut + qn);
        }
        if (!certified) fail_list[atomicAdd(fail_count, 1)] = qi;
    }
}

} // namespace

bool scl_knn_tc_supported(int R) { return R == 20 || R == 40; }

int scl_knn_tc_ranges(int Q)
{
    const int tiles = (Q + 127) / 128;
    int r = SCL_NUM_SMS / tiles;
    return r < 1 ? 1 : r;
}

int scl_knn_tc_kprime(int K) { (void)K; return kKPrime; }   /* fixed: the list is a register array */

template <int R, int NT>
static cudaError_t launch_tc(const float* qkeys, int Q, const float* keys, const float* knorm, int key_lo, int key_hi, int n_ranges,
                             int kprime, int sub_base, int n_sub_total, int* g_thr, float* prop_s, int32_t* prop_idx, float* prop_cut,
                             cudaStream_t stream)
{
    using C = TcCfg<R, NT>;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(knn_tc_kernel<R, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::TOTAL);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    int range_len = (key_hi - key_lo + n_ranges - 1) / n_ranges;
    range_len = (range_len + NT - 1) / NT * NT;
    if (range_len < NT) range_len = NT;
    const int tiles = (Q + 127) / 128;
    long long* times = nullptr;
    const bool want_times = getenv("SCL_TC_TIMES") != nullptr;       /* developer aid: per-role cycle counters on stderr */
    if (want_times) { cudaMalloc(&times, (size_t)tiles * n_ranges * 16 * sizeof(long long)); cudaMemset(times, 0, (size_t)tiles * n_ranges * 128); }
    knn_tc_kernel<R, NT><<<tiles * n_ranges, kThreads, C::TOTAL, stream>>>(qkeys, Q, keys, knorm, key_lo, key_hi, range_len, n_ranges, kprime,
                                                                          sub_base, n_sub_total, times, g_thr, prop_s, prop_idx, prop_cut);
    if (want_times) {
        const int nb = tiles * n_ranges;
        std::vector<long long> h((size_t)nb * 16);
        cudaStreamSynchronize(stream);
        cudaMemcpy(h.data(), times, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        double a[16] = {0};
        for (int b = 0; b < nb; b++) for (int i = 0; i < 16; i++) a[i] += (double)h[(size_t)b * 16 + i] / nb;
        fprintf(stderr, "[tc keys %d..%d] tiles/CTA %.0f | cycles per tile: epilogue wait %.0f work %.0f (tmem %.0f, slow path %.0f of which folds %.0f) | producer wait %.0f work %.0f | "
                        "mma wait_tmem %.0f wait_operands %.0f issue %.0f | per warp: slow chunks %.0f, pushes/lane %.1f, folds %.0f\n",
                key_lo, key_hi, a[7], a[0] / a[7], a[1] / a[7], a[11] / a[7], a[12] / a[7], a[13] / a[7], a[2] / a[7], a[3] / a[7], a[4] / a[7], a[5] / a[7], a[6] / a[7],
                a[8] / 32, a[9] / 32, a[10]);
        cudaFree(times);
    }
    return cudaGetLastError();
}

// The sample pass covers the first keys of the database (1/16 of it, at most 65,536): its only purpose is to hand the
// main pass a tight starting threshold per query, which removes the start-up transient of the streaming top-K'
// (~K' ln(n/K') hits per thread) from 15/16 of the stream.
int scl_knn_tc_sample(int n_db)
{
    int n_s = n_db / 16;
    if (n_s > 65536) n_s = 65536;
    n_s &= ~255;
    return n_s < 4096 ? 0 : n_s;
}

cudaError_t scl_launch_knn_tc(const float* qkeys, int Q, const float* keys, const float* knorm, const float* kn2max, int n_db, int R, int K,
                              int metric, int id_mul, int id_add, KnnTcWorkspace ws, int32_t* out_ids, float* out_d2,
                              int32_t* fail_list, int* fail_count, cudaStream_t stream)
{
    if (Q <= 0) return cudaSuccess;
    const int n_ranges = scl_knn_tc_ranges(Q);
    const int kprime = scl_knn_tc_kprime(K);
    if (K > kprime - 2) return cudaErrorInvalidValue;
    const int n_sub = 2 * n_ranges;                       /* proposal lists per launch: two column halves per CTA */
    const int n_s = scl_knn_tc_sample(n_db);
    const int n_sub_total = n_s > 0 ? 2 * n_sub : n_sub;
    if ((size_t)Q * n_sub_total * kprime > ws.capacity) return cudaErrorInvalidValue;
    cudaError_t err = cudaMemsetAsync(fail_count, 0, sizeof(int), stream);
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(ws.g_thr, 0x7f, (size_t)Q * sizeof(int), stream);   /* 0x7f7f7f7f = 3.4e38: "no threshold yet" */
    if (err != cudaSuccess) return err;
    if (R != 20 && R != 40) return cudaErrorNotSupported;
    if (n_db >= 4 * kBootKeys) {
        if (R == 20) knn_bootstrap_kernel<20><<<(Q + 7) / 8, 256, 0, stream>>>(qkeys, Q, keys, knorm, kn2max, n_db, kprime, ws.g_thr);
        else knn_bootstrap_kernel<40><<<(Q + 7) / 8, 256, 0, stream>>>(qkeys, Q, keys, knorm, kn2max, n_db, kprime, ws.g_thr);
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    for (int pass = (n_s > 0 ? 0 : 1); pass < 2; pass++) {
        const int lo = pass == 0 ? 0 : n_s, hi = pass == 0 ? n_s : n_db;
        const int sub_base = (pass == 1 && n_s > 0) ? n_sub : 0;
        if (R == 20) err = launch_tc<20, 256>(qkeys, Q, keys, knorm, lo, hi, n_ranges, kprime, sub_base, n_sub_total, ws.g_thr, ws.prop_s, ws.prop_idx, ws.prop_cut, stream);
        else err = launch_tc<40, 128>(qkeys, Q, keys, knorm, lo, hi, n_ranges, kprime, sub_base, n_sub_total, ws.g_thr, ws.prop_s, ws.prop_idx, ws.prop_cut, stream);
        if (err != cudaSuccess) return err;
        if (pass == 0) {
            knn_sample_thr_kernel<<<(Q + 3) / 4, 128, 0, stream>>>(ws.prop_s, Q, n_sub_total, n_sub, kprime, ws.g_thr);
            err = cudaGetLastError();
            if (err != cudaSuccess) return err;
        }
    }
    const int warps = 4;
    if (metric == 0)
        knn_rerank_kernel<0><<<(Q + warps - 1) / warps, warps * 32, 0, stream>>>(qkeys, Q, keys, R, K, n_sub_total, kprime, ws.prop_s, ws.prop_idx, ws.prop_cut,
                                                                                kn2max, id_mul, id_add, out_ids, out_d2, fail_list, fail_count);
    else
        knn_rerank_kernel<1><<<(Q + warps - 1) / warps, warps * 32, 0, stream>>>(qkeys, Q, keys, R, K, n_sub_total, kprime, ws.prop_s, ws.prop_idx, ws.prop_cut,
                                                                                kn2max, id_mul, id_add, out_ids, out_d2, fail_list, fail_count);
    return cudaGetLastError();
}
