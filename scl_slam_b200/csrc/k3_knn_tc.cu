// k3_knn_tc.cu — K3 on the 5th-generation tensor cores, second generation: a BF16x3 tcgen05 prefilter fed by
// TMA from a pre-split key image, followed by an exact re-rank that keeps the result bit-identical to k3_knn.cu.
//
// Replaces (together with k3_knn.cu) nanoflann findNeighbors, /root/reference/include/descriptor.h:1714-1716,
// and libnabo knn, descriptor.h:1642.
//
// Score. S(q,k) = |k|^2 - 2 q.k orders the keys of one query exactly like the squared distance. Every float is
// split into bfloat16 pieces x = x0 + x1 + (rest), and the three largest cross products are contracted in ONE
// K-concatenated GEMM on tcgen05.mma kind::f16 (BF16 inputs, FP32 accumulators in TMEM):
//      A row (query) = [ -2 q0 | -2 q0 | -2 q1 | 1 1 1 0.. ]      (3R + 3 columns, padded to a multiple of 16)
//      B row (key)   = [   k0  |   k1  |   k0  | n0 n1 n2 0.. ]   with |k|^2 = n0 + n1 + n2 exactly
//   The dropped products (q0.k2, q1.k1, q2.k0) are bounded by 1.5 * 2^-16 |q||k|; the certificate below uses
//   eps = 2^-15 (|q| + |k|max)^2, which leaves more than 5x room for the accumulation error of the tensor core.
//   R = 20: 64 columns = 4 MMAs of K = 16 per 128x256 tile (the TF32x3 kernel this replaces needed 9 at half the rate).
//   The MMA shape is M = 128, N = 256: measured (tools/mma_bench.cu) one cta_group::1 tcgen05.mma costs 93 cycles at any
//   N <= 128 but 128 cycles at N = 256, i.e. only N = 256 runs the tensor pipe at its floor (M*N/256 cycles per K = 16).
//
// Key image (key_image_kernel). The B operand is computed ONCE per inserted key and kept in HBM in exactly the
// shared-memory layout tcgen05 reads (K-major, no swizzle, 8x16-byte core matrices): one 256-key tile is one
// contiguous block (32 KB at R = 20), so a tile is moved by a single cp.async.bulk (TMA) and no thread of the
// query kernel ever touches a key.
//
// Query kernel (knn_tc_kernel). One CTA per SM owns 256 queries (two M = 128 accumulator sets) and one contiguous
// range of key tiles, so every key tile fetched from L2 is used for 256 queries. Warp roles:
//     warp 0      MMA issuer : one thread; per key tile 2 x KSTEPS tcgen05.mma, committed per query tile
//     warp 1      TMA issuer : one thread; ring of NSTAGE key tiles in shared memory
//     warps 2-3   threshold service (below)
//     warps 4-11  epilogue   : thread = query = TMEM lane. tcgen05.ld 32 columns at a time, double buffered
//                              (the next load is in flight while the current 32 scores are examined); the common
//                              case is a FMNMX3 min-tree and one compare against the query's threshold
//   TMEM: 2 query tiles x 256 FP32 columns = all 512 columns. The two query tiles are each other's double buffer:
//   while the four epilogue warps of one drain its accumulator, the tensor pipe fills the other's.
//
// Thresholds. A thread drops every 8-key group whose best score is at or above its threshold and appends the others
// (first key, group minimum: 8 bytes) to the hit queue of its (query, range) in global memory. The threshold is a union
// bound: every thread publishes the best score of its range; for a query, the K'-th smallest of the published range
// minima is backed by K' distinct keys, so it bounds the global K'-th best score from above. The service warps recompute
// that bound continuously for the CTA's share of the queries; epilogue threads read it once per tile. No bootstrap or
// sample pass is needed.
//
// Re-rank + certificate (knn_rerank_kernel). One CTA per query: the keys of every queued group whose minimum is at or
// below the query's cut (its final union bound) are re-scored with the reference's exact float order (k3_knn.cu),
// the top-K by (d2, id) is selected, and the result is CERTIFIED: every key the prefilter dropped has score >= cut,
// i.e. exact d2 > cut + |q|^2 - eps; if the K-th selected distance is below that, no dropped key can belong to (or
// tie with) the top-K. Queries that fail are appended to a list and redone by the exact kernel.
//
// Roofline: 2*R*Q*N algorithmic flops against the tensor pipe (the kernel issues 3.2x that in BF16), 4*R*N
// algorithmic bytes against HBM (the image is 128 B/key at R = 20, read once per 256 queries from L2).
#include "common.cuh"
#include "kernels.h"

#include <cuda_bf16.h>

#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace {

constexpr int kKPrime = 16;        /* slots per query (stride). K', the number of distinct keys that back a query's threshold, is chosen per
                                    * launch: K <= K' - 2. The bound is the LARGEST of K' slot minima, which about K' H(K') keys undercut
                                    * (37 at K' = 12, 54 at 16): smaller is tighter, so K <= 10 (the reference's default) runs with K' = 12. */
__host__ __device__ constexpr int kprime_for(int K) { return K <= 10 ? 12 : 16; }
constexpr int kQueueCap = 64;                /* hit queue per (query, range): entries of 8 words (first key of a 32-key chunk, the best scores of its four 8-key groups, 3 pad); ~10 are used; a tile adds at most 8 */
constexpr int kEpiThreads = 256;   /* 8 epilogue warps: query tile = (warp-4)/4, TMEM lane quadrant = warp%4 */
constexpr int kThreads = 384;
constexpr int kNT = 256;           /* keys per tile: one TMA copy, one N = 256 accumulator per query tile */
constexpr int kQPerCta = 256;
constexpr int kSlotStride = 20;    /* ints per query in the slot array: K' (<= 16) range minima, then the query's DIRECT bound (below), 3 spare */
constexpr int kDirect = 16;        /* index of the direct bound within a query's slots */
constexpr int kNoThr = 0x7f7f7f7f; /* memset pattern of the slots: 3.39e38 = "nothing yet" */
constexpr float kThrInit = 1.0e38f;

// order-preserving float <-> signed int image
__device__ __forceinline__ int ordered_int(float f) { const int b = __float_as_int(f); return b ^ ((b >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// ---- tcgen05 / mbarrier PTX wrappers ---------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(scl_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(scl_smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(scl_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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
// 32 consecutive fp32 columns of this warp's 32 TMEM lanes; asynchronous until tmem_wait32 on the same registers
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// wait for the outstanding tcgen05.ld; the registers are in/out operands so no use can be scheduled above the wait
__device__ __forceinline__ void tmem_wait32(uint32_t (&r)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
          "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]),
          "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
          "+r"(r[30]), "+r"(r[31])
        :: "memory");
}
__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   /* FMNMX3 */
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <int R> struct TcCfg {
    /* Keys longer than 40 values (the 80-row keys of the Lidar-Iris family, descriptor.h:1087-1250) are handled as SEGS segments of
     * RS values: the score is the sum of the segments' partial scores, i.e. ONE accumulation over SEGS key sub-tiles, which
     * stream through the same 32 KB stages (the whole 256-key x 256-column operand would not fit beside the queries). */
    static constexpr int SEGS = R > 40 ? R / 20 : 1;
    static constexpr int RS = R / SEGS;                       /* key values per segment */
    static constexpr int KTOT = (3 * RS + 3 + 15) / 16 * 16;  /* GEMM K per segment: 64 at RS = 20, 128 at RS = 40 */
    static constexpr int CHUNKS = KTOT / 8;                   /* 16-byte K chunks per row */
    static constexpr int KSTEPS = KTOT / 16;                  /* tcgen05.mma instructions per 128x256 tile and segment */
    static constexpr uint32_t SBO = 128;                      /* 8-row core-matrix groups follow each other */
    static constexpr uint32_t LBO_A = 128 * 16;               /* distance between 16-byte K chunks: rows x 16 B */
    static constexpr uint32_t LBO_B = kNT * 16;
    static constexpr uint32_t TILE_A = 128 * KTOT * 2;        /* one query tile, one segment: 16 KB / 32 KB */
    static constexpr uint32_t TILE_B = kNT * KTOT * 2;        /* one key tile, one segment: 32 KB / 64 KB */
    static constexpr uint32_t IMG_TILE = SEGS * TILE_B;       /* one key tile of the image: its segments one after the other */
    static constexpr int NSTAGE = R <= 20 ? 5 : (SEGS > 1 ? 3 : 2);   /* most key (sub-)tiles in flight in shared memory (the launch may ask for fewer) */
    static constexpr uint32_t OFF_BAR = 0;                    /* mbarriers, tmem slot, flags */
    static constexpr uint32_t OFF_THR = 1024;                 /* [256] union bounds */
    static constexpr uint32_t OFF_A = 2048;                   /* two query tiles x SEGS segments */
    static constexpr uint32_t OFF_B = OFF_A + 2 * SEGS * TILE_A;
    static constexpr uint32_t TOTAL = OFF_B + NSTAGE * TILE_B;
    static constexpr uint32_t total(int stages) { return OFF_B + (uint32_t)stages * TILE_B; }
    /* D = F32, A = B = BF16, both K-major, N = kNT, M = 128 */
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kNT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};

// value of GEMM column `idx` of the key row (B) / of the query row (A)
template <int R>
__device__ __forceinline__ float b_column(const float (&k0)[R], const float (&k1)[R], const float (&nn)[3], int idx)
{
    if (idx < R) return k0[idx];
    if (idx < 2 * R) return k1[idx - R];
    if (idx < 3 * R) return k0[idx - 2 * R];
    if (idx < 3 * R + 3) return nn[idx - 3 * R];
    return 0.0f;
}

// ---- key image: keys [n][R] fp32 -> B-operand tiles ------------------------------------------------
template <int R>
__global__ void __launch_bounds__(128) key_image_kernel(const float* __restrict__ keys, const float* __restrict__ knorm, int k_lo, int k_hi,
                                                        unsigned char* __restrict__ img)
{
    using C = TcCfg<R>;
    constexpr int RS = C::RS;
    const int key = k_lo + blockIdx.x * 128 + threadIdx.x;
    if (key >= k_hi) return;
    const float n = __ldg(knorm + key);
    const int row = key % kNT;
#pragma unroll 1
    for (int seg = 0; seg < C::SEGS; seg++) {
        float k0[RS], k1[RS], nn[3];
#pragma unroll
        for (int g = 0; g < RS / 4; g++) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(keys + (size_t)key * R + seg * RS) + g);
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                k0[4 * g + i] = bf16_round(xs[i]);
                k1[4 * g + i] = bf16_round(xs[i] - k0[4 * g + i]);
            }
        }
        /* the squared norm of the whole key travels with the first segment */
        nn[0] = seg == 0 ? bf16_round(n) : 0.0f; nn[1] = seg == 0 ? bf16_round(n - nn[0]) : 0.0f; nn[2] = seg == 0 ? bf16_round(n - nn[0] - nn[1]) : 0.0f;
        unsigned char* dst = img + (size_t)(key / kNT) * C::IMG_TILE + (uint32_t)seg * C::TILE_B + (uint32_t)(row >> 3) * C::SBO + (uint32_t)(row & 7) * 16;
#pragma unroll
        for (int c = 0; c < C::CHUNKS; c++) {
            uint4 v;
            v.x = pack_bf16x2(b_column<RS>(k0, k1, nn, 8 * c + 0), b_column<RS>(k0, k1, nn, 8 * c + 1));
            v.y = pack_bf16x2(b_column<RS>(k0, k1, nn, 8 * c + 2), b_column<RS>(k0, k1, nn, 8 * c + 3));
            v.z = pack_bf16x2(b_column<RS>(k0, k1, nn, 8 * c + 4), b_column<RS>(k0, k1, nn, 8 * c + 5));
            v.w = pack_bf16x2(b_column<RS>(k0, k1, nn, 8 * c + 6), b_column<RS>(k0, k1, nn, 8 * c + 7));
            *reinterpret_cast<uint4*>(dst + (uint32_t)c * C::LBO_B) = v;
        }
    }
}

// A nearly full hit queue (rare on scattered databases; the rule on a trajectory, where every good key comes with a run of
// near-duplicate neighbours in the same tile) is re-filtered. First a DIRECT bound is taken from the queue itself: its groups
// are disjoint sets of keys, so the K-th smallest group minimum in it is undercut (or met) by K distinct keys of this range
// alone and bounds the query's K-th best score from above — tight exactly when the best keys sit together, where the union
// bound over range minima is loose. It is published for the whole query (atomicMin on the query's direct-bound word: the
// service warps fold it into every range's threshold and the re-rank into its cut) and applied here; then chunks queued
// under an earlier, looser threshold whose best score is not below the new one are dropped like any other key.
// The whole WARP serves the lanes that need it, one queue at a time (lane <-> two entries of the queue): a thread compacting
// its own queue alone ran ~20 000 cycles with 31 lanes idle, and on a smooth trajectory every lane needs it about once
// (measured: 480 us instead of 130 us for the kernel).
// Out of line and by value (uint2 = new entry count, new threshold bits): the hot loop keeps its register allocation, and no
// hot variable has its address taken.
__device__ __noinline__ uint2 compact_queues_warp(bool need, uint4* my_q, int n_hit, float thr, int K, int* my_direct_word, int lane)
{
    const float inf = __int_as_float(0x7f800000);
    const int inf_img = ordered_int(inf);
    unsigned todo = __ballot_sync(0xffffffffu, need);
    while (todo) {
        const int L = __ffs(todo) - 1;
        todo &= todo - 1;
        const unsigned long long qp = (unsigned long long)my_q, dp = (unsigned long long)my_direct_word;
        uint4* q = reinterpret_cast<uint4*>(((unsigned long long)__shfl_sync(0xffffffffu, (unsigned)(qp >> 32), L) << 32) | __shfl_sync(0xffffffffu, (unsigned)qp, L));
        int* direct_word = reinterpret_cast<int*>(((unsigned long long)__shfl_sync(0xffffffffu, (unsigned)(dp >> 32), L) << 32) | __shfl_sync(0xffffffffu, (unsigned)dp, L));
        const int n = __shfl_sync(0xffffffffu, n_hit, L);
        float t = __shfl_sync(0xffffffffu, thr, L);
        uint4 ea[2]; uint32_t eb[2]; bool valid[2]; float v[8];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int e = lane + 32 * u;
            valid[u] = e < n;
            ea[u] = make_uint4(0u, 0x7f800000u, 0x7f800000u, 0x7f800000u); eb[u] = 0x7f800000u;
            if (valid[u]) { ea[u] = __ldcg(q + 2 * e); eb[u] = __ldcg(reinterpret_cast<const uint32_t*>(q + 2 * e + 1)); }
            v[4 * u] = __uint_as_float(ea[u].y); v[4 * u + 1] = __uint_as_float(ea[u].z); v[4 * u + 2] = __uint_as_float(ea[u].w); v[4 * u + 3] = __uint_as_float(eb[u]);
        }
        float cmin[2];
#pragma unroll
        for (int u = 0; u < 2; u++) cmin[u] = fminf(fminf(v[4 * u], v[4 * u + 1]), fminf(v[4 * u + 2], v[4 * u + 3]));
        /* the K-th smallest of the queue's group minima (equal minima are distinct keys and count separately) */
        float kth = inf; int have = 0;
        for (int r = 0; r < K; r++) {
            float loc = v[0];
#pragma unroll
            for (int j = 1; j < 8; j++) loc = fminf(loc, v[j]);
            const int li = ordered_int(loc);
            const int w = __reduce_min_sync(0xffffffffu, li);
            if (w >= inf_img) break;
            const unsigned holders = __ballot_sync(0xffffffffu, li == w);
            if (lane == __ffs(holders) - 1) {
                bool done = false;
#pragma unroll
                for (int j = 0; j < 8; j++) if (!done && ordered_int(v[j]) == w) { v[j] = inf; done = true; }
            }
            kth = ordered_float(w); have++;
        }
        if (have == K) {
            /* chunks holding a group at or below kth must stay: the threshold is the next float above it */
            const float up = kth == 0.0f ? 1.0e-45f : __int_as_float(__float_as_int(kth) + (kth >= 0.0f ? 1 : -1));
            if (up < t) { t = up; if (lane == 0) atomicMin(direct_word, ordered_int(up)); }
        }
        int n_new = 0;
        __syncwarp();                                    /* every entry is in registers before any is rewritten */
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const bool keep = valid[u] && cmin[u] < t;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            const int pos = n_new + __popc(m & ((1u << lane) - 1u));
            if (keep) { __stcg(q + 2 * pos, ea[u]); __stcg(q + 2 * pos + 1, make_uint4(eb[u], 0u, 0u, 0u)); }
            n_new += __popc(m);
        }
        if (lane == L) { n_hit = n_new; thr = t; }
        __syncwarp();
    }
    return make_uint2((unsigned)n_hit, __float_as_uint(thr));
}

// ---- the query kernel ---------------------------------------------------------------------------------
template <int R, bool TIMES>
__global__ void __launch_bounds__(kThreads, 1) knn_tc_kernel(
    const float* __restrict__ qkeys, int Q, const unsigned char* __restrict__ img, int key_hi, int n_ranges,
    long long* __restrict__ times /* null, or [grid][16] developer counters (SCL_TC_TIMES=1) */,
    int* __restrict__ slots /* [Q][K'] range minima by range % K' (ordered-int image) */,
    uint4* __restrict__ hq /* [Q][n_ranges][kQueueCap][2] hit queues: (first key of the chunk, best scores of its four groups) */, int* __restrict__ hq_cnt /* [Q][n_ranges] */,
    int* __restrict__ dbg /* null, or developer counters */, int dev_flags /* SCL_TC_FLAGS: timing experiments, results are then wrong */,
    int kp /* K' of this launch: 12 or 16 */, int K /* neighbours asked for */, int NS /* key tiles in flight: 2 .. C::NSTAGE (fewer leave shared memory to kernels of other lanes) */)
{
    using C = TcCfg<R>;
    constexpr int NSMAX = C::NSTAGE;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t *full = bars, *empty = bars + NSMAX, *tfull = bars + 2 * NSMAX, *tempty = bars + 2 * NSMAX + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSMAX + 4);
    volatile int* epi_done = reinterpret_cast<volatile int*>(tmem_slot + 1);
    volatile int* sthr = reinterpret_cast<volatile int*>(smem + C::OFF_THR);          /* [256] union bounds of the CTA's queries */
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_groups = gridDim.x / n_ranges;
    const int group = blockIdx.x % n_groups, range = blockIdx.x / n_groups;   /* neighbouring CTAs share a key range (L2 reuse) */
    const int q_base = group * kQPerCta;
    /* Key tiles are dealt round-robin: this CTA owns tiles range, range + n_ranges, ... Every CTA then sees a sample of the
     * WHOLE database, so the range minima that make up the union bound are alike even when the database is ordered
     * (a trajectory: neighbouring keys are neighbouring places, and whole stretches of it are far from the query). */
    const int n_tiles_all = (key_hi + kNT - 1) / kNT;
    const int n_tiles = range < n_tiles_all ? (n_tiles_all - range + n_ranges - 1) / n_ranges : 0;
    const int n_service = min(n_ranges, n_tiles_all);                         /* ranges that hold keys */
    /* ... and in a scattered order: the it-th tile a CTA visits is its (it * stride mod n_tiles)-th, stride about 0.618 n_tiles and
     * coprime to it (a low-discrepancy walk). On a smooth trajectory a sweep in key order approaches the query's place
     * monotonically: every new tile undercuts the bound so far, every chunk of it is queued and the queues churn; visited in
     * scattered order the first few tiles already sample the whole trajectory and the bound settles early. */
    int tile_stride = 1;
    if (n_tiles > 2) {
        tile_stride = (int)(0.6180339887f * (float)n_tiles);
        if (tile_stride < 1) tile_stride = 1;
        while (true) {
            int a = tile_stride, b = n_tiles;
            while (b) { const int t = a % b; a = b; b = t; }
            if (a == 1) break;
            tile_stride++;
        }
    }
    // ---- one-time setup -----------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { scl_mbar_init(&full[s], 1); scl_mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; s++) { scl_mbar_init(&tfull[s], 1); scl_mbar_init(&tempty[s], 4); }
        *epi_done = 0;
        scl_mbar_fence_init();
    }
    if (threadIdx.x < kQPerCta) sthr[threadIdx.x] = kNoThr;
    if (warp == 0) {   /* TMEM: 2 query tiles x 256 fp32 columns */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(scl_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x < kQPerCta) {
        /* A operand: rows = the CTA's 256 queries, columns [-2 q0 | -2 q0 | -2 q1 | 1 1 1 0..] per segment. One thread per row: its key
         * segment is fetched with RS/4 independent 16-byte loads (one L2 round trip), split, and written as CHUNKS 16-byte stores. */
        constexpr int RS = C::RS;
        const int m = threadIdx.x, qi = q_base + m;
        const int r = m & 127;
#pragma unroll 1
        for (int seg = 0; seg < C::SEGS; seg++) {
            float a0[RS], a1[RS];                           /* -2 q0, -2 q1 */
#pragma unroll
            for (int g = 0; g < RS / 4; g++) {
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qi < Q) x = __ldg(reinterpret_cast<const float4*>(qkeys + (size_t)qi * R + seg * RS) + g);
                const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float q0 = bf16_round(xs[i]);
                    a0[4 * g + i] = -2.0f * q0;
                    a1[4 * g + i] = -2.0f * bf16_round(xs[i] - q0);
                }
            }
            const float one = qi < Q ? 1.0f : 0.0f;
            unsigned char* dst = smem + C::OFF_A + (uint32_t)((m >> 7) * C::SEGS + seg) * C::TILE_A + (uint32_t)(r >> 3) * C::SBO + (uint32_t)(r & 7) * 16;
#pragma unroll
            for (int c = 0; c < C::CHUNKS; c++) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int idx = 8 * c + j;              /* compile-time after unrolling */
                    v[j] = idx < RS ? a0[idx < RS ? idx : 0] : idx < 2 * RS ? a0[idx < 2 * RS && idx >= RS ? idx - RS : 0]
                         : idx < 3 * RS ? a1[idx < 3 * RS && idx >= 2 * RS ? idx - 2 * RS : 0] : idx < 3 * RS + 3 ? one : 0.0f;
                }
                *reinterpret_cast<uint4*>(dst + (uint32_t)c * C::LBO_A) =
                    make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            }
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 4) {
        // ===== epilogue: thread = query = TMEM lane ====================================================
        const int qt = (warp - 4) >> 2;                 /* query tile of this warp */
        const int row = (warp & 3) * 32 + lane;         /* TMEM lane */
        const int qi = q_base + qt * 128 + row;
        const bool live = qi < Q;
        int n_hit = 0;                                  /* 8-key groups queued by this thread */
        int n_slow = 0;                                 /* developer counter (SCL_TC_TIMES) */
        float thr = (live && !(dev_flags & 1)) ? kThrInit : -kThrInit;        /* rows beyond Q never queue anything */
        float published = kThrInit;
        int* my_slot = slots + (size_t)(live ? qi : 0) * kSlotStride + (range % kp);
        volatile int* my_sthr = sthr + qt * 128 + row;
        uint4* my_q = hq + ((size_t)(live ? qi : 0) * n_ranges + range) * (size_t)(2 * kQueueCap);
        // The four epilogue warps of a query tile are coupled through the accumulator hand-off (it is refilled only when all
        // four have drained it), so the per-chunk code sits on the critical path of the whole CTA. It is BRANCH-FREE on purpose:
        // ptxas only hoists a tcgen05.ld above the min-tree of the previous chunk when both are in one basic block (with a
        // branch per chunk it sank every load to the end of its block, right in front of its consumers: 220 cycles per chunk
        // instead of 60). 18 FMNMX(3) per 32 scores; a hit APPENDS (first key, the four group minima) of the chunk to the
        // (query, range) queue in global memory with predicated stores; the re-rank kernel re-scores the keys of the groups at
        // or below the cut exactly. The queue has room for four tiles' worth of chunks and is compacted between tiles when nearly full.
        float tile_min = kThrInit;
        bool overflowed = false;
        auto examine = [&](const uint32_t (&r)[32], int key_first) {
            float g[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float x0 = __uint_as_float(r[8 * j]), x1 = __uint_as_float(r[8 * j + 1]), x2 = __uint_as_float(r[8 * j + 2]),
                            x3 = __uint_as_float(r[8 * j + 3]), x4 = __uint_as_float(r[8 * j + 4]), x5 = __uint_as_float(r[8 * j + 5]),
                            x6 = __uint_as_float(r[8 * j + 6]), x7 = __uint_as_float(r[8 * j + 7]);
                g[j] = fminf(fmin3(fmin3(x0, x1, x2), fmin3(x3, x4, x5), x6), x7);
            }
            const float m = fminf(fmin3(g[0], g[1], g[2]), g[3]);
            if (TIMES && __any_sync(0xffffffffu, m < thr)) n_slow++;
            /* one compare and two predicated 16-byte stores per chunk: the four group minima travel together and the re-rank
             * decides per group (per-group compares and stores here cost a quarter of the chunk's issue slots) */
            const bool hit = m < thr;
            if (hit) {
                __stcg(my_q + 2 * n_hit, make_uint4((uint32_t)key_first, __float_as_uint(g[0]), __float_as_uint(g[1]), __float_as_uint(g[2])));
                __stcg(my_q + 2 * n_hit + 1, make_uint4(__float_as_uint(g[3]), 0u, 0u, 0u));
            }
            n_hit += hit ? 1 : 0;
            tile_min = fminf(tile_min, m);
        };
        long long tw = 0, c0 = 0, t_first = 0, t_body = 0, t_tail = 0, q0 = 0;
        if (TIMES) c0 = clock64();
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(qt * kNT);
        uint32_t va[32], vb[32];
        float next_thr = thr;                               /* the service warps' bound, read one tile ahead of its use */
        if (n_tiles > 0) {
            /* Start-up: with no threshold yet, every score of the first tile would be a hit. Instead the first tile is
             * read twice: a first pass only finds the range minimum so far and publishes it; as soon as K' ranges have done
             * so the service warps deliver a union bound (a few microseconds), and the normal pass starts with it. */
            scl_mbar_wait(&tfull[qt], 0);
            tc_fence_after();
            if (n_service >= kp && (range + 1) * kNT <= key_hi && !(dev_flags & 2)) {
                float m0 = kThrInit;
#pragma unroll 1
                for (int c = 0; c < kNT / 32; c++) {
                    tmem_ld32_issue(lane_base + (uint32_t)(c * 32), va);
                    tmem_wait32(va);
#pragma unroll
                    for (int j = 0; j < 32; j += 2) m0 = fmin3(m0, __uint_as_float(va[j]), __uint_as_float(va[j + 1]));
                }
                if (live) { published = m0; atomicMin(my_slot, ordered_int(m0)); }
                const long long w0 = clock64();
                while (true) {
                    const bool ok = !live || *my_sthr < 0x7f000000;
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (clock64() - w0 > 40000) { if (dbg && lane == 0) atomicAdd(dbg + 5, 1); break; }
                    __nanosleep(100);
                }
            }
            if (live) next_thr = fminf(next_thr, ordered_float(*my_sthr));
        }
        // One accumulator (kNT keys of this query tile) per iteration, as eight 32-column chunks alternating between va and vb:
        // the load of the next chunk is in flight while the current one is examined, and the accumulator is handed back to the
        // MMA issuer as soon as its last chunk is in registers.
#pragma unroll 1
        int tile_pos = 0;                                   /* it * tile_stride mod n_tiles */
        for (int it = 0; it < n_tiles; it++) {
            const int key0 = (range + tile_pos * n_ranges) * kNT;
            tile_pos += tile_stride; if (tile_pos >= n_tiles) tile_pos -= n_tiles;
            if (it > 0) {
                long long p0 = 0;
                if (TIMES) p0 = clock64();
                scl_mbar_wait(&tfull[qt], (uint32_t)(it & 1));
                if (TIMES) tw += clock64() - p0;
                tc_fence_after();
            }
            thr = fminf(thr, next_thr);
            tile_min = kThrInit;
            if (TIMES) q0 = clock64();
            tmem_ld32_issue(lane_base, va);
#pragma unroll
            for (int c = 0; c < kNT / 32; c += 2) {
                tmem_wait32(va);
                if (TIMES && c == 0) { const long long q1 = clock64(); t_first += q1 - q0; q0 = q1; }
                tmem_ld32_issue(lane_base + (uint32_t)(32 * c + 32), vb);
                examine(va, key0 + 32 * c);
                tmem_wait32(vb);
                if (c + 2 < kNT / 32) tmem_ld32_issue(lane_base + (uint32_t)(32 * c + 64), va);
                else {
                    /* every score of this accumulator is in registers: hand it back to the MMA issuer */
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[qt]);
                }
                examine(vb, key0 + 32 * c + 32);
            }
            if (TIMES) { const long long q1 = clock64(); t_body += q1 - q0; q0 = q1; }
            /* a new range minimum feeds the union bound (tiles wholly below key_hi only; thr is -inf for rows beyond Q) */
            if (key0 + kNT <= key_hi && tile_min < fminf(thr, published)) { published = tile_min; atomicMin(my_slot, ordered_int(tile_min)); }
            {
                const bool need = n_hit > kQueueCap - kNT / 32;     /* no room for another tile's worth of chunks: compact */
                if (__any_sync(0xffffffffu, need)) {
                    const uint2 cq = compact_queues_warp(need, my_q, n_hit, thr, K, slots + (size_t)(live ? qi : 0) * kSlotStride + kDirect, lane);
                    n_hit = (int)cq.x; thr = __uint_as_float(cq.y);
                    if (need && n_hit > kQueueCap - kNT / 32) { overflowed = true; n_hit = 0; thr = -kThrInit; }   /* sticky: nothing more is queued, the query is redone exactly */
                }
            }
            if (live) next_thr = ordered_float(*my_sthr);
            if (TIMES) t_tail += clock64() - q0;
        }
        if (overflowed) n_hit = kQueueCap + 1;
        if (TIMES) {
            const int wp = __reduce_add_sync(0xffffffffu, min(n_hit, kQueueCap)), wmax = __reduce_max_sync(0xffffffffu, n_hit);
            if (lane == 0) {
                long long* o = times + (size_t)blockIdx.x * 16;
                atomicAdd(reinterpret_cast<unsigned long long*>(o + 0), (unsigned long long)tw);
                atomicAdd(reinterpret_cast<unsigned long long*>(o + 1), (unsigned long long)(clock64() - c0));
                atomicAdd(reinterpret_cast<unsigned long long*>(o + 2), (unsigned long long)n_slow);
                atomicAdd(reinterpret_cast<unsigned long long*>(o + 3), (unsigned long long)wp);
                atomicMax(reinterpret_cast<unsigned long long*>(o + 4), (unsigned long long)wmax);
                atomicAdd(reinterpret_cast<unsigned long long*>(o + 9), (unsigned long long)t_first);
                atomicAdd(reinterpret_cast<unsigned long long*>(o + 10), (unsigned long long)t_body);
                atomicAdd(reinterpret_cast<unsigned long long*>(o + 11), (unsigned long long)t_tail);
            }
        }
        /* Everything this thread did NOT queue scored >= the threshold in force at the time >= the maximum the query's slots
         * end at (slots only fall, and a thread only ever applies a bound computed from them): the re-rank takes that maximum
         * as the query's cut. The count is written even when it is zero, so the queues need no clearing between batches. */
        if (live) hq_cnt[(size_t)qi * n_ranges + range] = n_hit;
        __syncwarp();
        if (lane == 0) atomicAdd(const_cast<int*>(epi_done), 1);
    } else if (warp >= 2) {
        // ===== threshold service: union bound of the CTA's 256 queries ==================================
        // Range r publishes its best score so far into slot r % K' of the query (atomicMin). The K' slots then hold the
        // scores of K' DISTINCT keys (different ranges), so their maximum bounds the query's K'-th best score from above.
        // Each lane refreshes four queries: 4 K' bytes from L2 and K' - 1 max operations per query, a microsecond per sweep.
        int sweeps = 0;
        while (n_tiles > 0 && !(dev_flags & 2)) {
            const bool last = *epi_done >= 8;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int j = (warp - 2) * 128 + 32 * k + lane;
                const int qi = q_base + j;
                if (qi < Q) {
                    const int4* p = reinterpret_cast<const int4*>(slots + (size_t)qi * kSlotStride);
                    int4 v[kKPrime / 4];
#pragma unroll
                    for (int i = 0; i < kKPrime / 4; i++) v[i] = __ldcg(p + i);
                    const int direct = __ldcg(slots + (size_t)qi * kSlotStride + kDirect);
                    int m = (int)0x80000000;
#pragma unroll
                    for (int i = 0; i < kKPrime / 4; i++) if (4 * i < kp) m = max(m, max(max(v[i].x, v[i].y), max(v[i].z, v[i].w)));
                    m = min(m, direct);                  /* the union bound (valid once all K' slots are filled) or a range's direct bound */
                    if (m < 0x7f000000) sthr[j] = m;
                }
            }
            sweeps++;
            if (last) break;
            if (sweeps > 64) __nanosleep(1000);          /* the bound moves fast at the start: sweep back to back there */
        }
        if (TIMES && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(times + (size_t)blockIdx.x * 16 + 5), (unsigned long long)sweeps);
    } else if (warp == 1) {
        // ===== TMA issuer: one thread, one bulk copy per key tile =======================================
        if (lane == 0) {
            /* One copy per key tile, or, for segmented keys, per (key tile, segment): a sub-tile serves both query tiles. */
            const unsigned char* src = img + (size_t)range * C::IMG_TILE;
            const size_t step = (size_t)n_ranges * C::IMG_TILE;
            constexpr int kPerTile = C::SEGS;
            int ld = 0, tile = 0;
            for (int it = 0; it < n_tiles; it++) {
#pragma unroll 1
                for (int j = 0; j < kPerTile; j++, ld++) {
                    const int b = ld % NS; const uint32_t ph = (ld / NS) & 1;
                    scl_mbar_wait(&empty[b], ph ^ 1u);
                    scl_mbar_expect_tx(&full[b], C::TILE_B);
                    scl_bulk_g2s(smem + C::OFF_B + (uint32_t)b * C::TILE_B, src + (size_t)tile * step + (size_t)(j % C::SEGS) * C::TILE_B, C::TILE_B, &full[b]);
                }
                tile += tile_stride; if (tile >= n_tiles) tile -= n_tiles;
            }
        }
        __syncwarp();
    } else {
        // ===== MMA issuer: one thread ==============================================================
        if (lane == 0) {
            const uint32_t a_base = scl_smem_u32(smem + C::OFF_A), b_base = scl_smem_u32(smem + C::OFF_B);
            long long t_te = 0, t_fu = 0, c0 = 0;
            int ld = 0;                                             /* key (sub-)tiles consumed so far */
            for (int it = 0; it < n_tiles; it++) {
                if (C::SEGS == 1) {
                    const int b = ld % NS; const uint32_t bph = (ld / NS) & 1;
                    ld++;
                    if (TIMES) c0 = clock64();
                    scl_mbar_wait(&full[b], bph);                       /* key tile landed */
                    if (TIMES) { const long long c1 = clock64(); t_fu += c1 - c0; c0 = c1; }
                    const uint32_t bs = b_base + (uint32_t)b * C::TILE_B;
#pragma unroll
                    for (int qt = 0; qt < 2; qt++) {
                        /* accumulator drained by its four epilogue warps? This one thread SPINS (a sleeping try_wait wakes up late:
                         * measured 5 us per batch); a bound turns a protocol bug into a trap instead of a hung GPU */
                        if (!mbar_test(&tempty[qt], (uint32_t)((it & 1) ^ 1))) {
                            const long long w0 = clock64();
                            while (!mbar_test(&tempty[qt], (uint32_t)((it & 1) ^ 1))) { if (clock64() - w0 > (4ll << 30)) __trap(); }
                        }
                        tc_fence_after();
                        const uint32_t d = tmem_base + (uint32_t)(qt * kNT);
                        const uint32_t as = a_base + (uint32_t)qt * C::TILE_A;
#pragma unroll
                        for (int k = 0; k < C::KSTEPS; k++)
                            if (!(dev_flags & 4)) tc_mma_bf16(d, make_desc(as + 2 * k * C::LBO_A, C::LBO_A, C::SBO), make_desc(bs + 2 * k * C::LBO_B, C::LBO_B, C::SBO), C::IDESC, k > 0 ? 1u : 0u);
                        tc_commit(&tfull[qt]);                          /* accumulator ready for the epilogue */
                    }
                    tc_commit(&empty[b]);                               /* key tile reusable once these MMAs retire */
                } else {
                    /* segmented keys: the SEGS sub-tiles of the key tile accumulate into one score per query tile; every sub-tile
                     * is used for query tile 0, then 1 (one trip from L2 for both). An accumulator is handed to its epilogue after
                     * its last segment, so query tile 0 drains under the last segment of query tile 1. */
#pragma unroll 1
                    for (int seg = 0; seg < C::SEGS; seg++, ld++) {
                        const int b = ld % NS; const uint32_t bph = (ld / NS) & 1;
                        scl_mbar_wait(&full[b], bph);
                        tc_fence_after();
                        const uint32_t bs = b_base + (uint32_t)b * C::TILE_B;
#pragma unroll 1
                        for (int qt = 0; qt < 2; qt++) {
                            if (seg == 0) {
                                if (!mbar_test(&tempty[qt], (uint32_t)((it & 1) ^ 1))) {
                                    const long long w0 = clock64();
                                    while (!mbar_test(&tempty[qt], (uint32_t)((it & 1) ^ 1))) { if (clock64() - w0 > (4ll << 30)) __trap(); }
                                }
                                tc_fence_after();
                            }
                            const uint32_t d = tmem_base + (uint32_t)(qt * kNT);
                            const uint32_t as = a_base + (uint32_t)(qt * C::SEGS + seg) * C::TILE_A;
#pragma unroll
                            for (int k = 0; k < C::KSTEPS; k++)
                                tc_mma_bf16(d, make_desc(as + 2 * k * C::LBO_A, C::LBO_A, C::SBO), make_desc(bs + 2 * k * C::LBO_B, C::LBO_B, C::SBO), C::IDESC, (seg > 0 || k > 0) ? 1u : 0u);
                            if (seg == C::SEGS - 1) tc_commit(&tfull[qt]);
                        }
                        tc_commit(&empty[b]);                           /* sub-tile reusable once these MMAs retire */
                    }
                }
                if (TIMES) { const long long c1 = clock64(); t_te += c1 - c0; }
            }
            if (TIMES) { times[(size_t)blockIdx.x * 16 + 6] = t_fu; times[(size_t)blockIdx.x * 16 + 7] = t_te; times[(size_t)blockIdx.x * 16 + 8] = n_tiles; }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// the reference's float distance (the operation order of k3_knn.cu: nanoflann adds four squares at a time, libnabo one)
// on a key row held in registers
template <int METRIC, int R>
__device__ __forceinline__ float exact_d2(const float* __restrict__ q, const float4 (&kv)[R / 4])
{
    float result = 0.0f;
#pragma unroll
    for (int g = 0; g < R / 4; g++) {
        const float d0 = __fsub_rn(q[4 * g], kv[g].x), d1 = __fsub_rn(q[4 * g + 1], kv[g].y), d2 = __fsub_rn(q[4 * g + 2], kv[g].z), d3 = __fsub_rn(q[4 * g + 3], kv[g].w);
        if (METRIC == 0) {
            const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
            result = __fadd_rn(result, s);
        } else {
            result = __fadd_rn(result, __fmul_rn(d0, d0)); result = __fadd_rn(result, __fmul_rn(d1, d1));
            result = __fadd_rn(result, __fmul_rn(d2, d2)); result = __fadd_rn(result, __fmul_rn(d3, d3));
        }
    }
    return result;
}

// Phase B: exact re-rank + certificate. One WARP per query (four queries per CTA, no block-wide barriers): the kernel is
// a chain of dependent memory round trips (counts -> queue entries -> key rows), so what matters is how many loads a lane
// has in flight per trip, not how many threads share a query. The hit queues of all ranges are flattened (counts ->
// prefix sums in the warp's shared memory), every lane reads a few independent entries per trip, the 8 keys of every
// surviving group (best score at or below the cut) are re-scored exactly, four rows per lane in flight, and the warp
// selects the top-K and certifies it.
constexpr int kMaxGroups = 512;                         /* surviving groups per query the list holds (about 3 K' are expected: the cut is the LARGEST of K' slot minima) */
constexpr int kFilterGroups = 256;                      /* the second filter handles lists up to this long; longer ones are cut by the streaming pass first */
constexpr int kMaxRanges = 160;
constexpr int kMaxSel = 512;                            /* keys entering the top-K selection */
constexpr int kRrWarps = 4;                             /* queries per CTA */
template <int METRIC, int R>
__global__ void __launch_bounds__(32 * kRrWarps) knn_rerank_kernel(const float* __restrict__ qkeys, int Q, const float* __restrict__ keys, int K,
                                                          int n_ranges, int n_db, const uint4* __restrict__ hq, const int* __restrict__ hq_cnt,
                                                          const int* __restrict__ slots, const float* __restrict__ kn2max, int id_mul, int id_add,
                                                          int32_t* __restrict__ out_ids, float* __restrict__ out_d2, int q_off,
                                                          int32_t* __restrict__ fail_list, int* __restrict__ fail_count, float* __restrict__ err_probe, int dev_flags, int kp,
                                                          int* __restrict__ slots_rw /* the same slots, to be reset for the next launch */,
                                                          int* __restrict__ next_fail_count /* the other call parity's counter: zeroed here */)
{
    __shared__ int sh_key[kRrWarps][kMaxGroups];        /* first key of the group */
    __shared__ float sh_g[kRrWarps][kMaxGroups];        /* its best prefilter score */
    __shared__ float sh_d[kRrWarps][kMaxSel];           /* exact distances / ids of the keys that can still make the top-K */
    __shared__ int sh_id[kRrWarps][kMaxSel];
    __shared__ int sh_pre[kRrWarps][kMaxRanges + 1];
    __shared__ float sh_q[kRrWarps][96];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.x * kRrWarps + warp;
    if (qi >= Q) return;                                /* whole warp */
    int* s_key = sh_key[warp]; float* s_g = sh_g[warp]; float* s_d = sh_d[warp]; int* s_id = sh_id[warp]; int* s_pre = sh_pre[warp];
    float* s_q = sh_q[warp];
    const unsigned lt_mask = (1u << lane) - 1u;
    const float inf = __int_as_float(0x7f800000);
    /* first trip: slots, the query's key, the queue counts, the largest key norm: all independent */
    int gt = lane < kp ? __ldg(slots + (size_t)qi * kSlotStride + lane) : (int)0x80000000;
    const int direct = __ldg(slots + (size_t)qi * kSlotStride + kDirect);
    const float qv = lane < R ? __ldg(qkeys + (size_t)qi * R + lane) : 0.0f;
    const float qv2 = (R > 32 && lane + 32 < R) ? __ldg(qkeys + (size_t)qi * R + lane + 32) : 0.0f;
    const float qv3 = (R > 64 && lane + 64 < R) ? __ldg(qkeys + (size_t)qi * R + lane + 64) : 0.0f;
    const float knmax = __ldg(kn2max);
    constexpr int kCntPerLane = (kMaxRanges + 31) / 32;
    int cnt[kCntPerLane];
#pragma unroll
    for (int u = 0; u < kCntPerLane; u++) {
        const int r = u * 32 + lane;
        cnt[u] = r < n_ranges ? __ldg(hq_cnt + (size_t)qi * n_ranges + r) : 0;
    }
    /* The cut: the query's final union bound (the maximum of its K' slots). Every group that was not queued had all its
     * scores >= it, so every key scoring below it sits in a queued group; queued groups whose best score is above it
     * cannot be certified anyway and are skipped: about K' groups survive. */
    gt = min(__reduce_max_sync(0xffffffffu, gt), direct);   /* every threshold a thread applied was >= this */
    const float cut = gt < 0x7f000000 ? ordered_float(gt) : inf;
    /* housekeeping that used to be two memsets per batch: this query's slots go back to "no key yet" for the next launch
     * (nobody else reads them), and the fail counter of the NEXT call is zeroed (calls alternate between two counters) */
    if (lane < kSlotStride) slots_rw[(size_t)qi * kSlotStride + lane] = kNoThr;
    if (qi == 0 && lane == 0) *next_fail_count = 0;
    if (lane < R) s_q[lane] = qv;
    if (R > 32 && lane + 32 < R) s_q[lane + 32] = qv2;
    if (R > 64 && lane + 64 < R) s_q[lane + 64] = qv3;
    bool overflow = false;
    int carry = 0;                                      /* exclusive prefix sums of the counts, 32 ranges at a time */
#pragma unroll
    for (int u = 0; u < kCntPerLane; u++) {
        int v = cnt[u];
        if (v > kQueueCap) { overflow = true; v = kQueueCap; }
        int incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
        const int r = u * 32 + lane;
        if (r < n_ranges) s_pre[r + 1] = carry + incl;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_pre[0] = 0;
    overflow = __any_sync(0xffffffffu, overflow);
    __syncwarp();
    const int total = carry;
    if (dev_flags & 128) { if (lane == 0) out_ids[(size_t)qi * K] = total; return; }
    float qn = 0.0f;
    for (int d = 0; d < R; d++) qn = fmaf(s_q[d], s_q[d], qn);
    const float sn = sqrtf(qn) + sqrtf(knmax);
    const float eps0 = 3.0517578125e-05f * sn * sn;                 /* 2^-15 (|q| + |k|max)^2 */
    /* Which groups VOTE in the bounds below. A group's minimum is the prefilter score of one of its keys, a distinct key per
     * group. It counts as "a neighbour this close exists" unless (a) the group straddles the search bound (the minimum may
     * belong to a key outside it) or (b), libnabo flavour, the key may be one the self-match rule removes (exact
     * d2 <= FLT_EPSILON, i.e. score + |q|^2 <= FLT_EPSILON + eps). Groups that do not vote are always kept. */
    const float self_lim = FLT_EPSILON + eps0 - qn;
    /* The queue entries, UE per lane in flight; survivors are compacted with a ballot. If more groups survive the cut than the
     * list holds (a smooth trajectory: the union bound is only as tight as the K'-th nearest TILE, and every key of the dozen
     * tiles around the query's place passes it), the cut is tightened and the pass repeated: the K-th smallest voting group
     * minimum m_K over ALL queued groups is undercut by K distinct keys, so the K-th nearest neighbour has exact
     * d2 <= m_K + |q|^2 + eps and a voting group above m_K + 2 eps holds no key that can reach or tie with the top-K
     * (one streaming pass: every lane keeps the K smallest of its share in registers, K rounds of warp minimum merge them). */
    float cut_t = cut;
    int n_grp = 0;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; attempt++) {
        n_grp = 0;
        constexpr int UE = 4;
        int lo = 0;                                     /* the range that holds this lane's entry: advances monotonically */
        for (int base = 0; base < total; base += 32 * UE) {
            uint4 ea[UE]; uint32_t eb[UE];
#pragma unroll
            for (int u = 0; u < UE; u++) {
                const int g = base + u * 32 + lane;
                ea[u] = make_uint4(0x7fffffffu, 0x7f800000u, 0x7f800000u, 0x7f800000u); eb[u] = 0x7f800000u;
                if (g < total) {
                    while (s_pre[lo + 1] <= g) lo++;
                    const uint4* ep = hq + ((size_t)qi * n_ranges + lo) * (size_t)(2 * kQueueCap) + 2 * (g - s_pre[lo]);
                    ea[u] = __ldcg(ep);
                    eb[u] = __ldcg(reinterpret_cast<const uint32_t*>(ep + 1));
                }
            }
#pragma unroll
            for (int u = 0; u < UE; u++) {
                const float gm[4] = {__uint_as_float(ea[u].y), __uint_as_float(ea[u].z), __uint_as_float(ea[u].w), __uint_as_float(eb[u])};
#pragma unroll
                for (int j = 0; j < 4; j++) {           /* the four 8-key groups of the chunk */
                    const int key = (int)ea[u].x + 8 * j;
                    const bool votes = key + 7 < n_db && (METRIC == 0 || gm[j] > self_lim);
                    const bool keep = gm[j] <= (votes ? cut_t : cut) && key < n_db && ea[u].x != 0x7fffffffu;
                    const unsigned m = __ballot_sync(0xffffffffu, keep);
                    const int pos = n_grp + __popc(m & lt_mask);
                    if (keep && pos < kMaxGroups) { s_key[pos] = key; s_g[pos] = gm[j]; }
                    n_grp += __popc(m);
                }
            }
        }
        if (n_grp <= kFilterGroups || attempt == 1) break;
        float best[16];
#pragma unroll
        for (int i = 0; i < 16; i++) best[i] = inf;
        int lo2 = 0;
        for (int g = lane; g < total; g += 32) {
            while (s_pre[lo2 + 1] <= g) lo2++;
            const uint4* ep = hq + ((size_t)qi * n_ranges + lo2) * (size_t)(2 * kQueueCap) + 2 * (g - s_pre[lo2]);
            const uint4 a = __ldcg(ep);
            const uint32_t b = __ldcg(reinterpret_cast<const uint32_t*>(ep + 1));
            const float gm[4] = {__uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w), __uint_as_float(b)};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int key = (int)a.x + 8 * j;
                const bool votes = key + 7 < n_db && gm[j] == gm[j] && (METRIC == 0 || gm[j] > self_lim);
                float v = votes ? gm[j] : inf;
#pragma unroll
                for (int i = 0; i < 16; i++) {           /* bubble v through the sorted list (registers: no dynamic indexing) */
                    if (i < K) { const float lw = fminf(best[i], v); v = fmaxf(best[i], v); best[i] = lw; }
                }
            }
        }
        const int inf_img = ordered_int(inf);
        float mk = inf; int have = 0;
        for (int r = 0; r < K; r++) {
            const int head = ordered_int(best[0]);
            const int w = __reduce_min_sync(0xffffffffu, head);
            if (w >= inf_img) break;
            const unsigned holders = __ballot_sync(0xffffffffu, head == w);      /* equal minima are distinct keys: one is taken per round */
            if (lane == __ffs(holders) - 1) {
#pragma unroll
                for (int i = 0; i < 15; i++) best[i] = best[i + 1];
                best[15] = inf;
            }
            mk = ordered_float(w); have++;
        }
        if (have < K || !(mk + 2.0f * eps0 < cut)) break;           /* nothing to gain: the query is redone exactly */
        cut_t = mk + 2.0f * eps0;
        __syncwarp();
    }
    if (n_grp > kMaxGroups) { overflow = true; n_grp = kMaxGroups; }
    __syncwarp();
    if (dev_flags & 256) { if (lane == 0) out_ids[(size_t)qi * K] = n_grp; return; }
    /* a certified top-K lies wholly below cut + |q|^2 (see below): keys at or above it need not enter the selection */
    float d_lim = cut < inf ? cut + qn : inf;
    /* Second, tighter filter before the key rows are fetched. Every surviving voting group's minimum is the score of a
     * distinct key that is a valid neighbour, so the K-th smallest such minimum m_K is undercut by K keys: the K-th nearest
     * neighbour has exact d2 <= m_K + |q|^2 + eps, and a voting group whose minimum exceeds m_K + 2 eps holds no key that
     * can beat it. About K of the ~3 K' groups remain. Groups that do not vote (see above) are always kept. */
    if (n_grp > K && n_grp <= kFilterGroups) {
        constexpr int kPer = kFilterGroups / 32;       /* groups per lane */
        unsigned v[kPer]; bool whole[kPer];
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            const int c = lane + 32 * j;
            whole[j] = c < n_grp && s_key[c] + 7 < n_db && (METRIC == 0 || s_g[c] > self_lim);
            /* scores can be negative: order-preserving unsigned image of the float */
            const unsigned b = c < n_grp ? __float_as_uint(s_g[c]) : 0u;
            v[j] = whole[j] ? (b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u)) : 0xffffffffu;
        }
        unsigned mk = 0xffffffffu; int have = 0;
        for (int r = 0; r < K; r++) {
            unsigned loc = v[0];
#pragma unroll
            for (int j = 1; j < kPer; j++) loc = min(loc, v[j]);
            const unsigned w = __reduce_min_sync(0xffffffffu, loc);
            if (w == 0xffffffffu) break;
            /* remove ONE holder of the minimum (equal minima are distinct keys and count separately) */
            const unsigned holders = __ballot_sync(0xffffffffu, loc == w);
            if (lane == __ffs(holders) - 1) {
                bool done = false;
#pragma unroll
                for (int j = 0; j < kPer; j++) if (!done && v[j] == w) { v[j] = 0xffffffffu; done = true; }
            }
            mk = w; have++;
        }
        if (have == K) {
            const unsigned mb = mk ^ ((mk >> 31) ? 0x80000000u : 0xffffffffu);          /* back to float bits */
            const float lim = __uint_as_float(mb) + 2.0f * eps0;
            /* compact the list in place: positions only move down, lanes work in index order */
            int n_new = 0;
#pragma unroll
            for (int j = 0; j < kPer; j++) {
                const int c = lane + 32 * j;
                const int key = c < n_grp ? s_key[c] : 0; const float g = c < n_grp ? s_g[c] : 0.0f;
                const bool keep = c < n_grp && (!whole[j] || g <= lim);
                __syncwarp();
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                const int pos = n_new + __popc(m & lt_mask);
                if (keep) { s_key[pos] = key; s_g[pos] = g; }
                n_new += __popc(m);
                __syncwarp();
            }
            n_grp = n_new;
        }
    }
    float worst_err = 0.0f;
    int n_sel = 0;
    constexpr int UK = R <= 20 ? 4 : (R <= 40 ? 2 : 1); /* key rows in flight per lane */
    for (int base = 0; base < n_grp * 8; base += 32 * UK) {
        if (n_sel + 32 * UK > kMaxSel) {
            /* The selection list is about to fill up (hundreds of keys within the prefilter's resolution of the K-th best: a
             * flat stretch of a smooth trajectory). Keep its K best by (d2, id) and take their worst distance as the new
             * limit: a later key enters only if it can still reach the top-K (equal distances included, the final selection
             * decides them by id). */
            for (int r = 0; r < K && r < n_sel; r++) {
                float bd = inf; int bi = 0x7fffffff, bp = -1;
                for (int c = r + lane; c < n_sel; c += 32) {
                    const float d = s_d[c]; const int id = s_id[c];
                    if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; bp = c; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float od = __shfl_xor_sync(0xffffffffu, bd, off); const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    const int op = __shfl_xor_sync(0xffffffffu, bp, off);
                    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; bp = op; }
                }
                if (lane == 0 && bp >= 0 && bp != r) {
                    const float td = s_d[r]; const int ti = s_id[r];
                    s_d[r] = s_d[bp]; s_id[r] = s_id[bp]; s_d[bp] = td; s_id[bp] = ti;
                }
                __syncwarp();
            }
            if (n_sel > K) n_sel = K;
            if (n_sel == K) {
                const float dk = s_d[K - 1];
                const float up = dk == 0.0f ? 1.0e-45f : __int_as_float(__float_as_int(dk) + 1);     /* distances are non-negative: the next float up */
                d_lim = fminf(d_lim, up);
            }
            __syncwarp();
        }
        float4 kv[UK][R / 4];
        int id[UK];
#pragma unroll
        for (int u = 0; u < UK; u++) {
            const int c = base + u * 32 + lane;
            id[u] = c < n_grp * 8 ? s_key[c >> 3] + (c & 7) : n_db;
            if (id[u] < n_db) {
#pragma unroll
                for (int g = 0; g < R / 4; g++) kv[u][g] = __ldg(reinterpret_cast<const float4*>(keys + (size_t)id[u] * R) + g);
            }
        }
#pragma unroll
        for (int u = 0; u < UK; u++) {
            float d = inf;
            if (id[u] < n_db) {
                const int c = base + u * 32 + lane;
                d = exact_d2<METRIC, R>(s_q, kv[u]);
                /* the prefilter must not have OVER-estimated a key by more than eps (that is what the certificate relies on) */
                if (err_probe) worst_err = fmaxf(worst_err, (s_g[c >> 3] - (d - qn)) / eps0);
                if (METRIC == 1 && !(d > FLT_EPSILON)) d = inf;          /* libnabo self-match rule */
                if (!(d < (METRIC == 0 ? FLT_MAX : inf))) d = inf;       /* never accepted by the trees */
            }
            const bool keep = d < d_lim;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            const int pos = n_sel + __popc(m & lt_mask);
            if (keep && pos < kMaxSel) { s_d[pos] = d; s_id[pos] = id[u] * id_mul + id_add; }
            n_sel += __popc(m);
        }
    }
    if (err_probe) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) worst_err = fmaxf(worst_err, __shfl_xor_sync(0xffffffffu, worst_err, off));
        if (lane == 0 && worst_err > 0.0f) atomicMax(reinterpret_cast<int*>(err_probe), __float_as_int(worst_err));   /* non-negative floats order as ints */
    }
    int n_surv = n_sel;
    if (n_surv > kMaxSel) { overflow = true; n_surv = kMaxSel; }
    __syncwarp();
    if (dev_flags & 512) { if (lane == 0) out_ids[(size_t)qi * K] = n_surv; return; }
    /* K rounds of "smallest (d2, id) not picked yet" (a rank-by-counting selection measured slower: its cost grows with
     * the square of the list length, and the longest list of the batch sets the kernel's duration). Distances are
     * non-negative floats, whose bit patterns order like unsigned integers, so (d2, id) is ONE 64-bit key; lists of up
     * to 128 keys (the rule) live in registers, four per lane, and a round is a local minimum plus two warp reductions. */
    float dK = 0.0f; int found = 0;
    if (n_surv <= 128) {
        unsigned long long e[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = lane + 32 * j;
            e[j] = c < n_surv ? (((unsigned long long)__float_as_uint(s_d[c]) << 32) | (unsigned)s_id[c]) : ~0ull;
        }
        for (int r = 0; r < K; r++) {
            unsigned long long loc = e[0] < e[1] ? e[0] : e[1];
            const unsigned long long loc2 = e[2] < e[3] ? e[2] : e[3];
            loc = loc < loc2 ? loc : loc2;
            const unsigned hi = (unsigned)(loc >> 32);
            const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
            const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? (unsigned)loc : 0xffffffffu);
            const unsigned long long win = ((unsigned long long)mhi << 32) | mlo;
            if (win == ~0ull) break;
#pragma unroll
            for (int j = 0; j < 4; j++) if (e[j] == win) e[j] = ~0ull;     /* keys are distinct: exactly one owner */
            if (lane == 0) { out_ids[(size_t)qi * K + r] = (int)mlo; out_d2[(size_t)qi * K + r] = __uint_as_float(mhi); }
            dK = __uint_as_float(mhi); found++;
        }
    } else {
        float pd = -1.0f; int pi = -1;
        for (int r = 0; r < K; r++) {
            float bd = inf; int bi = 0x7fffffff;
            for (int c = lane; c < n_surv; c += 32) {
                const float d = s_d[c];
                const int id = s_id[c];
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
    }
    if (lane == 0) {
        for (int r = found; r < K; r++) { out_ids[(size_t)qi * K + r] = -1; out_d2[(size_t)qi * K + r] = FLT_MAX; }
        bool certified = !overflow;
        if (cut < inf) {
            /* dropped keys have S >= cut, i.e. exact d2 > cut + |q|^2 - eps */
            const float eps = eps0 + 1.0e-7f * (float)R * dK;       /* + exact-side rounding: R sequential float additions */
            certified = certified && (found == K) && (dK + eps < cut + qn);
        }
        if (!certified) { fail_list[atomicAdd(fail_count, 1)] = q_off + qi; atomicAdd(next_fail_count + 8, 1); }   /* + the engine's running total (third counter of the block) */
        if (err_probe) {                                            /* developer counters: list sizes */
            int* sz = reinterpret_cast<int*>(err_probe) + 6;
            atomicAdd(sz + 0, total); atomicAdd(sz + 1, n_grp); atomicAdd(sz + 2, n_sel); atomicAdd(sz + 3, 1); atomicMax(sz + 4, n_grp); atomicMax(sz + 5, total);
        }
        if (!certified && err_probe) {                              /* developer counters: why */
            int* why = reinterpret_cast<int*>(err_probe) + 1;
            if (overflow) atomicAdd(why + 0, 1);
            else if (found != K) atomicAdd(why + 1, 1);
            else atomicAdd(why + 2, 1);
            if (n_grp >= kMaxGroups) atomicAdd(why + 3, 1);
        }
    }
}

} // namespace

bool scl_knn_tc_supported(int R) { return R == 20 || R == 40 || R == 80; }

int scl_knn_tc_ranges(int Q)
{
    const int groups = (Q + kQPerCta - 1) / kQPerCta;
    int r = SCL_NUM_SMS / groups;
    return r < 1 ? 1 : r;
}
int scl_knn_tc_max_batch() { return 1024; }          /* larger batches are cut into launches of this many queries */
int scl_knn_tc_kprime() { return kKPrime; }
int scl_knn_tc_slot_stride() { return kSlotStride; }
size_t scl_knn_tc_queue_bytes() { return (size_t)kQueueCap * 32; }          /* per (query, range) */
size_t scl_knn_tc_image_bytes(int R, int n_keys)
{
    const size_t tiles = ((size_t)n_keys + kNT - 1) / kNT;
    return tiles * (R == 20 ? TcCfg<20>::IMG_TILE : R == 40 ? TcCfg<40>::IMG_TILE : TcCfg<80>::IMG_TILE);
}

cudaError_t scl_launch_key_image(const float* keys, const float* knorm, int k_lo, int k_hi, int R, unsigned char* img, cudaStream_t stream)
{
    if (k_hi <= k_lo) return cudaSuccess;
    const int blocks = (k_hi - k_lo + 127) / 128;
    if (R == 20) key_image_kernel<20><<<blocks, 128, 0, stream>>>(keys, knorm, k_lo, k_hi, img);
    else if (R == 40) key_image_kernel<40><<<blocks, 128, 0, stream>>>(keys, knorm, k_lo, k_hi, img);
    else if (R == 80) key_image_kernel<80><<<blocks, 128, 0, stream>>>(keys, knorm, k_lo, k_hi, img);
    else return cudaErrorNotSupported;
    return cudaGetLastError();
}

// Developer switches (per-role cycle counters, timing experiments that make results wrong) exist only in builds with
// -DSCL_DEV_SWITCHES; a release build has no environment lookups on the launch path.
#ifdef SCL_DEV_SWITCHES
static int dev_flags_env() { const char* f = getenv("SCL_TC_FLAGS"); return f ? atoi(f) : 0; }
static bool dev_times_env() { return getenv("SCL_TC_TIMES") != nullptr; }
#else
static constexpr int dev_flags_env() { return 0; }
#endif

template <int R>
static cudaError_t launch_tc(const float* qkeys, int Q, const unsigned char* img, int n_db, int n_ranges, int* slots,
                              uint4* hq, int* hq_cnt, int* dbg, int kp, int K, int stages, cudaStream_t stream)
{
    using C = TcCfg<R>;
    if (stages < 2) stages = 2;
    if (stages > C::NSTAGE) stages = C::NSTAGE;
    const uint32_t smem_bytes = C::total(stages);
    {
        /* a function attribute belongs to the current device's context: once per device, not once per process */
        static bool attr[64] = {false};
        static std::mutex mu;
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lk(mu);
        if (dev >= 64 || !attr[dev]) {
            e = cudaFuncSetAttribute(knn_tc_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::TOTAL);
#ifdef SCL_DEV_SWITCHES
            if (e == cudaSuccess) e = cudaFuncSetAttribute(knn_tc_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::TOTAL);
#endif
            if (e != cudaSuccess) return e;
            if (dev < 64) attr[dev] = true;
        }
    }
    const int groups = (Q + kQPerCta - 1) / kQPerCta;
    const int nb = groups * n_ranges;
#ifdef SCL_DEV_SWITCHES
    long long* times = nullptr;
    const bool want_times = dev_times_env();
    const int dev_flags = dev_flags_env();
    if (want_times) { cudaMalloc(&times, (size_t)nb * 16 * sizeof(long long)); cudaMemsetAsync(times, 0, (size_t)nb * 128, stream); }
    if (dev_flags & 64) {}
    else if (want_times) knn_tc_kernel<R, true><<<nb, kThreads, smem_bytes, stream>>>(qkeys, Q, img, n_db, n_ranges, times, slots, hq, hq_cnt, dbg, dev_flags, kp, K, stages);
    else knn_tc_kernel<R, false><<<nb, kThreads, smem_bytes, stream>>>(qkeys, Q, img, n_db, n_ranges, nullptr, slots, hq, hq_cnt, dbg, dev_flags, kp, K, stages);
    if (want_times) {
        std::vector<long long> h((size_t)nb * 16);
        cudaStreamSynchronize(stream);
        cudaMemcpy(h.data(), times, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        double a[16] = {0}, amax4 = 0;
        for (int b = 0; b < nb; b++) for (int i = 0; i < 16; i++) a[i] += (double)h[(size_t)b * 16 + i] / nb;
        for (int b = 0; b < nb; b++) if ((double)h[(size_t)b * 16 + 4] > amax4) amax4 = (double)h[(size_t)b * 16 + 4];
        fprintf(stderr, "[tc n_db %d] tiles/CTA %.0f | per tile, per epilogue warp: total %.0f cycles, waiting for the accumulator %.0f, first chunk %.0f, other seven %.0f, tail %.0f | chunks with a hit per warp %.0f of %.0f, "
                        "chunks queued per (query, range) %.1f (largest %.0f) | mma thread per tile: wait key tile %.0f, wait accumulators + issue %.0f | service sweeps %.0f\n",
                n_db, a[8], a[1] / 8 / a[8], a[0] / 8 / a[8], a[9] / 8 / a[8], a[10] / 8 / a[8], a[11] / 8 / a[8], a[2] / 8, a[8] * 8, a[3] / 8 / 32, amax4, a[6] / a[8], a[7] / a[8], a[5] / 2);
        cudaFree(times);
    }
#else
    SCL_PREFER_SMEM((knn_tc_kernel<R, false>));
    knn_tc_kernel<R, false><<<nb, kThreads, smem_bytes, stream>>>(qkeys, Q, img, n_db, n_ranges, nullptr, slots, hq, hq_cnt, dbg, 0, kp, K, stages);
#endif
    return cudaGetLastError();
}

cudaError_t scl_launch_knn_tc(const float* qkeys, int Q, const float* keys, const unsigned char* img, const float* kn2max, int n_db, int R, int K,
                               int metric, int id_mul, int id_add, KnnTcWorkspace ws, int32_t* out_ids, float* out_d2,
                               int32_t* fail_list, int* fail_count, int* next_fail_count, bool init_state, int stages, cudaStream_t stream)
{
    if (Q <= 0) return cudaSuccess;
    if (!scl_knn_tc_supported(R)) return cudaErrorNotSupported;
    if (K > kKPrime - 2) return cudaErrorInvalidValue;
    cudaError_t err = cudaSuccess;
    const int max_b = scl_knn_tc_max_batch();
    if (init_state) {
        /* first use of these buffers (or a call that was cut short): afterwards the re-rank kernel keeps them clean */
        err = cudaMemsetAsync(fail_count, 0, sizeof(int), stream);
        if (err == cudaSuccess) err = cudaMemsetAsync(ws.slots, 0x7f, (size_t)(Q < max_b ? Q : max_b) * kSlotStride * 4, stream);     /* 3.39e38: "no key yet" */
        if (err != cudaSuccess) return err;
    }
    for (int q0 = 0; q0 < Q; q0 += max_b) {
        const int Qc = Q - q0 < max_b ? Q - q0 : max_b;
        const int n_ranges = scl_knn_tc_ranges(Qc);
        if ((size_t)Qc * n_ranges > ws.capacity) return cudaErrorInvalidValue;
        const float* qk = qkeys + (size_t)q0 * R;
        if (R == 20) err = launch_tc<20>(qk, Qc, img, n_db, n_ranges, ws.slots, reinterpret_cast<uint4*>(ws.hq), ws.hq_cnt, reinterpret_cast<int*>(ws.err_probe), kprime_for(K), K, stages, stream);
        else if (R == 40) err = launch_tc<40>(qk, Qc, img, n_db, n_ranges, ws.slots, reinterpret_cast<uint4*>(ws.hq), ws.hq_cnt, reinterpret_cast<int*>(ws.err_probe), kprime_for(K), K, stages, stream);
        else err = launch_tc<80>(qk, Qc, img, n_db, n_ranges, ws.slots, reinterpret_cast<uint4*>(ws.hq), ws.hq_cnt, reinterpret_cast<int*>(ws.err_probe), kprime_for(K), K, stages, stream);
        if (err != cudaSuccess) return err;
#define SCL_RERANK(M, RR)                                                                                                              \
    SCL_PREFER_SMEM((knn_rerank_kernel<M, RR>));                                                                                          \
    knn_rerank_kernel<M, RR><<<(Qc + kRrWarps - 1) / kRrWarps, 32 * kRrWarps, 0, stream>>>(qk, Qc, keys, K, n_ranges, n_db, reinterpret_cast<const uint4*>(ws.hq), ws.hq_cnt,   \
                                                            ws.slots, kn2max, id_mul, id_add, out_ids + (size_t)q0 * K, out_d2 + (size_t)q0 * K, \
                                                            q0, fail_list, fail_count, ws.err_probe, dev_flags, kprime_for(K), ws.slots, next_fail_count)
        const int dev_flags = dev_flags_env();     /* 0 in release builds */
        if (dev_flags & 16) {}
        else if (R == 20) { if (metric == 0) { SCL_RERANK(0, 20); } else { SCL_RERANK(1, 20); } }
        else if (R == 40) { if (metric == 0) { SCL_RERANK(0, 40); } else { SCL_RERANK(1, 40); } }
        else { if (metric == 0) { SCL_RERANK(0, 80); } else { SCL_RERANK(1, 80); } }
#undef SCL_RERANK
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
#ifdef SCL_DEV_SWITCHES
    if (ws.err_probe && getenv("SCL_TC_DEBUG")) {                   /* developer aid */
        int h[12];
        cudaStreamSynchronize(stream);
        {
            const int Qc = Q < max_b ? Q : max_b;
            std::vector<int> sl((size_t)Qc * kSlotStride);
            cudaMemcpy(sl.data(), ws.slots, sl.size() * 4, cudaMemcpyDeviceToHost);
            int per_slot[kKPrime] = {0};
            for (int q = 0; q < Qc; q++) for (int k = 0; k < kKPrime; k++) if (sl[(size_t)q * kSlotStride + k] >= 0x7f000000) per_slot[k]++;
            fprintf(stderr, "[tc slots of the last launch, unfilled per slot]");
            for (int k = 0; k < kKPrime; k++) fprintf(stderr, " %d", per_slot[k]);
            fprintf(stderr, "\n");
        }
        cudaMemcpy(h, ws.err_probe, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[tc Q %d n_db %d] worst |score error| / eps %.3f | uncertified so far: queue overflow %d, fewer than K survivors %d, certificate %d (survivor list full %d) | start-up waits timed out (warps) %d\n",
                Q, n_db, *reinterpret_cast<float*>(&h[0]), h[1], h[2], h[3], h[4], h[5]);
        if (h[9] > 0) fprintf(stderr, "[tc re-rank, averages over %d queries] queue entries %.0f (largest %d), surviving groups %.1f (largest %d), keys in the selection %.1f\n",
                              h[9], (double)h[6] / h[9], h[11], (double)h[7] / h[9], h[10], (double)h[8] / h[9]);
    }
#endif
    return cudaSuccess;
}

void scl_preload_k3_tc()
{
    SCL_TOUCH(key_image_kernel<20>); SCL_TOUCH(key_image_kernel<40>);
    SCL_TOUCH((knn_tc_kernel<20, false>)); SCL_TOUCH((knn_tc_kernel<40, false>));
    SCL_TOUCH((knn_rerank_kernel<0, 20>)); SCL_TOUCH((knn_rerank_kernel<1, 20>)); SCL_TOUCH((knn_rerank_kernel<0, 40>)); SCL_TOUCH((knn_rerank_kernel<1, 40>));
}

void scl_preload_k3_tc80()
{
    SCL_TOUCH(key_image_kernel<80>); SCL_TOUCH((knn_tc_kernel<80, false>)); SCL_TOUCH((knn_rerank_kernel<0, 80>)); SCL_TOUCH((knn_rerank_kernel<1, 80>));
}
