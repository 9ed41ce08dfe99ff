// k3_knn.cu — K3, the ring-key kNN that replaces the KD-trees of the reference:
//   nanoflann findNeighbors   /root/reference/include/descriptor.h:1714-1716 (nanoflann.hpp:1222-1242)
//   libnabo knn               /root/reference/include/descriptor.h:1631,1642
//
// This file holds the EXACT variant: brute force over all keys with the reference's own float
// accumulation order (nanoflann L2_Adaptor 4-wide groups, nanoflann.hpp:383-408, or libnabo's
// sequential loop), every operation an explicit round-to-nearest intrinsic, so the distances
// are bit-identical to the CPU path and the candidate set is identical except on exact ties
// (ties: lowest key first; the trees keep the first-visited). The tensor-core prefilter
// (k3_knn_tc.cu) only proposes candidates; what it proposes is re-ranked with the arithmetic here.
//
// Shape: grid = (key splits, query tiles of 128). A thread owns one query (its R key values in
// registers) and streams its split of the key matrix through a shared-memory tile (coalesced
// float4 loads, broadcast LDS.128 reads), keeping its own sorted top-K in shared memory
// ([K][128] layout, conflict free). A second kernel k-way-merges the per-split lists of a
// query with one warp.
//
// Roofline: the key matrix (4*R bytes/key) is streamed once per query tile from HBM/L2; the
// kernel is FP32-issue bound (3 ops per dimension per pair), not memory bound.
#include "common.cuh"
#include "kernels.h"

#include <cfloat>

namespace {

constexpr int kTQ = 128;      /* queries per CTA = threads per CTA */
constexpr int kTK = 64;       /* keys per shared-memory tile */
constexpr int kMaxK = 32;
constexpr int kMaxSplits = 256;

template <int R, int METRIC>
__device__ __forceinline__ float key_d2(const float (&q)[R], const float* __restrict__ k)
{
    float result = 0.0f;
    if (METRIC == 0) {
        /* nanoflann.hpp:391-397: result += d0*d0 + d1*d1 + d2*d2 + d3*d3 (left to right) */
#pragma unroll
        for (int d = 0; d + 3 < R; d += 4) {
            float4 kv;
            if (R % 4 == 0) kv = *reinterpret_cast<const float4*>(k + d);      /* broadcast LDS.128 */
            else kv = make_float4(k[d], k[d + 1], k[d + 2], k[d + 3]);
            const float d0 = __fsub_rn(q[d], kv.x), d1 = __fsub_rn(q[d + 1], kv.y);
            const float d2 = __fsub_rn(q[d + 2], kv.z), d3 = __fsub_rn(q[d + 3], kv.w);
            const float g = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
            result = __fadd_rn(result, g);
        }
#pragma unroll
        for (int d = R & ~3; d < R; d++) { const float d0 = __fsub_rn(q[d], k[d]); result = __fadd_rn(result, __fmul_rn(d0, d0)); }
    } else {
#pragma unroll
        for (int d = 0; d < R; d++) { const float d0 = __fsub_rn(q[d], k[d]); result = __fadd_rn(result, __fmul_rn(d0, d0)); }
    }
    return result;
}

// k-way merge of the per-split sorted lists of one query by one warp; order = (d2, id). The lists were written by other
// CTAs of the same launch: they are read through L2 (__ldcg). Every lane keeps the HEAD RECORDS of its (up to eight) lists in
// registers: one round trip fetches them all, and a round costs one more only for the lane whose head was taken. (The first
// version re-read every head in every round inside the compare chain, and the compiler kept those loads in program order:
// 16 dependent L2 round trips per round, 48 us for the merge of 256 splits — measured as the tail of knn_fallback_kernel,
// profiles/r02i_launches_summary.md — against ~4 us now.)
__device__ void merge_splits_warp(const int32_t* __restrict__ pi, const float* __restrict__ pd, int splits, int K, int lane,
                                  int32_t* __restrict__ out_ids, float* __restrict__ out_d2)
{
    constexpr int kPer = kMaxSplits / 32;
    const float inf = __int_as_float(0x7f800000);
    int head[kPer]; float hd[kPer]; int hi[kPer];
#pragma unroll
    for (int s = 0; s < kPer; s++) {
        const int sp = lane + 32 * s;
        head[s] = 0; hd[s] = inf; hi[s] = 0x7fffffff;
        if (sp < splits) { hd[s] = __ldcg(pd + (size_t)sp * K); hi[s] = __ldcg(pi + (size_t)sp * K); }
    }
    for (int r = 0; r < K; r++) {
        float bd = inf; int bi = 0x7fffffff; int bs = -1;
#pragma unroll
        for (int s = 0; s < kPer; s++)
            if (hd[s] < bd || (hd[s] == bd && hi[s] < bi)) { bd = hd[s]; bi = hi[s]; bs = s; }
        /* warp argmin on (d2, id) */
        float wd = bd; int wi = bi;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, wd, off); const int oi = __shfl_xor_sync(0xffffffffu, wi, off);
            if (od < wd || (od == wd && oi < wi)) { wd = od; wi = oi; }
        }
        const bool found = wi != 0x7fffffff;
        if (bs >= 0 && bi == wi && bd == wd && found) {     /* keys are distinct: exactly one lane owns the winner */
#pragma unroll
            for (int s = 0; s < kPer; s++) {
                if (s == bs) {
                    const int sp = lane + 32 * s;
                    head[s]++;
                    hd[s] = inf; hi[s] = 0x7fffffff;
                    if (head[s] < K) { hd[s] = __ldcg(pd + (size_t)sp * K + head[s]); hi[s] = __ldcg(pi + (size_t)sp * K + head[s]); }
                }
            }
        }
        if (lane == 0) {
            out_ids[r] = found ? wi : -1;
            out_d2[r] = found ? wd : FLT_MAX;
        }
    }
}

// bx, by, gdx: the CTA's split, its query tile and the number of splits; nthreads >= kTQ threads run it (threads beyond kTQ
// own no query and only help to stage the key tiles: the fallback kernel below runs this body in 256-thread CTAs).
template <int R, int METRIC>
__device__ __forceinline__ void exact_body(const float* __restrict__ qkeys, int Q, const float* __restrict__ keys,
                                           int n_db, int K, int split_len, int id_mul, int id_add,
                                           const int32_t* __restrict__ qlist, const int* __restrict__ qcount,
                                           int32_t* __restrict__ part_ids, float* __restrict__ part_d2,
                                           int* __restrict__ tickets /* [query tiles], zero between launches */,
                                           int32_t* __restrict__ out_ids, float* __restrict__ out_d2,
                                           int min_count /* lists of up to this many queries are taken by the thread-per-key body */,
                                           const int bx, const int by, const int gdx, const int nthreads)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sk = reinterpret_cast<float*>(smem_raw);                 /* [kTK][R] */
    float* ld = sk + kTK * R;                                       /* [K][kTQ] */
    int* li = reinterpret_cast<int*>(ld + K * kTQ);                 /* [K][kTQ] */
    const int t = threadIdx.x;
    /* optional indirection: only the queries in qlist[0..*qcount) (the tensor-core path's fallback) */
    const int nq = qlist ? *qcount : Q;
    if (by * kTQ >= nq || (qlist && nq <= min_count)) return;
    const int slot = by * kTQ + (t < kTQ ? t : 0);
    const bool active = t < kTQ && slot < nq;
    const int qi = active ? (qlist ? qlist[slot] : slot) : 0;
    float q[R];
#pragma unroll
    for (int d = 0; d < R; d++) q[d] = active ? __ldg(qkeys + (size_t)qi * R + d) : 0.0f;

    const int k0 = bx * split_len;
    const int k1 = min(n_db, k0 + split_len);
    int count = 0;
    const float limit = (METRIC == 0) ? FLT_MAX : __int_as_float(0x7f800000);
    float worst = limit;

    for (int base = k0; base < k1; base += kTK) {
        const int nk = min(kTK, k1 - base);
        __syncthreads();
        {   /* coalesced tile load; base*R*4 is a multiple of 16 because split_len and kTK are multiples of 4 */
            const float4* src = reinterpret_cast<const float4*>(keys + (size_t)base * R);
            float4* dst = reinterpret_cast<float4*>(sk);
            const int n4 = nk * R / 4;
            for (int i = t; i < n4; i += nthreads) dst[i] = __ldg(src + i);
            for (int i = n4 * 4 + t; i < nk * R; i += nthreads) sk[i] = __ldg(keys + (size_t)base * R + i);
        }
        __syncthreads();
        if (!active) continue;
        for (int j = 0; j < nk; j++) {
            const float d2 = key_d2<R, METRIC>(q, sk + j * R);
            if (METRIC == 1 && !(d2 > FLT_EPSILON)) continue;       /* libnabo self-match rule */
            if (!(d2 < worst)) continue;                             /* strict <: ties with the worst are rejected */
            int i = count < K ? count : K - 1;
            for (; i > 0 && ld[(i - 1) * kTQ + t] > d2; --i) {
                ld[i * kTQ + t] = ld[(i - 1) * kTQ + t];
                li[i * kTQ + t] = li[(i - 1) * kTQ + t];
            }
            ld[i * kTQ + t] = d2;
            li[i * kTQ + t] = (base + j) * id_mul + id_add;
            if (count < K) count++;
            if (count == K) worst = ld[(K - 1) * kTQ + t];
        }
    }
    if (active) {
        const size_t o = ((size_t)slot * gdx + bx) * K;
        for (int i = 0; i < K; i++) {
            part_d2[o + i] = i < count ? ld[i * kTQ + t] : __int_as_float(0x7f800000);
            part_ids[o + i] = i < count ? li[i * kTQ + t] : 0x7fffffff;
        }
    }
    /* The last CTA of this query tile to get here (ticket) merges the tile's per-split lists: one launch instead of two,
     * and a launch that has nothing to do (the tensor-core path's empty fallback list) ends at the early return above. */
    __threadfence();
    __syncthreads();
    __shared__ int s_last;
    if (t == 0) s_last = (atomicAdd(&tickets[by], 1) == gdx - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (t == 0) tickets[by] = 0;
    const int warp = t >> 5, lane = t & 31;
    for (int s = warp; s < kTQ; s += nthreads / 32) {
        const int sl = by * kTQ + s;
        if (sl >= nq) break;
        const int qq = qlist ? qlist[sl] : sl;
        merge_splits_warp(part_ids + (size_t)sl * gdx * K, part_d2 + (size_t)sl * gdx * K, gdx, K, lane,
                          out_ids + (size_t)qq * K, out_d2 + (size_t)qq * K);
    }
}

template <int R, int METRIC>
__global__ void __launch_bounds__(kTQ) knn_exact_kernel(const float* __restrict__ qkeys, int Q, const float* __restrict__ keys,
                                                        int n_db, int K, int split_len, int id_mul, int id_add,
                                                        const int32_t* __restrict__ qlist, const int* __restrict__ qcount,
                                                        int32_t* __restrict__ part_ids, float* __restrict__ part_d2, int* __restrict__ tickets,
                                                        int32_t* __restrict__ out_ids, float* __restrict__ out_d2, int min_count)
{
    exact_body<R, METRIC>(qkeys, Q, keys, n_db, K, split_len, id_mul, id_add, qlist, qcount, part_ids, part_d2, tickets, out_ids, out_d2, min_count,
                          (int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x, (int)blockDim.x);
}

// The same search for a HANDFUL of queries (the reference's own call pattern: one detect*LoopClosureID per keyframe).
// With thread = query a single query would leave 127 of 128 threads idle, so here thread = key: a CTA of 256 threads scores
// KPT * 256 keys of one query (rows read straight from global memory, 80 contiguous bytes per thread), then picks its K
// smallest (d2, id) in K rounds of block-wide argmin, and the last CTA of the query (ticket) merges the splits.
// grid = (key splits, queries).
constexpr int kSmallThreads = 256;
// bx, gdx: the CTA's key split and the number of splits; query qi, whose partial lists and ticket live in slot `slot`.
// MERGE = false: the per-split lists are left for the caller to merge (knn_fallback_kernel merges all listed queries at the end).
template <int R, int METRIC, int KPT, bool MERGE = true>
__device__ __forceinline__ void small_body(const float* __restrict__ qkeys, const float* __restrict__ keys, int n_db,
                                           int K, int split_len, int id_mul, int id_add, int32_t* __restrict__ part_ids,
                                           float* __restrict__ part_d2, int* __restrict__ tickets,
                                           int32_t* __restrict__ out_ids, float* __restrict__ out_d2,
                                           const int bx, const int gdx, const int slot, const int qi)
{
    __shared__ float s_d[kSmallThreads / 32];
    __shared__ int s_i[kSmallThreads / 32];
    __shared__ int s_last;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float inf = __int_as_float(0x7f800000);
    float q[R];
#pragma unroll
    for (int d = 0; d < R; d++) q[d] = __ldg(qkeys + (size_t)qi * R + d);
    const int k0 = bx * split_len;
    const int k1 = min(n_db, k0 + split_len);
    const float limit = (METRIC == 0) ? FLT_MAX : inf;
    float d[KPT];
#pragma unroll
    for (int i = 0; i < KPT; i++) {
        const int j = k0 + i * kSmallThreads + t;
        float dd = inf;
        if (j < k1) {
            dd = key_d2<R, METRIC>(q, keys + (size_t)j * R);
            if (METRIC == 1 && !(dd > FLT_EPSILON)) dd = inf;       /* libnabo self-match rule */
            if (!(dd < limit)) dd = inf;                             /* never accepted by the trees (NaN included) */
        }
        d[i] = dd;
    }
    const size_t o = ((size_t)slot * gdx + bx) * K;
    int r = 0;
    for (; r < K; r++) {
        float bd = inf; int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < KPT; i++) if (d[i] < bd) { bd = d[i]; bi = k0 + i * kSmallThreads + t; }   /* ascending keys: first minimum = lowest id */
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, off); const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) { s_d[warp] = bd; s_i[warp] = bi; }
        __syncthreads();
        float wd = s_d[0]; int wi = s_i[0];
#pragma unroll
        for (int w = 1; w < kSmallThreads / 32; w++) { const float od = s_d[w]; const int oi = s_i[w]; if (od < wd || (od == wd && oi < wi)) { wd = od; wi = oi; } }
        __syncthreads();
        if (wi == 0x7fffffff) break;                                 /* fewer than K acceptable keys in this split */
        if ((wi - k0) % kSmallThreads == t) {
            const int slot = (wi - k0) / kSmallThreads;
#pragma unroll
            for (int i = 0; i < KPT; i++) if (i == slot) d[i] = inf;
        }
        if (t == 0) { part_d2[o + r] = wd; part_ids[o + r] = wi * id_mul + id_add; }
    }
    if (t == 0) {
        for (; r < K; r++) { part_d2[o + r] = inf; part_ids[o + r] = 0x7fffffff; }
        if (MERGE) {
            __threadfence();
            s_last = (atomicAdd(&tickets[slot], 1) == gdx - 1);
        }
    }
    if (!MERGE) return;
    __syncthreads();
    if (!s_last) return;                                /* block-uniform */
    __threadfence();
    if (t == 0) tickets[slot] = 0;
    if (warp == 0)
        merge_splits_warp(part_ids + (size_t)slot * gdx * K, part_d2 + (size_t)slot * gdx * K, gdx, K, lane,
                          out_ids + (size_t)qi * K, out_d2 + (size_t)qi * K);
}

// grid = (key splits, queries): a handful of direct queries
template <int R, int METRIC, int KPT>
__global__ void __launch_bounds__(kSmallThreads) knn_exact_small_kernel(const float* __restrict__ qkeys, const float* __restrict__ keys, int n_db,
                                                                        int K, int split_len, int id_mul, int id_add, int32_t* __restrict__ part_ids,
                                                                        float* __restrict__ part_d2, int* __restrict__ tickets,
                                                                        int32_t* __restrict__ out_ids, float* __restrict__ out_d2)
{
    small_body<R, METRIC, KPT>(qkeys, keys, n_db, K, split_len, id_mul, id_add, part_ids, part_d2, tickets, out_ids, out_d2,
                               (int)blockIdx.x, (int)gridDim.x, (int)blockIdx.y, (int)blockIdx.y);
}

// The fallback of the tensor-core path as ONE launch: the queries qlist[0 .. *qcount) that could not be certified. Normally
// there are none and every CTA leaves at its first branch. A short list (up to kSmallList queries: the rule when there is one
// at all) is taken by the first n_small CTAs, thread = key, one listed query after the other; a longer list by the remaining
// CTAs, thread = query (exact_body in 256-thread CTAs). Two launches — one per variant — cost 5 us per batch in launch gaps.
constexpr int kSmallList = 16;
template <int R, int METRIC, int KPT>
__global__ void __launch_bounds__(kSmallThreads) knn_fallback_kernel(const float* __restrict__ qkeys, int Q, const float* __restrict__ keys, int n_db, int K,
                                                                     int id_mul, int id_add, const int32_t* __restrict__ qlist, const int* __restrict__ qcount,
                                                                     int32_t* __restrict__ part_ids, float* __restrict__ part_d2, int* __restrict__ tickets,
                                                                     int32_t* __restrict__ out_ids, float* __restrict__ out_d2,
                                                                     int n_small, int split_len_small, int splits_gen, int split_len_gen)
{
    const int c = *qcount;
    if (c <= 0) return;
    if ((int)blockIdx.x < n_small) {
        if (c > kSmallList) return;
        for (int slot = 0; slot < c; slot++) {
            small_body<R, METRIC, KPT, false>(qkeys, keys, n_db, K, split_len_small, id_mul, id_add, part_ids, part_d2, tickets, out_ids, out_d2,
                                              (int)blockIdx.x, n_small, slot, qlist[slot]);
            __syncthreads();                             /* the body's shared scratch is reused by the next query */
        }
        /* ONE ticket for the whole list, and the last CTA merges every listed query, a warp each. A ticket and a merge per query
         * made a chain: the CTA that merges query s arrives late at query s + 1, is the last one there again, merges again ...
         * (five listed queries: 83 us of scans, 186 us of kernel; profiles/r02i_launches_summary.md). */
        __shared__ int s_last_all;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last_all = (atomicAdd(&tickets[0], 1) == n_small - 1);
        __syncthreads();
        if (!s_last_all) return;
        __threadfence();
        if (threadIdx.x == 0) tickets[0] = 0;
        for (int slot = (int)(threadIdx.x >> 5); slot < c; slot += kSmallThreads / 32) {
            const int qq = qlist[slot];
            merge_splits_warp(part_ids + (size_t)slot * n_small * K, part_d2 + (size_t)slot * n_small * K, n_small, K, (int)(threadIdx.x & 31),
                              out_ids + (size_t)qq * K, out_d2 + (size_t)qq * K);
        }
    } else {
        if (c <= kSmallList) return;
        const int g = (int)blockIdx.x - n_small;
        exact_body<R, METRIC>(qkeys, Q, keys, n_db, K, split_len_gen, id_mul, id_add, qlist, qcount, part_ids, part_d2, tickets, out_ids, out_d2, 0,
                              g % splits_gen, g / splits_gen, splits_gen, (int)blockDim.x);
    }
}

template <int R>
cudaError_t launch_exact_small(const float* qkeys, int Q, const float* keys, int n_db, int K, int metric, int id_mul, int id_add,
                               KnnWorkspace ws, int32_t* out_ids, float* out_d2, cudaStream_t stream, bool* done)
{
    *done = false;
    int splits = (n_db + 1023) / 1024;
    if (splits > kMaxSplits) splits = kMaxSplits;
    if (splits < 1) splits = 1;
    int split_len = (n_db + splits - 1) / splits;
    split_len = (split_len + kSmallThreads - 1) / kSmallThreads * kSmallThreads;
    const int kpt = split_len / kSmallThreads;
    if (kpt > 32 || (size_t)Q * splits * K > ws.capacity) return cudaSuccess;        /* the general kernel takes it */
    dim3 grid(splits, Q);
#define SCL_SMALL(M, P) SCL_PREFER_SMEM((knn_exact_small_kernel<R, M, P>)); knn_exact_small_kernel<R, M, P><<<grid, kSmallThreads, 0, stream>>>(qkeys, keys, n_db, K, split_len, id_mul, id_add, ws.part_ids, \
                                                                                          ws.part_d2, ws.tickets, out_ids, out_d2)
    if (kpt <= 16) { if (metric == 0) { SCL_SMALL(0, 16); } else { SCL_SMALL(1, 16); } }
    else { if (metric == 0) { SCL_SMALL(0, 32); } else { SCL_SMALL(1, 32); } }
#undef SCL_SMALL
    *done = true;
    return cudaGetLastError();
}

// the uncertified queries of a tensor-core batch (qlist[0 .. *qcount), normally none): one launch, see knn_fallback_kernel
template <int R>
cudaError_t launch_fallback(const float* qkeys, int Q, const float* keys, int n_db, int K, int metric, int id_mul, int id_add,
                            const int32_t* qlist, const int* qcount, int splits_gen, int split_len_gen,
                            KnnWorkspace ws, int32_t* out_ids, float* out_d2, cudaStream_t stream, bool* done)
{
    *done = false;
    int splits = (n_db + 1023) / 1024;
    if (splits > kMaxSplits) splits = kMaxSplits;
    if (splits < 1) splits = 1;
    int split_len = (n_db + splits - 1) / splits;
    split_len = (split_len + kSmallThreads - 1) / kSmallThreads * kSmallThreads;
    const int kpt = split_len / kSmallThreads;
    const size_t smem = (size_t)kTK * R * 4 + (size_t)K * kTQ * 8;
    if (kpt > 32 || (size_t)kSmallList * splits * K > ws.capacity || smem > 48 * 1024) return cudaSuccess;   /* knn_exact_kernel alone takes it */
    const int n_gen = splits_gen * ((Q + kTQ - 1) / kTQ);
#define SCL_FB(M, P) SCL_PREFER_SMEM((knn_fallback_kernel<R, M, P>)); knn_fallback_kernel<R, M, P><<<splits + n_gen, kSmallThreads, smem, stream>>>(qkeys, Q, keys, n_db, K, id_mul, id_add, qlist, qcount, \
                                                      ws.part_ids, ws.part_d2, ws.tickets, out_ids, out_d2, splits, split_len, splits_gen, split_len_gen)
    if (kpt <= 16) { if (metric == 0) { SCL_FB(0, 16); } else { SCL_FB(1, 16); } }
    else { if (metric == 0) { SCL_FB(0, 32); } else { SCL_FB(1, 32); } }
#undef SCL_FB
    *done = true;
    return cudaGetLastError();
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ rows, int n, int width, float* __restrict__ dst)
{
    const size_t total = (size_t)n * width;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / width), c = (int)(i % width);
        dst[i] = __ldg(src + (size_t)rows[r] * width + c);
    }
}

__global__ void ids_to_local_kernel(const int32_t* __restrict__ ids, int n, int id_mul, int id_add, int missing_to,
                                    int32_t* __restrict__ ids_rewrite, int32_t* __restrict__ local)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = ids[i];
    local[i] = id < 0 ? missing_to : ((id - id_add) % id_mul == 0 ? (id - id_add) / id_mul : -1);   /* -1: owned by another shard */
    if (ids_rewrite != nullptr && id < 0) ids_rewrite[i] = missing_to;
}

template <int R>
cudaError_t launch_exact(const float* qkeys, int Q, const float* keys, int n_db, int K, int metric, int splits, int split_len,
                         int id_mul, int id_add, const int32_t* qlist, const int* qcount, KnnWorkspace ws, int32_t* out_ids, float* out_d2,
                         cudaStream_t stream, int min_count)
{
    const size_t smem = (size_t)kTK * R * 4 + (size_t)K * kTQ * 8;
    dim3 grid(splits, (Q + kTQ - 1) / kTQ);
    if (metric == 0) {
        SCL_PREFER_SMEM((knn_exact_kernel<R, 0>));
        knn_exact_kernel<R, 0><<<grid, kTQ, smem, stream>>>(qkeys, Q, keys, n_db, K, split_len, id_mul, id_add, qlist, qcount, ws.part_ids, ws.part_d2,
                                                            ws.tickets, out_ids, out_d2, min_count);
    } else {
        SCL_PREFER_SMEM((knn_exact_kernel<R, 1>));
        knn_exact_kernel<R, 1><<<grid, kTQ, smem, stream>>>(qkeys, Q, keys, n_db, K, split_len, id_mul, id_add, qlist, qcount, ws.part_ids, ws.part_d2,
                                                            ws.tickets, out_ids, out_d2, min_count);
    }
    return cudaGetLastError();
}

} // namespace

int scl_knn_splits(int Q, int n_db)
{
    const int tiles = (Q + kTQ - 1) / kTQ;
    int splits = (4 * SCL_NUM_SMS + tiles - 1) / tiles;
    const int max_by_keys = (n_db + 4 * kTK - 1) / (4 * kTK);
    if (splits > max_by_keys) splits = max_by_keys;
    if (splits > kMaxSplits) splits = kMaxSplits;
    if (splits < 1) splits = 1;
    return splits;
}

cudaError_t scl_launch_knn_exact(const float* qkeys, int Q, const float* keys, int n_db, int R, int K, int metric,
                                 int id_mul, int id_add, const int32_t* qlist, const int* qcount, KnnWorkspace ws,
                                 int32_t* out_ids, float* out_d2, cudaStream_t stream)
{
    if (Q <= 0) return cudaSuccess;
    if (K < 1 || K > kMaxK) return cudaErrorInvalidValue;
    if (!qlist && Q <= 8 && ws.tickets && (R == 20 || R == 40 || R == 80)) {       /* a handful of queries: thread = key */
        bool done = false;
        const cudaError_t e = R == 20 ? launch_exact_small<20>(qkeys, Q, keys, n_db, K, metric, id_mul, id_add, ws, out_ids, out_d2, stream, &done)
                            : R == 40 ? launch_exact_small<40>(qkeys, Q, keys, n_db, K, metric, id_mul, id_add, ws, out_ids, out_d2, stream, &done)
                                      : launch_exact_small<80>(qkeys, Q, keys, n_db, K, metric, id_mul, id_add, ws, out_ids, out_d2, stream, &done);
        if (e != cudaSuccess || done) return e;
    }
    const int splits = scl_knn_splits(Q, n_db);
    int split_len = (n_db + splits - 1) / splits;
    split_len = (split_len + kTK - 1) / kTK * kTK;
    if (split_len < kTK) split_len = kTK;
    if ((size_t)Q * splits * K > ws.capacity || !ws.tickets) return cudaErrorInvalidValue;
    if (qlist && (R == 20 || R == 40 || R == 80)) {
        /* the tensor-core path's uncertified queries: a short list goes thread = key (thread = query took 8.5 ms for six queries
         * on a million smooth keys, thread = key 0.2 ms), a long one thread = query, both in one launch */
        bool done = false;
        const cudaError_t e = R == 20 ? launch_fallback<20>(qkeys, Q, keys, n_db, K, metric, id_mul, id_add, qlist, qcount, splits, split_len, ws, out_ids, out_d2, stream, &done)
                            : R == 40 ? launch_fallback<40>(qkeys, Q, keys, n_db, K, metric, id_mul, id_add, qlist, qcount, splits, split_len, ws, out_ids, out_d2, stream, &done)
                                      : launch_fallback<80>(qkeys, Q, keys, n_db, K, metric, id_mul, id_add, qlist, qcount, splits, split_len, ws, out_ids, out_d2, stream, &done);
        if (e != cudaSuccess || done) return e;
    }
    const int min_count = 0;
    cudaError_t err;
    if (R == 20) err = launch_exact<20>(qkeys, Q, keys, n_db, K, metric, splits, split_len, id_mul, id_add, qlist, qcount, ws, out_ids, out_d2, stream, min_count);
    else if (R == 40) err = launch_exact<40>(qkeys, Q, keys, n_db, K, metric, splits, split_len, id_mul, id_add, qlist, qcount, ws, out_ids, out_d2, stream, min_count);
    else if (R == 10) err = launch_exact<10>(qkeys, Q, keys, n_db, K, metric, splits, split_len, id_mul, id_add, qlist, qcount, ws, out_ids, out_d2, stream, min_count);
    else if (R == 80) err = launch_exact<80>(qkeys, Q, keys, n_db, K, metric, splits, split_len, id_mul, id_add, qlist, qcount, ws, out_ids, out_d2, stream, min_count);
    else return cudaErrorNotSupported;
    return err;
}

cudaError_t scl_launch_gather_rows(const float* src, const int32_t* rows, int n, int width, float* dst, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const size_t total = (size_t)n * width;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 16 * SCL_NUM_SMS) blocks = 16 * SCL_NUM_SMS;
    SCL_PREFER_SMEM(gather_rows_kernel);
    gather_rows_kernel<<<blocks, 256, 0, stream>>>(src, rows, n, width, dst);
    return cudaGetLastError();
}

cudaError_t scl_launch_ids_to_local(const int32_t* ids, int n, int id_mul, int id_add, int missing_to, int32_t* ids_rewrite,
                                    int32_t* local, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    SCL_PREFER_SMEM(ids_to_local_kernel);
    ids_to_local_kernel<<<(n + 255) / 256, 256, 0, stream>>>(ids, n, id_mul, id_add, missing_to, ids_rewrite, local);
    return cudaGetLastError();
}

void scl_preload_k3()
{
    SCL_TOUCH((knn_exact_kernel<10, 0>)); SCL_TOUCH((knn_exact_kernel<10, 1>)); SCL_TOUCH((knn_exact_kernel<20, 0>)); SCL_TOUCH((knn_exact_kernel<20, 1>));
    SCL_TOUCH((knn_exact_kernel<40, 0>)); SCL_TOUCH((knn_exact_kernel<40, 1>));
    SCL_TOUCH((knn_exact_small_kernel<20, 0, 16>)); SCL_TOUCH((knn_exact_small_kernel<20, 1, 16>)); SCL_TOUCH((knn_exact_small_kernel<20, 0, 32>)); SCL_TOUCH((knn_exact_small_kernel<20, 1, 32>));
    SCL_TOUCH((knn_exact_small_kernel<40, 0, 16>)); SCL_TOUCH((knn_exact_small_kernel<40, 1, 16>)); SCL_TOUCH((knn_exact_small_kernel<40, 0, 32>)); SCL_TOUCH((knn_exact_small_kernel<40, 1, 32>));
    SCL_TOUCH(gather_rows_kernel); SCL_TOUCH(ids_to_local_kernel);
    /* the 80-row keys of the row-key family (rowkey.cu) */
    SCL_TOUCH((knn_exact_kernel<80, 0>)); SCL_TOUCH((knn_exact_kernel<80, 1>));
    SCL_TOUCH((knn_exact_small_kernel<80, 0, 16>)); SCL_TOUCH((knn_exact_small_kernel<80, 1, 16>)); SCL_TOUCH((knn_exact_small_kernel<80, 0, 32>)); SCL_TOUCH((knn_exact_small_kernel<80, 1, 32>));
    SCL_TOUCH((knn_fallback_kernel<20, 0, 16>)); SCL_TOUCH((knn_fallback_kernel<20, 1, 16>)); SCL_TOUCH((knn_fallback_kernel<20, 0, 32>)); SCL_TOUCH((knn_fallback_kernel<20, 1, 32>));
    SCL_TOUCH((knn_fallback_kernel<40, 0, 16>)); SCL_TOUCH((knn_fallback_kernel<40, 1, 16>)); SCL_TOUCH((knn_fallback_kernel<40, 0, 32>)); SCL_TOUCH((knn_fallback_kernel<40, 1, 32>));
    SCL_TOUCH((knn_fallback_kernel<80, 0, 16>)); SCL_TOUCH((knn_fallback_kernel<80, 1, 16>)); SCL_TOUCH((knn_fallback_kernel<80, 0, 32>)); SCL_TOUCH((knn_fallback_kernel<80, 1, 32>));
}
