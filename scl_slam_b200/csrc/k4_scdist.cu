// k4_scdist.cu — K4, the shift-aligned column-cosine Scan Context distance.
//
// Replaces distanceBtnScanContext (/root/reference/include/descriptor.h:1538-1569) with its
// parts makeSectorkeyFromScancontext (:1477-1489), fastAlignUsingVkey (:1491-1511), circshift
// (:1376-1395) and distDirectSC (:1513-1536), and the candidate scan of
// detectInterLoopClosureID (:1721-1737), for every (query, candidate) pair of a batch.
//
// One CTA per query, one warp per candidate (warps loop when K exceeds the warps that fit in
// shared memory). The query descriptor and each candidate descriptor (R*S floats, contiguous
// in the database) are fetched with ONE bulk-copy instruction each (cp.async.bulk, the TMA
// engine; completion on an mbarrier), so the gather is issued by a single lane and overlaps
// the arithmetic of the other warps.
//
// Arithmetic is FP64 with the reference's own operation order — sequential sums in index
// order, explicit round-to-nearest mul/add (no FMA), IEEE sqrt and divide — so the distance,
// and therefore the argmin shift and the winning candidate, are bit-identical to the CPU path:
//   * lane <-> column for the column sums, norms and per-column cosine terms,
//   * lane <-> shift for the S sector-key alignment norms,
//   * one lane per window shift for the in-order sum over columns,
//   * window shifts visited in ascending order with strict <, candidates in kNN order with
//     strict < and the self-skip rule.
// Nothing is cached per database entry: sector keys and column norms are recomputed from the
// descriptor tile that is in shared memory anyway, so the only HBM traffic is the descriptors.
//
// Roofline: HBM gather, 4*R*S*(K+1) bytes per query; ~31k FP64 operations per pair at 20x60.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int kQExt = 8;                /* query rows are stored S + kQExt wide (the first columns repeated): circular windows read straight */
constexpr int kShiftChunk = 7;          /* the default window (radius 3) in one chunk; keeps two CTAs per SM at 20x60 */

struct ScLayout {
    int RS, S, warps;
    size_t off_q, off_c, off_qdd, off_vq, off_nq, off_warp, warp_stride, off_res, total;
};

__host__ __device__ inline ScLayout sc_layout(int R, int S, int K, int warps)
{
    ScLayout L;
    L.RS = R * S; L.S = S; L.warps = warps;
    size_t o = 16 * ((size_t)(1 + warps) * 8 / 16 + 1);        /* mbarriers */
    L.off_q = o; o += (size_t)L.RS * 4;
    L.off_c = o; o += (size_t)warps * L.RS * 4;
    o = (o + 15) / 16 * 16;
    L.off_qdd = o; o += (size_t)R * (S + kQExt) * 8;            /* the query tile widened to double once per CTA, rows S + kQExt wide */
    L.off_vq = o; o += (size_t)S * 8;
    L.off_nq = o; o += (size_t)S * 8;
    L.off_warp = o;
    L.warp_stride = (size_t)(3 + kShiftChunk) * S * 8;          /* vc (twice in a row: circular reads need no wrap), nc, sim[kShiftChunk][S] */
    o += (size_t)warps * L.warp_stride;
    L.off_res = o; o += (size_t)K * 16;                         /* dist[K] doubles, shift[K] ints */
    L.total = o;
    return L;
}

// sequential column statistics of a row-major R x S float tile: mean (sector key, :1477-1489)
// and Euclidean norm (col.norm(), :1523) in double
// Columns j0, j0+stride, ... (up to kJC of them) are reduced side by side: each column's sums keep the reference's
// sequential order (bit-exact), while the independent chains hide the FP64 add latency.
constexpr int kJC = 4;
__device__ __forceinline__ void column_stats_batch(const float* __restrict__ d, int R, int S, int j0, int stride,
                                                   double* __restrict__ mean, double* __restrict__ norm)
{
    for (int jb = j0; jb < S; jb += kJC * stride) {
        double s[kJC], ss[kJC];
#pragma unroll
        for (int c = 0; c < kJC; c++) { s[c] = 0.0; ss[c] = 0.0; }
        for (int r = 0; r < R; r++) {
#pragma unroll
            for (int c = 0; c < kJC; c++) {
                const int j = jb + c * stride;
                if (j < S) {
                    const double v = (double)d[r * S + j];
                    s[c] = __dadd_rn(s[c], v);
                    ss[c] = __dadd_rn(ss[c], __dmul_rn(v, v));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kJC; c++) {
            const int j = jb + c * stride;
            if (j < S) { mean[j] = __ddiv_rn(s[c], (double)R); norm[j] = __dsqrt_rn(ss[c]); }
        }
    }
}

// RT/ST: compile-time rings/sectors (20x60, 40x120) so the row loops unroll and the index arithmetic folds; 0 = run time.
template <int kMaxThreads, int kMinBlocks, int RT, int ST>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) scdist_kernel(
    const float* __restrict__ db_desc, const float* __restrict__ q_desc, const int32_t* __restrict__ q_local,
    const int32_t* __restrict__ q_ids, const int32_t* __restrict__ cand_local, const int32_t* __restrict__ cand_ids,
    int K, int R_arg, int S_arg, int search_radius, int use_bulk,
    double* __restrict__ cand_dist, int32_t* __restrict__ cand_shift,
    int32_t* __restrict__ best_id, double* __restrict__ best_dist, int32_t* __restrict__ best_shift)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int R = RT ? RT : R_arg, S = ST ? ST : S_arg;
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ScLayout L = sc_layout(R, S, K, warps);
    double* qdd = reinterpret_cast<double*>(smem + L.off_qdd);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* qd = reinterpret_cast<float*>(smem + L.off_q);
    float* cd = reinterpret_cast<float*>(smem + L.off_c) + (size_t)warp * L.RS;
    double* vq = reinterpret_cast<double*>(smem + L.off_vq);
    double* nq = reinterpret_cast<double*>(smem + L.off_nq);
    double* vc = reinterpret_cast<double*>(smem + L.off_warp + (size_t)warp * L.warp_stride);
    double* nc = vc + 2 * S;
    double* sim = nc + S;
    double* res_dist = reinterpret_cast<double*>(smem + L.off_res);
    int* res_shift = reinterpret_cast<int*>(res_dist + K);
    const int qi = blockIdx.x;
    const int RS = L.RS;
    const uint32_t bytes = (uint32_t)RS * 4u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 1 + warps; i++) scl_mbar_init(&bars[i], 1);
        scl_mbar_fence_init();
    }
    __syncthreads();

    const float* qsrc = q_desc ? q_desc + (size_t)qi * RS : db_desc + (size_t)q_local[qi] * RS;
    if (use_bulk) {
        if (threadIdx.x == 0) { scl_mbar_expect_tx(&bars[0], bytes); scl_bulk_g2s(qd, qsrc, bytes, &bars[0]); }
    } else {
        for (int i = threadIdx.x; i < RS; i += blockDim.x) qd[i] = __ldg(qsrc + i);
    }
    /* The candidates this engine holds (all of them on an unsharded engine, about K / world on a shard) are listed first, so
     * that the warps share the work evenly whatever the ownership pattern; slots of other shards report NaN at once. */
    __shared__ int s_owned[32];
    __shared__ int s_n_owned;
    if (warp == 0) {
        const int cl = lane < K ? cand_local[(size_t)qi * K + lane] : -1;
        const unsigned m = __ballot_sync(0xffffffffu, cl >= 0);
        if (cl >= 0) s_owned[__popc(m & ((1u << lane) - 1u))] = lane;
        if (lane == 0) s_n_owned = __popc(m);
        if (lane < K && cl < 0) {
            res_dist[lane] = __longlong_as_double(0x7ff8000000000000LL); res_shift[lane] = 0;
            if (cand_dist) cand_dist[(size_t)qi * K + lane] = __longlong_as_double(0x7ff8000000000000LL);
            if (cand_shift) cand_shift[(size_t)qi * K + lane] = 0;
        }
    }
    __syncthreads();
    const int n_owned = s_n_owned;
    /* first candidate of every warp goes in flight before anyone waits */
    int oi = warp;                                    /* position in the owned list */
    int it = oi < n_owned ? s_owned[oi] : K;
    int c_local = it < K ? cand_local[(size_t)qi * K + it] : -1;
    if (use_bulk && lane == 0 && c_local >= 0) {
        scl_mbar_expect_tx(&bars[1 + warp], bytes);
        scl_bulk_g2s(cd, db_desc + (size_t)c_local * RS, bytes, &bars[1 + warp]);
    }
    if (use_bulk) scl_mbar_wait(&bars[0], 0);
    else __syncthreads();
    column_stats_batch(qd, R, S, threadIdx.x, blockDim.x, vq, nq);
    for (int i = threadIdx.x; i < R * (S + kQExt); i += blockDim.x) {
        const int r = i / (S + kQExt), j = i - r * (S + kQExt);
        qdd[i] = (double)qd[r * S + (j < S ? j : j - S)];
    }
    __syncthreads();

    uint32_t parity = 0;
    for (; oi < n_owned; oi += warps) {
        double out_dist = __longlong_as_double(0x7ff8000000000000LL); /* NaN: candidate missing */
        int out_shift = 0;
        if (c_local >= 0) {
            if (use_bulk) { scl_mbar_wait(&bars[1 + warp], parity); parity ^= 1u; }
            else { for (int i = lane; i < RS; i += 32) cd[i] = __ldg(db_desc + (size_t)c_local * RS + i); __syncwarp(); }
            /* a. candidate sector key and column norms */
            column_stats_batch(cd, R, S, lane, 32, vc, nc);
            __syncwarp();
            for (int j = lane; j < S; j += 32) vc[S + j] = vc[j];
            __syncwarp();
            /* b. fastAlignUsingVkey: lane <-> shift, sequential over columns (:1496-1508). Shifted key element j of shift sh is
             * vc[(j - sh) mod S] = vc2[S - sh + j]: two shifts of this lane (sh, sh + 32) run side by side. */
            double bestn = 10000000.0; int bests = 0x7fffffff;
            constexpr int kSC = 2;
            for (int sb = lane; sb < S; sb += kSC * 32) {
                double ss[kSC]; const double* vs[kSC];
#pragma unroll
                for (int c = 0; c < kSC; c++) { ss[c] = 0.0; const int sh = sb + 32 * c; vs[c] = vc + (sh < S ? S - sh : 0); }
#pragma unroll 4
                for (int j = 0; j < S; j++) {
                    const double a = vq[j];
#pragma unroll
                    for (int c = 0; c < kSC; c++) {
                        const double d = __dsub_rn(a, vs[c][j]);
                        ss[c] = __dadd_rn(ss[c], __dmul_rn(d, d));
                    }
                }
#pragma unroll
                for (int c = 0; c < kSC; c++) {
                    const int sh = sb + 32 * c;
                    if (sh < S) {
                        const double nrm = __dsqrt_rn(ss[c]);
                        if (nrm < bestn) { bestn = nrm; bests = sh; }   /* ascending shift per lane: first minimum wins */
                    }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double on = __shfl_xor_sync(0xffffffffu, bestn, off); const int os = __shfl_xor_sync(0xffffffffu, bests, off);
                if (on < bestn || (on == bestn && os < bests)) { bestn = on; bests = os; }
            }
            const int align = (bests == 0x7fffffff) ? 0 : bests;   /* no norm below 1e7: argmin stays 0 (:1493) */
            /* c. window of shifts around the alignment (:1545-1566). The reference visits the shifts within search_radius of
             * the alignment in ascending order and keeps the first strict minimum, i.e. the smallest distance and, among equal
             * distances, the smallest shift. Here the window is walked in circular order from align - radius (consecutive
             * shifts read consecutive query columns) and that rule is applied to (distance, shift) pairs. */
            double min_sc = 10000000.0; int argmin_shift = 0; bool found = false;
            const int win = min(2 * search_radius + 1, S);
            int s_first = win == S ? 0 : align - search_radius; if (s_first < 0) s_first += S;
            for (int p0 = 0; p0 < win; p0 += kShiftChunk) {
                const int ns = min(kShiftChunk, win - p0);
                int s0 = s_first + p0; if (s0 >= S) s0 -= S;               /* shift of window position p0; position w has s0 + w (mod S) */
                int cnt[kShiftChunk];                                       /* columns that count, per shift (warp-uniform) */
#pragma unroll
                for (int w = 0; w < kShiftChunk; w++) cnt[w] = 0;
                if (ns == kShiftChunk && (S & 1) == 0) {
                    /* Full chunk: lane <-> a PAIR of adjacent candidate columns (cb, cb + 1). Window position w of column cb
                     * meets query column jb + w, and of column cb + 1 query column jb + w + 1: the pair shares kShiftChunk + 1
                     * consecutive doubles of the widened query row, read once (shared memory is the binding resource of this
                     * kernel). Every (column, shift) dot product still accumulates over the rows in order: unchanged bit for bit. */
                    for (int cb0 = 0; cb0 < S; cb0 += 64) {
                        const int cb = cb0 + 2 * lane;
                        const bool on = cb < S;                            /* S is even: cb + 1 < S as well */
                        const double nba = on ? nc[cb] : 0.0, nbb = on ? nc[cb + 1] : 0.0;
                        int jb = (on ? cb : 0) + s0; if (jb >= S) jb -= S;
                        double da[kShiftChunk], db[kShiftChunk];
#pragma unroll
                        for (int w = 0; w < kShiftChunk; w++) { da[w] = 0.0; db[w] = 0.0; }
                        if (on && ((nba != 0.0) | (nbb != 0.0))) {
                            const double* qrow = qdd + jb;
                            const float2* crow = reinterpret_cast<const float2*>(cd + cb);
#pragma unroll 2
                            for (int r = 0; r < R; r++) {
                                const float2 c2 = crow[r * (S / 2)];
                                const double a = (double)c2.x, b = (double)c2.y;
#pragma unroll
                                for (int w = 0; w <= kShiftChunk; w++) {
                                    const double qv = qrow[r * (S + kQExt) + w];
                                    if (w < kShiftChunk) da[w] = __dadd_rn(da[w], __dmul_rn(qv, a));
                                    if (w > 0) db[w - 1] = __dadd_rn(db[w - 1], __dmul_rn(qv, b));
                                }
                            }
                        }
#pragma unroll
                        for (int w = 0; w < kShiftChunk; w++) {
                            int ja = jb + w; if (ja >= S) ja -= S;
                            int jn = ja + 1; if (jn >= S) jn -= S;
                            const double naa = nq[ja], nab = nq[jn];
                            const bool ca = on && !((naa == 0.0) | (nba == 0.0)), cbn = on && !((nab == 0.0) | (nbb == 0.0));
                            if (on) {
                                sim[w * S + ja] = ca ? __ddiv_rn(da[w], __dmul_rn(naa, nba)) : 0.0;
                                sim[w * S + jn] = cbn ? __ddiv_rn(db[w], __dmul_rn(nab, nbb)) : 0.0;
                            }
                            cnt[w] += __popc(__ballot_sync(0xffffffffu, ca)) + __popc(__ballot_sync(0xffffffffu, cbn));
                        }
                    }
                } else
                /* lane <-> CANDIDATE column cb: each candidate element is widened to double once and meets the query columns
                 * cb + s0 + w of all the chunk's shifts (consecutive doubles of the widened query row). Every (j, s) dot
                 * product still accumulates over the rows in order, so the result is unchanged bit for bit. */
                for (int cb0 = 0; cb0 < S; cb0 += 32) {
                    const int cb = cb0 + lane;
                    const bool on = cb < S;
                    const double nb = on ? nc[cb] : 0.0;
                    int jb = (on ? cb : 0) + s0; if (jb >= S) jb -= S;      /* query column of window position 0 */
                    double dot[kShiftChunk];
#pragma unroll
                    for (int w = 0; w < kShiftChunk; w++) dot[w] = 0.0;
                    if (nb != 0.0) {
                        const double* qrow = qdd + jb;
                        const float* crow = cd + cb;
#pragma unroll 2
                        for (int r = 0; r < R; r++) {
                            const double b = (double)crow[r * S];
#pragma unroll
                            for (int w = 0; w < kShiftChunk; w++)
                                if (w < ns) dot[w] = __dadd_rn(dot[w], __dmul_rn(qrow[r * (S + kQExt) + w], b));
                        }
                    }
#pragma unroll
                    for (int w = 0; w < kShiftChunk; w++) {
                        if (w < ns) {                                      /* warp-uniform */
                            int j = jb + w; if (j >= S) j -= S;
                            const double na = nq[j];
                            const bool counts = on && !((na == 0.0) | (nb == 0.0));
                            if (on) sim[w * S + j] = counts ? __ddiv_rn(dot[w], __dmul_rn(na, nb)) : 0.0;
                            cnt[w] += __popc(__ballot_sync(0xffffffffu, counts));
                        }
                    }
                }
                __syncwarp();
                /* one lane per shift: in-order sum over the columns. Columns that do not count hold +0.0, and adding +0.0 never
                 * changes the running sum (it starts at +0.0 and a dot product that starts at +0.0 is never -0.0). */
                double dist = 0.0; int my_s = 0x7fffffff;
                if (lane < ns) {
                    int my_cnt = 0;
#pragma unroll
                    for (int w = 0; w < kShiftChunk; w++) if (w == lane) my_cnt = cnt[w];
                    const double* sp = sim + lane * S;
                    double sum = 0.0;
#pragma unroll 4
                    for (int j = 0; j < S; j++) sum = __dadd_rn(sum, sp[j]);
                    dist = __dsub_rn(1.0, __ddiv_rn(sum, (double)my_cnt));  /* 0/0 = NaN when no column counts (:1534) */
                    my_s = s0 + lane; if (my_s >= S) my_s -= S;
                }
                /* smallest (distance, shift) of the chunk; NaN and anything not below the running minimum never wins */
                bool cand = lane < ns && dist < 10000000.0;
                double bd = cand ? dist : 10000000.0; int bs = cand ? my_s : 0x7fffffff;
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) {                    /* kShiftChunk <= 8 lanes hold values */
                    const double od = __shfl_xor_sync(0xffffffffu, bd, off); const int os = __shfl_xor_sync(0xffffffffu, bs, off);
                    if (od < bd || (od == bd && os < bs)) { bd = od; bs = os; }
                }
                bd = __shfl_sync(0xffffffffu, bd, 0); bs = __shfl_sync(0xffffffffu, bs, 0);
                if (bs != 0x7fffffff && (bd < min_sc || (found && bd == min_sc && bs < argmin_shift))) { min_sc = bd; argmin_shift = bs; found = true; }
                __syncwarp();
            }
            out_dist = min_sc; out_shift = argmin_shift;
        }
        if (lane == 0) {
            res_dist[it] = out_dist; res_shift[it] = out_shift;
            if (cand_dist) cand_dist[(size_t)qi * K + it] = out_dist;
            if (cand_shift) cand_shift[(size_t)qi * K + it] = out_shift;
        }
        /* next candidate of this warp into the same tile */
        const int nxt = oi + warps < n_owned ? s_owned[oi + warps] : K;
        c_local = nxt < K ? cand_local[(size_t)qi * K + nxt] : -1;
        it = nxt;
        __syncwarp();
        if (use_bulk && lane == 0 && c_local >= 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            scl_mbar_expect_tx(&bars[1 + warp], bytes);
            scl_bulk_g2s(cd, db_desc + (size_t)c_local * RS, bytes, &bars[1 + warp]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        /* candidate scan in kNN order, strict <, query itself skipped (:1721-1737) */
        double min_dist = 10000000.0; int nn_align = 0, nn_idx = -1;
        const int self = q_ids ? q_ids[qi] : -1;
        for (int i = 0; i < K; i++) {
            const int id = cand_ids[(size_t)qi * K + i];
            if (id < 0) continue;
            if (res_dist[i] < min_dist && id != self) { min_dist = res_dist[i]; nn_align = res_shift[i]; nn_idx = id; }
        }
        if (best_id) best_id[qi] = nn_idx;
        if (best_dist) best_dist[qi] = min_dist;
        if (best_shift) best_shift[qi] = nn_align;
    }
}

// Multi-GPU merge: per query, world*K records -> global top-K by (d2, id), then the winner scan.
__global__ void merge_shards_kernel(int world, int Q, int K, const int32_t* __restrict__ q_ids,
                                    const int32_t* __restrict__ all_ids, const float* __restrict__ all_d2,
                                    const double* __restrict__ all_dist, const int32_t* __restrict__ all_shift,
                                    int32_t* __restrict__ out_ids, float* __restrict__ out_d2, double* __restrict__ out_dist,
                                    int32_t* __restrict__ out_shift, int32_t* __restrict__ best_id, double* __restrict__ best_dist,
                                    int32_t* __restrict__ best_shift)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    int head[16];
    for (int w = 0; w < world; w++) head[w] = 0;
    double min_dist = 10000000.0; int nn_align = 0, nn_idx = -1;
    const int self = q_ids ? q_ids[qi] : -1;
    for (int r = 0; r < K; r++) {
        int bw = -1; float bd = 0.f; int bi = 0;
        for (int w = 0; w < world; w++) {
            if (head[w] >= K) continue;
            const size_t o = ((size_t)w * Q + qi) * K + head[w];
            const int id = all_ids[o];
            if (id < 0) { head[w] = K; continue; }
            const float d = all_d2[o];
            if (bw < 0 || d < bd || (d == bd && id < bi)) { bw = w; bd = d; bi = id; }
        }
        int id = -1; float d2 = 3.402823466e+38f; double dist = __longlong_as_double(0x7ff8000000000000LL); int shift = 0;
        if (bw >= 0) {
            const size_t o = ((size_t)bw * Q + qi) * K + head[bw];
            id = bi; d2 = bd; dist = all_dist[o]; shift = all_shift[o];
            head[bw]++;
            if (dist < min_dist && id != self) { min_dist = dist; nn_align = shift; nn_idx = id; }
        }
        if (out_ids) out_ids[(size_t)qi * K + r] = id;
        if (out_d2) out_d2[(size_t)qi * K + r] = d2;
        if (out_dist) out_dist[(size_t)qi * K + r] = dist;
        if (out_shift) out_shift[(size_t)qi * K + r] = shift;
    }
    if (best_id) best_id[qi] = nn_idx;
    if (best_dist) best_dist[qi] = min_dist;
    if (best_shift) best_shift[qi] = nn_align;
}

// Two-phase multi-GPU exchange, phase 1: per query, world x K (id, d2) records (each rank's list in its own kNN
// order; rank blocks `stride` bytes apart) -> the global top-K by (d2, id). Every rank computes the same lists.
__global__ void merge_topk_kernel(int world, int Q, int K, const unsigned char* __restrict__ ids_base, const unsigned char* __restrict__ d2_base,
                                  size_t stride, int32_t* __restrict__ out_ids, float* __restrict__ out_d2)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    int head[16];
    for (int w = 0; w < world; w++) head[w] = 0;
    for (int r = 0; r < K; r++) {
        int bw = -1; float bd = 0.f; int bi = 0;
        for (int w = 0; w < world; w++) {
            if (head[w] >= K) continue;
            const size_t o = (size_t)qi * K + head[w];
            const int id = reinterpret_cast<const int32_t*>(ids_base + w * stride)[o];
            if (id < 0) { head[w] = K; continue; }
            const float d = reinterpret_cast<const float*>(d2_base + w * stride)[o];
            if (bw < 0 || d < bd || (d == bd && id < bi)) { bw = w; bd = d; bi = id; }
        }
        if (bw >= 0) head[bw]++;
        out_ids[(size_t)qi * K + r] = bw >= 0 ? bi : -1;
        out_d2[(size_t)qi * K + r] = bw >= 0 ? bd : 3.402823466e+38f;
    }
}

// Phase 2: every candidate's SC distance was computed by the rank that owns it (id mod world); pick it from that
// rank's block and run the winner scan in global kNN order (descriptor.h:1721-1737).
__global__ void combine_owned_kernel(int world, int Q, int K, const int32_t* __restrict__ q_ids, const int32_t* __restrict__ cand_ids,
                                     const unsigned char* __restrict__ dist_base, const unsigned char* __restrict__ shift_base, size_t stride,
                                     double* __restrict__ out_dist, int32_t* __restrict__ out_shift, int32_t* __restrict__ best_id,
                                     double* __restrict__ best_dist, int32_t* __restrict__ best_shift)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    double min_dist = 10000000.0; int nn_align = 0, nn_idx = -1;
    const int self = q_ids ? q_ids[qi] : -1;
    for (int r = 0; r < K; r++) {
        const size_t o = (size_t)qi * K + r;
        const int id = cand_ids[o];
        double dist = __longlong_as_double(0x7ff8000000000000LL); int shift = 0;
        if (id >= 0) {
            const int owner = id % world;
            dist = reinterpret_cast<const double*>(dist_base + owner * stride)[o];
            shift = reinterpret_cast<const int32_t*>(shift_base + owner * stride)[o];
            if (dist < min_dist && id != self) { min_dist = dist; nn_align = shift; nn_idx = id; }
        }
        if (out_dist) out_dist[o] = dist;
        if (out_shift) out_shift[o] = shift;
    }
    if (best_id) best_id[qi] = nn_idx;
    if (best_dist) best_dist[qi] = min_dist;
    if (best_shift) best_shift[qi] = nn_align;
}

} // namespace

cudaError_t scl_launch_scdist(const float* db_desc, const float* q_desc, const int32_t* q_local, const int32_t* q_ids,
                              const int32_t* cand_local, const int32_t* cand_ids, int Q, int K, int R, int S, int search_radius,
                              double* cand_dist, int32_t* cand_shift, int32_t* best_id, double* best_dist, int32_t* best_shift,
                              int owned_per_query /* expected candidates per query held here; <= 0: all K */, cudaStream_t stream)
{
    if (Q <= 0) return cudaSuccess;
    const size_t budget = 200 * 1024;
    int warps = K < 16 ? K : 16;
    /* a shard holds about K / world of a query's candidates: fewer warps per CTA then, so that more CTAs (queries) share an SM */
    if (owned_per_query > 0 && owned_per_query < warps) warps = owned_per_query < 2 ? 2 : owned_per_query;
    while (warps > 1 && sc_layout(R, S, K, warps).total > budget) warps--;
    const ScLayout L = sc_layout(R, S, K, warps);
    if (L.total > 227 * 1024) return cudaErrorNotSupported;
    const int use_bulk = ((R * S) % 4 == 0) && ((reinterpret_cast<uintptr_t>(db_desc) & 15) == 0) &&
                         (q_desc == nullptr || (reinterpret_cast<uintptr_t>(q_desc) & 15) == 0);
#define SCL_SCDIST_LAUNCH(MT, MB, RT, ST)                                                                                         \
    do {                                                                                                                       \
        cudaError_t ea = cudaFuncSetAttribute(scdist_kernel<MT, MB, RT, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total); \
        if (ea != cudaSuccess) return ea;                                                                                      \
        scdist_kernel<MT, MB, RT, ST><<<Q, warps * 32, L.total, stream>>>(db_desc, q_desc, q_local, q_ids, cand_local, cand_ids, K, R, S, \
                                                                          search_radius, use_bulk, cand_dist, cand_shift, best_id, best_dist, best_shift); \
    } while (0)
    /* up to 10 warps and <= 113 KB: two CTAs per SM (20 warps hide the FP64 latencies); otherwise one big CTA */
    const bool two = warps <= 10 && L.total <= 113 * 1024;
    if (R == 20 && S == 60) { if (two) SCL_SCDIST_LAUNCH(320, 2, 20, 60); else SCL_SCDIST_LAUNCH(512, 1, 20, 60); }
    else if (R == 40 && S == 120) { if (two) SCL_SCDIST_LAUNCH(320, 2, 40, 120); else SCL_SCDIST_LAUNCH(512, 1, 40, 120); }
    else { if (two) SCL_SCDIST_LAUNCH(320, 2, 0, 0); else SCL_SCDIST_LAUNCH(512, 1, 0, 0); }
#undef SCL_SCDIST_LAUNCH
    return cudaGetLastError();
}

cudaError_t scl_launch_merge_shards(int world, int Q, int K, const int32_t* q_ids, const int32_t* all_ids, const float* all_d2,
                                    const double* all_dist, const int32_t* all_shift, int32_t* out_ids, float* out_d2,
                                    double* out_dist, int32_t* out_shift, int32_t* best_id, double* best_dist, int32_t* best_shift,
                                    cudaStream_t stream)
{
    if (Q <= 0) return cudaSuccess;
    if (world < 1 || world > 16) return cudaErrorInvalidValue;
    merge_shards_kernel<<<(Q + 127) / 128, 128, 0, stream>>>(world, Q, K, q_ids, all_ids, all_d2, all_dist, all_shift,
                                                            out_ids, out_d2, out_dist, out_shift, best_id, best_dist, best_shift);
    return cudaGetLastError();
}

cudaError_t scl_launch_merge_topk(int world, int Q, int K, const void* ids_base, const void* d2_base, size_t rank_stride_bytes,
                                  int32_t* out_ids, float* out_d2, cudaStream_t stream)
{
    if (Q <= 0) return cudaSuccess;
    if (world < 1 || world > 16) return cudaErrorInvalidValue;
    merge_topk_kernel<<<(Q + 127) / 128, 128, 0, stream>>>(world, Q, K, static_cast<const unsigned char*>(ids_base),
                                                          static_cast<const unsigned char*>(d2_base), rank_stride_bytes, out_ids, out_d2);
    return cudaGetLastError();
}

cudaError_t scl_launch_combine_owned(int world, int Q, int K, const int32_t* q_ids, const int32_t* cand_ids, const void* dist_base,
                                     const void* shift_base, size_t rank_stride_bytes, double* out_dist, int32_t* out_shift,
                                     int32_t* best_id, double* best_dist, int32_t* best_shift, cudaStream_t stream)
{
    if (Q <= 0) return cudaSuccess;
    if (world < 1 || world > 16) return cudaErrorInvalidValue;
    combine_owned_kernel<<<(Q + 127) / 128, 128, 0, stream>>>(world, Q, K, q_ids, cand_ids, static_cast<const unsigned char*>(dist_base),
                                                             static_cast<const unsigned char*>(shift_base), rank_stride_bytes, out_dist, out_shift,
                                                             best_id, best_dist, best_shift);
    return cudaGetLastError();
}
