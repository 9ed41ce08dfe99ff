// k4_scdist.cu — K4, the shift-aligned column-cosine Scan Context distance.
//
// Replaces distanceBtnScanContext (/root/reference/include/descriptor.h:1538-1569) with its
// parts makeSectorkeyFromScancontext (:1477-1489), fastAlignUsingVkey (:1491-1511), circshift
// (:1376-1395) and distDirectSC (:1513-1536), and the candidate scan of
// detectInterLoopClosureID (:1721-1737), for every (query, candidate) pair of a batch.
//
// One CTA per query, one warp per candidate. The query descriptor, each candidate descriptor
// (R*S floats, contiguous in the database) and their cached column statistics (sector key and
// column norms, 2*S doubles per entry, written once at insert time: the reference recomputes
// them for both operands of every pair, :1541-1542) arrive by cp.async.bulk (the TMA engine)
// on mbarriers, issued by single lanes, so the gather overlaps the arithmetic of other warps.
//
// The result is bit-identical to the sequential-order CPU path (oracle/sc_oracle.cpp; parity with
// a vectorising Eigen build is within an ulp and unpinned). What makes it fast is
// that only the shifts that can WIN are evaluated in the reference's FP64 operation order:
//   1. alignment (:1491-1511): all S shifted sector-key distances in FP32 (lane <-> shift), then
//      the shifts within a rigorous error bound of the FP32 minimum (one, as a rule) in FP64 with
//      the reference's rounding sequence; ascending strict-< rule among them.
//   2. window (:1545-1566): the 2*radius+1 column-cosine distances in FP32 (lane <-> a pair of
//      candidate columns sharing 8 query columns), then the shifts within the error bound of the
//      FP32 minimum in FP64: per column the row-ordered dot product (a float*float product is exact
//      in double, so one fused multiply-add rounds exactly like the reference's mul-then-add),
//      IEEE divide, and the in-order sum over the columns.
//   A pair whose magnitudes leave the range where the FP32 bounds hold, or with more near-ties than
//   the candidate lists take, is evaluated for ALL shifts in FP64 (the path scl_set_scdist_mode(1)
//   forces for every pair: the tests compare both on adversarial inputs).
//
// Roofline: HBM gather, (4*R*S + 16*S) * (K+1) bytes per query.
#include "common.cuh"
#include "kernels.h"

#include <mutex>

namespace {

constexpr int kQExt = 8;                /* query rows are stored S + kQExt wide (the first columns repeated): circular windows read straight */
constexpr int kWin = 7;                 /* window positions per FP32 pass (the default window, radius 3, in one pass) */
constexpr int kMaxCand = 4;             /* shifts evaluated exactly per pass */
constexpr float kAlignTol = 2.4e-7f;    /* 4 u, u = 2^-24: see the alignment estimate in the kernel */
constexpr float kDistTol = 4.0e-5f;     /* window: |d32 - d| <= R u + 1e-6 (R-term FP32 dots against na nb, 9-deep sum): 6e-6 at R = 80; doubled + margin */
constexpr float kSafeHi = 1.0e15f, kSafeLo = 1.0e-15f;   /* magnitudes outside: no FP32 prefilter for the pair */

struct ScLayout {
    int RS, S, warps, use_qdd;
    size_t off_qd, off_qx, off_qdd, off_qs, off_vq32, off_warp, warp_stride;
    size_t w_cd, w_cs, w_vc32, w_nc32, w_T, w_d32, w_cnt, w_list;
    size_t off_res, total;
};

__host__ __device__ inline size_t al16(size_t x) { return (x + 15) / 16 * 16; }

__host__ __device__ inline ScLayout sc_layout(int R, int S, int K, int warps, int use_qdd = 1)
{
    ScLayout L;
    L.RS = R * S; L.S = S; L.warps = warps; L.use_qdd = use_qdd;
    size_t o = al16((size_t)(1 + warps) * 8);                  /* mbarriers */
    L.off_qd = o; o += al16((size_t)L.RS * 4);                 /* query descriptor as it arrives */
    L.off_qx = o; o += al16((size_t)R * (S + kQExt) * 4);      /* the same with rows S + kQExt wide */
    L.off_qdd = o; if (use_qdd) o += al16((size_t)L.RS * 8);   /* widened to double once per CTA (not on a shard: few pairs per query, shared memory buys more CTAs per SM) */
    L.off_qs = o; o += al16((size_t)2 * S * 8);                /* query sector key | column norms (double) */
    L.off_vq32 = o; o += al16((size_t)S * 4);
    L.off_warp = o;
    size_t w = 0;
    L.w_cd = w; w += al16((size_t)L.RS * 4);
    L.w_cs = w; w += al16((size_t)2 * S * 8);
    L.w_vc32 = w; w += al16((size_t)2 * S * 4);                /* twice in a row: circular reads need no wrap */
    L.w_nc32 = w; w += al16((size_t)S * 4);
    L.w_T = w; w += al16((size_t)kMaxCand * (S + 1) * 8);      /* per exact shift: squared differences / column similarities */
    L.w_d32 = w; w += al16((size_t)S * 4);                     /* FP32 distance per window position */
    L.w_cnt = w; w += al16((size_t)S * 4);                     /* columns that count per window position */
    L.w_list = w; w += 64;
    L.warp_stride = w;
    o += (size_t)warps * w;
    L.off_res = o; o += al16((size_t)K * 16);                  /* dist[K] doubles, shift[K] ints */
    L.total = o;
    return L;
}

struct ScArgs {
    ScLayout L;                                         /* computed once on the host */
    const float* db_desc; const double* db_stat;
    const float* q_desc; const double* q_stat;          /* fresh queries (+ their stats from K2), or null: queries are entries q_local[i] */
    const int32_t* q_local; const int32_t* q_ids;
    const int32_t* cand_local; const int32_t* cand_ids;  /* cand_local null: derived from the reported ids (id = local * id_mul + id_add) */
    int id_mul, id_add;
    int K, R, S, search_radius, use_bulk, exact_all;
    double* cand_dist; int32_t* cand_shift; int32_t* best_id; double* best_dist; int32_t* best_shift;
};

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ float warp_min_nan_last(float v)
{
    /* fminf returns the other operand for a NaN: NaNs never become the minimum unless everything is NaN */
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// the S-bit mask (z0..z3, bit k of word k / 32) rotated left by sh (0 <= sh < S): bit j of the result = bit (j - sh) mod S
__device__ __forceinline__ void rot_mask(unsigned z0, unsigned z1, unsigned z2, unsigned z3, int S, int sh,
                                         unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3)
{
    if (S <= 64) {
        const unsigned long long v = ((unsigned long long)z1 << 32) | z0;
        const unsigned long long keep = S == 64 ? ~0ull : ((1ull << S) - 1ull);
        const unsigned long long r = sh == 0 ? v : (((v << sh) | (v >> (S - sh))) & keep);
        r0 = (unsigned)r; r1 = (unsigned)(r >> 32); r2 = 0u; r3 = 0u;
    } else {
        const unsigned __int128 v = ((unsigned __int128)(((unsigned long long)z3 << 32) | z2) << 64) | (((unsigned long long)z1 << 32) | z0);
        const unsigned __int128 keep = S == 128 ? ~(unsigned __int128)0 : ((((unsigned __int128)1) << S) - 1);
        const unsigned __int128 r = sh == 0 ? v : (((v << sh) | (v >> (S - sh))) & keep);
        r0 = (unsigned)r; r1 = (unsigned)(r >> 32); r2 = (unsigned)(r >> 64); r3 = (unsigned)(r >> 96);
    }
}

// fastAlignUsingVkey for ALL shifts in the reference's order (:1496-1508): lane <-> shift, sequential over the columns.
// vc2 holds the candidate sector key twice in a row: element j of shift sh is vc[(j - sh) mod S] = vc2[S - sh + j].
__device__ __forceinline__ int align_exact_all(const double* __restrict__ vq, const double* __restrict__ vc2, int S, int lane)
{
    double bestn = 10000000.0; int bests = 0x7fffffff;
    for (int sb = lane; sb < S; sb += 64) {
        double ss[2]; const double* vs[2];
#pragma unroll
        for (int c = 0; c < 2; c++) { ss[c] = 0.0; const int sh = sb + 32 * c; vs[c] = vc2 + (sh < S ? S - sh : 0); }
#pragma unroll 4
        for (int j = 0; j < S; j++) {
            const double a = vq[j];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const double d = __dsub_rn(a, vs[c][j]);
                ss[c] = __dadd_rn(ss[c], __dmul_rn(d, d));
            }
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
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
    return bests == 0x7fffffff ? 0 : bests;                     /* no norm below 1e7: argmin stays 0 (:1493) */
}

// RT/ST: compile-time rings/sectors (20x60, 40x120) so the row loops unroll and the index arithmetic folds; 0 = run time.
// Registers: the ten-warp CTA is held to 88 per thread so that ONE such CTA fits beside a knn_tc_kernel CTA of another
// query lane on the same SM (384 x 96 + 320 x 88 registers = 65 024 of 65 536; shared memory 100 KB + 113 KB): the SC
// distances of one batch then run under the tensor-core pass of the next instead of after it.
template <int kMaxThreads, int kMinBlocks, int RT, int ST>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) __maxnreg__(kMaxThreads == 320 ? 88 : (kMaxThreads == 128 ? 80 : 128))
scdist_kernel(const ScArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int R = RT ? RT : a.R, S = ST ? ST : a.S, K = a.K;
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ScLayout& L = a.L;
    const int QP = S + kQExt;                                   /* pitch of the widened query rows */
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* qd = reinterpret_cast<float*>(smem + L.off_qd);
    float* qx = reinterpret_cast<float*>(smem + L.off_qx);
    double* qdd = reinterpret_cast<double*>(smem + L.off_qdd);
    double* vq = reinterpret_cast<double*>(smem + L.off_qs);
    double* nq = vq + S;
    float* vq32 = reinterpret_cast<float*>(smem + L.off_vq32);
    unsigned char* wbase = smem + L.off_warp + (size_t)(warp < L.warps ? warp : 0) * L.warp_stride;   /* helper warps never touch it */
    float* cd = reinterpret_cast<float*>(wbase + L.w_cd);
    double* vc = reinterpret_cast<double*>(wbase + L.w_cs);
    double* nc = vc + S;
    float* vc32 = reinterpret_cast<float*>(wbase + L.w_vc32);
    float* inc32 = reinterpret_cast<float*>(wbase + L.w_nc32);
    double* T = reinterpret_cast<double*>(wbase + L.w_T);
    float* d32 = reinterpret_cast<float*>(wbase + L.w_d32);
    int* cntw = reinterpret_cast<int*>(wbase + L.w_cnt);
    int* list = reinterpret_cast<int*>(wbase + L.w_list);
    double* res_dist = reinterpret_cast<double*>(smem + L.off_res);
    int* res_shift = reinterpret_cast<int*>(res_dist + K);
    const int qi = blockIdx.x;
    const int RS = L.RS;
    const uint32_t bytes = (uint32_t)RS * 4u, sbytes = (uint32_t)S * 16u;
    const int TP = S + 1;                                       /* pitch of T */

    const int slots = L.warps;                        /* candidate tiles in shared memory: warps 0 .. slots-1 score pairs, the others only help with the query side */
    if (threadIdx.x == 0) {
        for (int i = 0; i < 1 + slots; i++) scl_mbar_init(&bars[i], 1);
        scl_mbar_fence_init();
    }
    /* The candidates this engine holds (all of them on an unsharded engine, about K / world on a shard) are listed first, so
     * that the warps share the work evenly whatever the ownership pattern; slots of other shards report NaN at once. */
    __shared__ int s_owned[32];
    __shared__ int s_local[32];                       /* local key of every candidate slot (-1: not held here) */
    __shared__ int s_n_owned;
    __shared__ float s_q[4];                          /* centre of the query sector key, its centred squared norm, FP32-safe flag */
    __shared__ unsigned s_zq[4];                      /* bit j: query column j is not empty */
    __shared__ float s_red[4][32];
    if (warp == 0) {
        int cl = -1;
        if (lane < K) {
            if (a.cand_local) cl = a.cand_local[(size_t)qi * K + lane];
            else { const int id = a.cand_ids[(size_t)qi * K + lane]; cl = (id >= 0 && id % a.id_mul == a.id_add) ? id / a.id_mul : -1; }
        }
        s_local[lane] = cl;
        const unsigned m = __ballot_sync(0xffffffffu, cl >= 0);
        if (cl >= 0) s_owned[__popc(m & ((1u << lane) - 1u))] = lane;
        if (lane == 0) s_n_owned = __popc(m);
        if (lane < K && cl < 0) {
            res_dist[lane] = __longlong_as_double(0x7ff8000000000000LL); res_shift[lane] = 0;
            if (a.cand_dist) a.cand_dist[(size_t)qi * K + lane] = __longlong_as_double(0x7ff8000000000000LL);
            if (a.cand_shift) a.cand_shift[(size_t)qi * K + lane] = 0;
        }
    }
    __syncthreads();
    const int n_owned = s_n_owned;
    if (n_owned == 0) {                               /* nothing to score here (a shard that owns none of this query's candidates): nothing was fetched */
        if (threadIdx.x == 0) {
            if (a.best_id) a.best_id[qi] = -1;
            if (a.best_dist) a.best_dist[qi] = 10000000.0;
            if (a.best_shift) a.best_shift[qi] = 0;
        }
        return;
    }
    const float* qsrc = a.q_desc ? a.q_desc + (size_t)qi * RS : a.db_desc + (size_t)a.q_local[qi] * RS;
    const double* qssrc = a.q_desc ? a.q_stat + (size_t)qi * 2 * S : a.db_stat + (size_t)a.q_local[qi] * 2 * S;
    if (a.use_bulk) {
        if (threadIdx.x == 0) {
            scl_mbar_expect_tx(&bars[0], bytes + sbytes);
            scl_bulk_g2s(qd, qsrc, bytes, &bars[0]);
            scl_bulk_g2s(vq, qssrc, sbytes, &bars[0]);
        }
    } else {
        for (int i = threadIdx.x; i < RS; i += blockDim.x) qd[i] = __ldg(qsrc + i);
        for (int i = threadIdx.x; i < 2 * S; i += blockDim.x) vq[i] = __ldg(qssrc + i);
    }
    /* first candidate of every scoring warp goes in flight before anyone waits */
    int oi = warp < slots ? warp : n_owned;           /* position in the owned list */
    int it = oi < n_owned ? s_owned[oi] : K;
    int c_local = it < K ? s_local[it] : -1;
    if (a.use_bulk && lane == 0 && c_local >= 0) {
        scl_mbar_expect_tx(&bars[1 + warp], bytes + sbytes);
        scl_bulk_g2s(cd, a.db_desc + (size_t)c_local * RS, bytes, &bars[1 + warp]);
        scl_bulk_g2s(vc, a.db_stat + (size_t)c_local * 2 * S, sbytes, &bars[1 + warp]);
    }
    if (a.use_bulk) scl_mbar_wait(&bars[0], 0);
    else __syncthreads();
    /* ---- query side, once per CTA ------------------------------------------------------------------------------
     * qdd: the tile in double (exact evaluation). qx: the FP32 estimate's operand, every column divided by its norm
     * (empty columns: 0), rows S + kQExt wide. vq32: the sector key minus its mean (a common offset of both keys
     * leaves their differences unchanged and shrinks the FP32 error bound). */
    if (warp == 0) {
        float sm = 0.0f;
        for (int j = lane; j < S; j += 32) sm += (float)vq[j];
        sm = warp_sum(sm);
        if (lane == 0) s_q[0] = sm / (float)S;         /* NaN / inf here end up in the unsafe path below */
    }
    if (L.use_qdd) for (int i = threadIdx.x; i < RS; i += blockDim.x) qdd[i] = (double)qd[i];
    __syncthreads();
    const float centre = s_q[0];
    const double centre_d = (double)centre;
    float aq = 0.0f, nhi = 0.0f, nlo = 3.0e38f;
    for (int j0 = 32 * warp; j0 < S; j0 += blockDim.x) {          /* whole warps: the ballot below needs every lane */
        const int j = j0 + lane;
        bool nz = false;
        if (j < S) {
            const double n = nq[j];
            /* a norm is zero exactly when its column is all zeros; a nonzero norm stays nonzero in float */
            const float nf = n == 0.0 ? 0.0f : fmaxf((float)n, 1.0e-37f);
            nz = nf != 0.0f;
            if (nz) { nhi = fmaxf(nhi, nf); nlo = fminf(nlo, nf); }
            const float v = (float)(vq[j] - centre_d);
            vq32[j] = v;
            aq = fmaf(v, v, aq);
        }
        const unsigned zb = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) s_zq[j0 >> 5] = zb;
    }
    for (int i = threadIdx.x; i < R * QP; i += blockDim.x) {
        const int r = i / QP, j = i - r * QP;
        const int jj = j < S ? j : j - S;
        const double n = nq[jj];
        const float inv = n == 0.0 ? 0.0f : 1.0f / fmaxf((float)n, 1.0e-37f);
        qx[i] = qd[r * S + jj] * inv;
    }
    aq = warp_sum(aq); nhi = warp_max(nhi); nlo = -warp_max(-nlo);
    if (lane == 0) { s_red[0][warp] = aq; s_red[1][warp] = nhi; s_red[2][warp] = nlo; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float x = 0.0f, y = 0.0f, z = 3.0e38f;
        for (int w = 0; w < warps; w++) { x += s_red[0][w]; y = fmaxf(y, s_red[1][w]); z = fminf(z, s_red[2][w]); }
        s_q[1] = x;
        /* query magnitudes inside the range where the FP32 estimates and their error bounds hold */
        s_q[2] = (y <= kSafeHi && z >= kSafeLo && x <= kSafeHi * kSafeHi && fabsf(centre) <= kSafeHi) ? 1.0f : 0.0f;
    }
    __syncthreads();
    const float A32 = s_q[1];
    const bool q_safe = s_q[2] != 0.0f;
    const unsigned zq0 = s_zq[0], zq1 = S > 32 ? s_zq[1] : 0u, zq2 = S > 64 ? s_zq[2] : 0u, zq3 = S > 96 ? s_zq[3] : 0u;

    uint32_t parity = 0;
    for (; oi < n_owned; oi += slots) {
        double out_dist = __longlong_as_double(0x7ff8000000000000LL); /* NaN: candidate missing */
        int out_shift = 0;
        if (c_local >= 0) {
            if (a.use_bulk) { scl_mbar_wait(&bars[1 + warp], parity); parity ^= 1u; }
            else {
                for (int i = lane; i < RS; i += 32) cd[i] = __ldg(a.db_desc + (size_t)c_local * RS + i);
                for (int i = lane; i < 2 * S; i += 32) vc[i] = __ldg(a.db_stat + (size_t)c_local * 2 * S + i);
                __syncwarp();
            }
            /* a. FP32 copies of the candidate's statistics; is the pair inside the range where the FP32 bounds hold? */
            float ac = 0.0f, chi = 0.0f, clo = 3.0e38f;
            unsigned zc[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int h = 0; h < 4; h++) {
                const int j = lane + 32 * h;
                bool nz = false;
                if (j < S) {
                    const float v = (float)(vc[j] - centre_d);
                    vc32[j] = v; vc32[S + j] = v;
                    ac = fmaf(v, v, ac);
                    const double n = nc[j];
                    const float nf = n == 0.0 ? 0.0f : fmaxf((float)n, 1.0e-37f);
                    inc32[j] = nf == 0.0f ? 0.0f : 1.0f / nf;
                    nz = nf != 0.0f;
                    if (nz) { chi = fmaxf(chi, nf); clo = fminf(clo, nf); }
                }
                zc[h] = __ballot_sync(0xffffffffu, nz);
            }
            ac = warp_sum(ac); chi = warp_max(chi); clo = -warp_max(-clo);
            const float C32 = ac;
            const bool safe = !a.exact_all && q_safe && chi <= kSafeHi && clo >= kSafeLo && C32 <= kSafeHi * kSafeHi;
            __syncwarp();
            /* b. fastAlignUsingVkey (:1491-1511) */
            int align = -1;
            if (safe) {
                /* FP32 estimate of every shift's squared distance |vq|^2 + |vc|^2 - 2 corr(s) on the centred keys: lane <-> shifts
                 * lane, lane + 32, ...; element j of shift sh is vc[(j - sh) mod S] = vc32[S - sh + j]. Two accumulators per shift. */
                float best32 = 3.0e38f;
                float ss32[4];                                  /* S <= 128 */
                const float base = A32 + C32;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int sb = lane + 64 * h;
                    float c0a = 0.0f, c0b = 0.0f, c1a = 0.0f, c1b = 0.0f;
                    const float* p0 = vc32 + (sb < S ? S - sb : 0);
                    const float* p1 = vc32 + (sb + 32 < S ? S - sb - 32 : 0);
                    if (64 * h < S) {
#pragma unroll 2
                        for (int j = 0; j + 3 < S; j += 4) {
                            const float4 x = *reinterpret_cast<const float4*>(vq32 + j);     /* broadcast */
                            c0a = fmaf(x.x, p0[j], c0a); c1a = fmaf(x.x, p1[j], c1a);
                            c0b = fmaf(x.y, p0[j + 1], c0b); c1b = fmaf(x.y, p1[j + 1], c1b);
                            c0a = fmaf(x.z, p0[j + 2], c0a); c1a = fmaf(x.z, p1[j + 2], c1a);
                            c0b = fmaf(x.w, p0[j + 3], c0b); c1b = fmaf(x.w, p1[j + 3], c1b);
                        }
                        for (int j = S & ~3; j < S; j++) { const float x = vq32[j]; c0a = fmaf(x, p0[j], c0a); c1a = fmaf(x, p1[j], c1a); }
                    }
                    ss32[2 * h] = sb < S ? base - 2.0f * (c0a + c0b) : __int_as_float(0x7fc00000);
                    ss32[2 * h + 1] = sb + 32 < S ? base - 2.0f * (c1a + c1b) : __int_as_float(0x7fc00000);
                    best32 = fminf(best32, fminf(ss32[2 * h], ss32[2 * h + 1]));
                }
                best32 = warp_min_nan_last(best32);
                /* |ss32 - ss| <= (S + 4) u (|vq| + |vc|)^2, u = 2^-24 (S u from the sums, 2 u from the final subtraction, 2 u from rounding the
                 * centred keys to float): a shift that can win lies within twice that of the FP32 minimum;
                 * the factor below leaves a further 2x */
                const float sq = __fsqrt_rn(A32) + __fsqrt_rn(C32);
                const float lim = best32 + kAlignTol * (float)(S + 4) * sq * sq;
                int n_al = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {                   /* ascending shift: i-major, then lane */
                    const int sh = lane + 32 * i;
                    const bool isc = sh < S && !(ss32[i] > lim); /* a NaN is a candidate */
                    const unsigned m = __ballot_sync(0xffffffffu, isc);
                    const int pos = n_al + __popc(m & ((1u << lane) - 1u));
                    if (isc && pos < kMaxCand) list[pos] = sh;
                    n_al += __popc(m);
                }
                __syncwarp();
                if (n_al == 1 && best32 < 1.0e12f) {
                    /* one shift can win: it is the reference's argmin (its exact norm is far below the 1e7 the scan starts from) */
                    align = list[0];
                } else if (n_al >= 2 && n_al <= kMaxCand && best32 < 1.0e12f) {
                    /* the candidates in the reference's rounding sequence: lane <-> column for the squared differences, one
                     * lane per candidate for the in-order sum */
                    for (int k = 0; k < n_al; k++) {
                        const int sh = list[k];
                        for (int j = lane; j < S; j += 32) {
                            int jc = j - sh; if (jc < 0) jc += S;
                            const double d = __dsub_rn(vq[j], vc[jc]);
                            T[k * TP + j] = __dmul_rn(d, d);
                        }
                    }
                    __syncwarp();
                    double bestn = 10000000.0; int bests = 0x7fffffff;
                    if (lane < n_al) {
                        const double* tp = T + lane * TP;
                        double ss = 0.0;
#pragma unroll 4
                        for (int j = 0; j < S; j++) ss = __dadd_rn(ss, tp[j]);
                        const double nrm = __dsqrt_rn(ss);
                        if (nrm < bestn) { bestn = nrm; bests = list[lane]; }
                    }
#pragma unroll
                    for (int off = 2; off > 0; off >>= 1) {     /* kMaxCand <= 4 lanes hold values */
                        const double on = __shfl_xor_sync(0xffffffffu, bestn, off); const int os = __shfl_xor_sync(0xffffffffu, bests, off);
                        if (on < bestn || (on == bestn && os < bests)) { bestn = on; bests = os; }
                    }
                    bests = __shfl_sync(0xffffffffu, bests, 0);
                    align = bests == 0x7fffffff ? 0 : bests;
                    __syncwarp();
                }
            }
            if (align < 0) {
                /* every shift in FP64: the candidate sector key twice in a row in T (2 S doubles fit: kMaxCand >= 2) */
                for (int j = lane; j < S; j += 32) { const double v = vc[j]; T[j] = v; T[S + j] = v; }
                __syncwarp();
                align = align_exact_all(vq, T, S, lane);
                __syncwarp();
            }
            /* c. window of shifts around the alignment (:1545-1566). The reference visits the shifts within search_radius of
             * the alignment in ascending order and keeps the first strict minimum, i.e. the smallest distance and, among equal
             * distances, the smallest shift. The window is walked in circular order from align - radius (consecutive shifts
             * read consecutive query columns) and that rule is applied to (distance, shift) pairs. */
            const int win = min(2 * a.search_radius + 1, S);
            int s_first = win == S ? 0 : align - a.search_radius; if (s_first < 0) s_first += S;
            float min32 = 3.0e38f;
            if (safe) {
                for (int p0 = 0; p0 < win; p0 += kWin) {
                    const int ns = min(kWin, win - p0);
                    int s0 = s_first + p0; if (s0 >= S) s0 -= S;    /* shift of window position p0; position p0 + w has s0 + w (mod S) */
                    float sum[kWin];
#pragma unroll
                    for (int w = 0; w < kWin; w++) sum[w] = 0.0f;
                    for (int cb0 = 0; cb0 < S; cb0 += 64) {
                        /* lane <-> a PAIR of adjacent candidate columns (cb, cb + 1). Window position w of column cb meets query
                         * column jb + w, and of column cb + 1 query column jb + w + 1: the pair shares kWin + 1 consecutive floats
                         * of the widened, column-normalised query row (S is even here: odd S takes the exact path). jb has the
                         * parity of s0 for every lane, so the eight floats come as four (even) or three + two single (odd)
                         * 8-byte loads. */
                        const int cb = cb0 + 2 * lane;
                        const bool on = cb < S;
                        const float ia = on ? inc32[cb] : 0.0f, ib = on ? inc32[cb + 1] : 0.0f;
                        int jb = (on ? cb : 0) + s0; if (jb >= S) jb -= S;
                        float da[kWin], db[kWin];
#pragma unroll
                        for (int w = 0; w < kWin; w++) { da[w] = 0.0f; db[w] = 0.0f; }
                        if (on && ((ia != 0.0f) | (ib != 0.0f))) {
                            const float* qrow = qx + jb;
                            const float2* crow = reinterpret_cast<const float2*>(cd + cb);
                            if ((s0 & 1) == 0) {
#pragma unroll 4
                                for (int r = 0; r < R; r++) {
                                    const float2 c2 = crow[r * (S / 2)];
                                    const float2* q2 = reinterpret_cast<const float2*>(qrow + r * QP);
                                    const float2 t0 = q2[0], t1 = q2[1], t2 = q2[2], t3 = q2[3];
                                    const float qv[kWin + 1] = {t0.x, t0.y, t1.x, t1.y, t2.x, t2.y, t3.x, t3.y};
#pragma unroll
                                    for (int w = 0; w < kWin; w++) { da[w] = fmaf(qv[w], c2.x, da[w]); db[w] = fmaf(qv[w + 1], c2.y, db[w]); }
                                }
                            } else {
#pragma unroll 4
                                for (int r = 0; r < R; r++) {
                                    const float2 c2 = crow[r * (S / 2)];
                                    const float* q1 = qrow + r * QP;
                                    const float2* q2 = reinterpret_cast<const float2*>(q1 + 1);
                                    const float h0 = q1[0], h7 = q1[7];
                                    const float2 t0 = q2[0], t1 = q2[1], t2 = q2[2];
                                    const float qv[kWin + 1] = {h0, t0.x, t0.y, t1.x, t1.y, t2.x, t2.y, h7};
#pragma unroll
                                    for (int w = 0; w < kWin; w++) { da[w] = fmaf(qv[w], c2.x, da[w]); db[w] = fmaf(qv[w + 1], c2.y, db[w]); }
                                }
                            }
                        }
#pragma unroll
                        for (int w = 0; w < kWin; w++) sum[w] = fmaf(da[w], ia, fmaf(db[w], ib, sum[w]));
                    }
                    /* columns that count at shift sh: query column j and candidate column j - sh both not empty, i.e. the bits the
                     * query mask shares with the candidate mask rotated left by sh (lane w does window position p0 + w) */
                    int my_cnt = 0;
                    {
                        int sh = s0 + lane; if (sh >= S) sh -= S;
                        if (lane < ns) {
                            unsigned r0, r1, r2, r3;
                            rot_mask(zc[0], zc[1], zc[2], zc[3], S, sh, r0, r1, r2, r3);
                            my_cnt = __popc(r0 & zq0) + __popc(r1 & zq1) + __popc(r2 & zq2) + __popc(r3 & zq3);
                        }
                    }
#pragma unroll
                    for (int w = 0; w < kWin; w++) {
                        const float s = warp_sum(sum[w]);
                        if (lane == w && w < ns) {
                            d32[p0 + w] = my_cnt > 0 ? 1.0f - s / (float)my_cnt : __int_as_float(0x7fc00000);
                            cntw[p0 + w] = my_cnt;
                        }
                    }
                    float mine = (lane < ns && my_cnt > 0) ? d32[p0 + lane] : 3.0e38f;
                    if (!(mine == mine)) mine = 3.0e38f;          /* a NaN never becomes the minimum */
                    min32 = fminf(min32, warp_min_nan_last(mine));
                }
                __syncwarp();
            }
            /* exact evaluation of the shifts that can win, kMaxCand at a time */
            double min_sc = 10000000.0; int argmin_shift = 0; bool found = false;
            const float limd = min32 + kDistTol;
            int p_next = 0;
            while (p_next < win) {
                /* collect up to kMaxCand window positions from p_next on */
                int n_c = 0;
                for (int pb = p_next; pb < win && n_c < kMaxCand; pb += 32) {
                    const int p = pb + lane;
                    bool isc = p < win;
                    if (safe && isc) isc = cntw[p] > 0 && !(d32[p] > limd);   /* no column counts: NaN, never wins; NaN estimate: a candidate */
                    const unsigned m = __ballot_sync(0xffffffffu, isc);
                    const int pos = n_c + __popc(m & ((1u << lane) - 1u));
                    if (isc && pos < kMaxCand) list[pos] = p;
                    const int total = n_c + __popc(m);
                    if (total > kMaxCand) {
                        /* the list is full: resume after the last position taken */
                        int last = 0;
                        for (int b = 0, seen = n_c; b < 32; b++) if ((m >> b) & 1u) { if (seen < kMaxCand) last = pb + b; seen++; }
                        p_next = last + 1; n_c = kMaxCand;
                    } else { n_c = total; p_next = min(pb + 32, win); }
                }
                __syncwarp();
                if (n_c == 0) break;
                int my_cnt = 0;
                for (int k = 0; k < n_c; k++) {
                    int sh = s_first + list[k]; if (sh >= S) sh -= S;
                    int cnt = 0;
                    for (int c0 = 0; c0 < S; c0 += 64) {
                        /* lane <-> query columns c and c + 32, candidate columns c - sh (:1517-1519 on circshift(sc2, sh)); the two
                         * columns' dot products and divisions are independent chains */
                        const int ca = c0 + lane, cb = c0 + 32 + lane;
                        int ja = (ca < S ? ca : 0) - sh; if (ja < 0) ja += S;
                        int jb = (cb < S ? cb : 0) - sh; if (jb < 0) jb += S;
                        const double naa = ca < S ? nq[ca] : 0.0, nba = ca < S ? nc[ja] : 0.0;
                        const double nab = cb < S ? nq[cb] : 0.0, nbb = cb < S ? nc[jb] : 0.0;
                        const bool cnta = ca < S && !((naa == 0.0) | (nba == 0.0)), cntb = cb < S && !((nab == 0.0) | (nbb == 0.0));
                        double dota = 0.0, dotb = 0.0;
                        const int oa = ca < S ? ca : 0, ob = cb < S ? cb : 0;
                        const float* ka = cd + ja; const float* kb = cd + jb;
                        if (cnta | cntb) {
                            if (L.use_qdd) {
                                const double* qa = qdd + oa; const double* qb = qdd + ob;
#pragma unroll 4
                                for (int r = 0; r < R; r++) {      /* exact products: fused multiply-add == mul, then add */
                                    dota = __fma_rn(qa[r * S], (double)ka[r * S], dota);
                                    dotb = __fma_rn(qb[r * S], (double)kb[r * S], dotb);
                                }
                            } else {
                                const float* qa = qd + oa; const float* qb = qd + ob;
#pragma unroll 4
                                for (int r = 0; r < R; r++) {
                                    dota = __fma_rn((double)qa[r * S], (double)ka[r * S], dota);
                                    dotb = __fma_rn((double)qb[r * S], (double)kb[r * S], dotb);
                                }
                            }
                        }
                        const double sima = cnta ? __ddiv_rn(dota, __dmul_rn(naa, nba)) : 0.0;
                        const double simb = cntb ? __ddiv_rn(dotb, __dmul_rn(nab, nbb)) : 0.0;
                        if (ca < S) T[k * TP + ca] = sima;      /* columns that do not count hold +0.0: adding it changes nothing */
                        if (cb < S) T[k * TP + cb] = simb;
                        cnt += __popc(__ballot_sync(0xffffffffu, cnta)) + __popc(__ballot_sync(0xffffffffu, cntb));
                    }
                    if (lane == k) my_cnt = cnt;
                }
                __syncwarp();
                double dist = 0.0; int my_s = 0x7fffffff;
                if (lane < n_c) {
                    const double* sp = T + lane * TP;
                    double sum = 0.0;
#pragma unroll 4
                    for (int j = 0; j < S; j++) sum = __dadd_rn(sum, sp[j]);
                    dist = __dsub_rn(1.0, __ddiv_rn(sum, (double)my_cnt));  /* 0/0 = NaN when no column counts (:1534) */
                    my_s = s_first + list[lane]; if (my_s >= S) my_s -= S;
                }
                /* smallest (distance, shift) of the pass; NaN and anything not below the running minimum never wins */
                const bool cand = lane < n_c && dist < 10000000.0;
                double bd = cand ? dist : 10000000.0; int bs = cand ? my_s : 0x7fffffff;
#pragma unroll
                for (int off = 2; off > 0; off >>= 1) {
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
            if (a.cand_dist) a.cand_dist[(size_t)qi * K + it] = out_dist;
            if (a.cand_shift) a.cand_shift[(size_t)qi * K + it] = out_shift;
        }
        /* next candidate of this warp into the same tile */
        const int nxt = oi + slots < n_owned ? s_owned[oi + slots] : K;
        c_local = nxt < K ? s_local[nxt] : -1;
        it = nxt;
        __syncwarp();
        if (a.use_bulk && lane == 0 && c_local >= 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            scl_mbar_expect_tx(&bars[1 + warp], bytes + sbytes);
            scl_bulk_g2s(cd, a.db_desc + (size_t)c_local * RS, bytes, &bars[1 + warp]);
            scl_bulk_g2s(vc, a.db_stat + (size_t)c_local * 2 * S, sbytes, &bars[1 + warp]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        /* candidate scan in kNN order, strict <, query itself skipped (:1721-1737) */
        double min_dist = 10000000.0; int nn_align = 0, nn_idx = -1;
        const int self = a.q_ids ? a.q_ids[qi] : -1;
        for (int i = 0; i < K; i++) {
            const int id = a.cand_ids[(size_t)qi * K + i];
            if (id < 0) continue;
            if (res_dist[i] < min_dist && id != self) { min_dist = res_dist[i]; nn_align = res_shift[i]; nn_idx = id; }
        }
        if (a.best_id) a.best_id[qi] = nn_idx;
        if (a.best_dist) a.best_dist[qi] = min_dist;
        if (a.best_shift) a.best_shift[qi] = nn_align;
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

cudaError_t set_smem(void (*kern)(const ScArgs), size_t bytes)
{
    /* per kernel and device (a function attribute belongs to the current device's context), raised on demand */
    struct Seen { const void* k; int dev; size_t bytes; };
    static Seen seen[64]; static int n_seen = 0; static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    Seen* hit = nullptr;
    for (int i = 0; i < n_seen; i++) if (seen[i].k == (const void*)kern && seen[i].dev == dev) hit = &seen[i];
    if (hit && hit->bytes >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    if (hit) hit->bytes = bytes;
    else if (n_seen < 64) seen[n_seen++] = Seen{(const void*)kern, dev, bytes};
    return cudaSuccess;
}

} // namespace

cudaError_t scl_launch_scdist(const float* db_desc, const double* db_stat, const float* q_desc, const double* q_stat,
                              const int32_t* q_local, const int32_t* q_ids,
                              const int32_t* cand_local, const int32_t* cand_ids, int id_mul, int id_add, int Q, int K, int R, int S, int search_radius,
                              double* cand_dist, int32_t* cand_shift, int32_t* best_id, double* best_dist, int32_t* best_shift,
                              int owned_per_query /* expected candidates per query held here; <= 0: all K */, int exact_all, cudaStream_t stream)
{
    if (Q <= 0) return cudaSuccess;
    if (S > 128 || K > 32) return cudaErrorNotSupported;
    const size_t budget = 227 * 1024 - 2048;
    const bool shard = owned_per_query > 0;
    const int use_qdd = shard ? 0 : 1;
    int slots = K < 16 ? K : 16;
    /* a shard holds about K / world of a query's candidates: fewer candidate tiles per CTA then, so that more CTAs (queries) share an SM */
    if (shard && owned_per_query < slots) slots = owned_per_query < 2 ? 2 : owned_per_query;
    while (slots > 1 && sc_layout(R, S, K, slots, use_qdd).total > budget) slots--;
    const ScLayout L = sc_layout(R, S, K, slots, use_qdd);
    if (L.total > budget) return cudaErrorNotSupported;
    const int warps = slots < 4 ? 4 : slots;           /* at least four warps prepare the query side */
    const int use_bulk = ((R * S) % 4 == 0) && ((reinterpret_cast<uintptr_t>(db_desc) & 15) == 0) &&
                         (q_desc == nullptr || ((reinterpret_cast<uintptr_t>(q_desc) & 15) == 0 && (reinterpret_cast<uintptr_t>(q_stat) & 15) == 0)) &&
                         ((reinterpret_cast<uintptr_t>(db_stat) & 15) == 0);
    if (S & 1) exact_all = 1;                                   /* the FP32 window pass pairs adjacent columns */
    ScArgs a{L, db_desc, db_stat, q_desc, q_stat, q_local, q_ids, cand_local, cand_ids, id_mul < 1 ? 1 : id_mul, id_add, K, R, S, search_radius, use_bulk, exact_all,
             cand_dist, cand_shift, best_id, best_dist, best_shift};
#define SCL_SCDIST_LAUNCH(MT, MB, RT, ST)                                                                                       \
    do {                                                                                                                       \
        cudaError_t ea = set_smem(scdist_kernel<MT, MB, RT, ST>, L.total);                                                       \
        if (ea != cudaSuccess) return ea;                                                                                      \
        SCL_PREFER_SMEM((scdist_kernel<MT, MB, RT, ST>));                                                                       \
        scdist_kernel<MT, MB, RT, ST><<<Q, warps * 32, L.total, stream>>>(a);                                                   \
    } while (0)
    /* up to 10 warps and half an SM's shared memory: two CTAs per SM; otherwise one big CTA; a shard's small CTAs: six per SM */
    const bool two = warps <= 10 && L.total <= 113 * 1024;
    if (R == 20 && S == 60 && shard && warps == 4 && L.total <= 48 * 1024) SCL_SCDIST_LAUNCH(128, 6, 20, 60);
    else if (R == 20 && S == 60) { if (two) SCL_SCDIST_LAUNCH(320, 2, 20, 60); else SCL_SCDIST_LAUNCH(512, 1, 20, 60); }
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

void scl_preload_k4()
{
    SCL_TOUCH((scdist_kernel<320, 2, 20, 60>)); SCL_TOUCH((scdist_kernel<512, 1, 20, 60>)); SCL_TOUCH((scdist_kernel<128, 6, 20, 60>)); SCL_TOUCH((scdist_kernel<320, 2, 40, 120>));
    SCL_TOUCH((scdist_kernel<512, 1, 40, 120>)); SCL_TOUCH((scdist_kernel<320, 2, 0, 0>)); SCL_TOUCH((scdist_kernel<512, 1, 0, 0>));
    SCL_TOUCH(merge_shards_kernel); SCL_TOUCH(merge_topk_kernel); SCL_TOUCH(combine_owned_kernel);
}
