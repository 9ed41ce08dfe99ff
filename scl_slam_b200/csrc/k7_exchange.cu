// k7_exchange.cu — the multi-GPU exchange of the sharded loop query fused with its merge kernels, over NVLink peer
// memory instead of NCCL: DESIGN.md §7.
//
// The two-phase sharded query (include/scl_engine.h) has two exchange points: the per-shard (id, d2) lists before the
// global top-K (descriptor.h:1714-1716 evaluated on the unsharded key set), and the per-owner (dist, shift) results
// before the winner scan (:1721-1737). Each was an NCCL all-gather followed by a small kernel. Here every rank owns one
// exchange buffer that all peers have mapped (CUDA IPC); ONE kernel per exchange point
//   1. stores this rank's block straight into its slot of every peer's buffer (st.global over NVLink),
//   2. makes it visible (__threadfence_system) and raises this rank's flag in every peer's buffer to the step number
//      (last CTA of the grid, found by a ticket),
//   3. waits until the flags of all ranks have reached the step number, and
//   4. does the merge (global top-K by (d2, id) / owner pick + strict-< winner scan) on the now complete local buffer.
// Data slots are double-buffered by step parity: a rank can run at most one exchange point ahead of a peer, because
// it needs that peer's flag for the point in between (see the reuse argument in DESIGN.md §7). The grids are small
// (Q / 128 CTAs, at most 32) so all of their CTAs are resident while they wait; the wait is bounded (a lost peer becomes a
// trap, not a hung GPU). Every query lane of an engine (engine_internal.h) has its own region of the buffer, flags and
// step counter, so batches on different lanes exchange independently and their waits overlap other lanes' kernels.
#include "common.cuh"
#include "kernels.h"

namespace {

__device__ __forceinline__ int ld_acquire_sys(const int* p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// common prologue: publish `bytes` from `src` at byte offset `dst_off` of every rank's buffer, raise this rank's flag of
// exchange point `phase` in every rank's buffer, wait for the flags of all ranks
__device__ void publish_and_wait(const XchgView& x, int phase, int seq, const unsigned char* src, size_t bytes, size_t dst_off)
{
    const int nthreads = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n16 = bytes / 16;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    for (int r = 0; r < x.world; r++) {
        uint4* d4 = reinterpret_cast<uint4*>(x.peer[r] + dst_off);
        if (reinterpret_cast<const unsigned char*>(d4) == src) continue;          /* already in place (own slice of a gather) */
        for (size_t i = tid; i < n16; i += nthreads) d4[i] = s4[i];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    int* ticket = reinterpret_cast<int*>(x.peer[x.rank] + x.ticket_off) + phase;
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < x.world) st_release_sys(reinterpret_cast<int*>(x.peer[threadIdx.x] + x.flag_off) + phase * 16 + x.rank, seq);
        if (threadIdx.x == 0) *ticket = 0;
    }
    /* every CTA waits for every rank's flag of this exchange point */
    if (threadIdx.x < x.world) {
        const int* f = reinterpret_cast<const int*>(x.peer[x.rank] + x.flag_off) + phase * 16 + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < seq) {
            if (clock64() - t0 > (40ll << 30)) __trap();         /* ~20 s: a peer is gone */
            __nanosleep(40);
        }
    }
    __syncthreads();     /* the polling threads' acquire loads + this barrier order every thread's (volatile) reads of the slots after the flags */
}

// exchange point 0: all-gather of the step's query descriptors. Every rank brings the rows [row0, row0 + rows) of the batch
// (its share of the host upload) and stores them into the same rows of every rank's query area (parity seq & 1).
__global__ void __launch_bounds__(256) xchg_gather_queries_kernel(XchgView x, int seq, const unsigned char* __restrict__ my_rows, size_t row0_bytes, size_t bytes)
{
    publish_and_wait(x, 2, seq, my_rows, bytes, x.data_off[2] + (size_t)(seq & 1) * x.slot_bytes[2] + row0_bytes);
}

// exchange point 1 + global top-K by (d2, id): block = [ids i32 Q*K | d2 f32 Q*K]
__global__ void __launch_bounds__(128) xchg_merge_topk_kernel(XchgView x, int seq, int Q, int K, const unsigned char* __restrict__ my_block,
                                                               int32_t* __restrict__ out_ids, float* __restrict__ out_d2)
{
    const size_t QK = (size_t)Q * K;
    publish_and_wait(x, 0, seq, my_block, QK * 8, x.data_off[0] + ((size_t)(seq & 1) * x.world + x.rank) * x.slot_bytes[0]);
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    const unsigned char* base = x.peer[x.rank] + x.data_off[0] + (size_t)(seq & 1) * x.world * x.slot_bytes[0];
    int head[16];
    for (int w = 0; w < x.world; w++) head[w] = 0;
    for (int r = 0; r < K; r++) {
        int bw = -1; float bd = 0.f; int bi = 0;
        for (int w = 0; w < x.world; w++) {
            if (head[w] >= K) continue;
            const size_t o = (size_t)qi * K + head[w];
            const unsigned char* blk = base + (size_t)w * x.slot_bytes[0];
            const int id = __ldcv(reinterpret_cast<const int32_t*>(blk) + o);
            if (id < 0) { head[w] = K; continue; }
            const float d = __ldcv(reinterpret_cast<const float*>(blk + QK * 4) + o);
            if (bw < 0 || d < bd || (d == bd && id < bi)) { bw = w; bd = d; bi = id; }
        }
        if (bw >= 0) head[bw]++;
        out_ids[(size_t)qi * K + r] = bw >= 0 ? bi : -1;
        out_d2[(size_t)qi * K + r] = bw >= 0 ? bd : 3.402823466e+38f;
    }
}

// exchange point 2 + owner pick + winner scan (descriptor.h:1721-1737): block = [dist f64 Q*K | shift i32 Q*K]
__global__ void __launch_bounds__(128) xchg_combine_kernel(XchgView x, int seq, int Q, int K, const unsigned char* __restrict__ my_block,
                                                            const int32_t* __restrict__ q_ids, const int32_t* __restrict__ cand_ids,
                                                            double* __restrict__ out_dist, int32_t* __restrict__ out_shift,
                                                            int32_t* __restrict__ best_id, double* __restrict__ best_dist, int32_t* __restrict__ best_shift)
{
    const size_t QK = (size_t)Q * K;
    publish_and_wait(x, 1, seq, my_block, QK * 12, x.data_off[1] + ((size_t)(seq & 1) * x.world + x.rank) * x.slot_bytes[1]);
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    const unsigned char* base = x.peer[x.rank] + x.data_off[1] + (size_t)(seq & 1) * x.world * x.slot_bytes[1];
    double min_dist = 10000000.0; int nn_align = 0, nn_idx = -1;
    const int self = q_ids ? q_ids[qi] : -1;
    for (int r = 0; r < K; r++) {
        const size_t o = (size_t)qi * K + r;
        const int id = cand_ids[o];
        double dist = __longlong_as_double(0x7ff8000000000000LL); int shift = 0;
        if (id >= 0) {
            const unsigned char* blk = base + (size_t)(id % x.world) * x.slot_bytes[1];
            dist = __ldcv(reinterpret_cast<const double*>(blk) + o);
            shift = __ldcv(reinterpret_cast<const int32_t*>(blk + QK * 8) + o);
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

cudaError_t scl_launch_xchg_merge_topk(const XchgView& x, int seq, int Q, int K, const void* my_block, int32_t* out_ids, float* out_d2, cudaStream_t stream)
{
    if (Q <= 0) return cudaSuccess;
    SCL_PREFER_SMEM(xchg_merge_topk_kernel);
    xchg_merge_topk_kernel<<<(Q + 127) / 128, 128, 0, stream>>>(x, seq, Q, K, static_cast<const unsigned char*>(my_block), out_ids, out_d2);
    return cudaGetLastError();
}

cudaError_t scl_launch_xchg_combine(const XchgView& x, int seq, int Q, int K, const void* my_block, const int32_t* q_ids, const int32_t* cand_ids,
                                    double* out_dist, int32_t* out_shift, int32_t* best_id, double* best_dist, int32_t* best_shift, cudaStream_t stream)
{
    if (Q <= 0) return cudaSuccess;
    SCL_PREFER_SMEM(xchg_combine_kernel);
    xchg_combine_kernel<<<(Q + 127) / 128, 128, 0, stream>>>(x, seq, Q, K, static_cast<const unsigned char*>(my_block), q_ids, cand_ids,
                                                           out_dist, out_shift, best_id, best_dist, best_shift);
    return cudaGetLastError();
}

cudaError_t scl_launch_xchg_gather_queries(const XchgView& x, int seq, const void* my_rows, size_t row0_bytes, size_t bytes, cudaStream_t stream)
{
    if (bytes % 16 || row0_bytes % 16) return cudaErrorInvalidValue;
    int blocks = (int)((bytes / 16 + 255) / 256);
    if (blocks > 32) blocks = 32;                     /* every CTA must be resident while it waits: a handful */
    if (blocks < 1) blocks = 1;
    SCL_PREFER_SMEM(xchg_gather_queries_kernel);
    xchg_gather_queries_kernel<<<blocks, 256, 0, stream>>>(x, seq, static_cast<const unsigned char*>(my_rows), row0_bytes, bytes);
    return cudaGetLastError();
}

void scl_preload_k7()
{
    SCL_TOUCH(xchg_gather_queries_kernel); SCL_TOUCH(xchg_merge_topk_kernel); SCL_TOUCH(xchg_combine_kernel);
}
