// alu_bench2.cu — warp-instructions per clock per SMSP of the min flavours the kNN epilogue could use, with every value
// passed through inline asm so that the compiler can neither fold nor re-fuse the chains.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define OPAQUE_I(x) asm volatile("" : "+r"(x))
template <int OP> __device__ __forceinline__ void step(float (&a)[8], int (&ia)[8], float b, float c, int ib, int ic)
{
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (OP == 0) { asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); }                    /* FMNMX */
        if (OP == 1) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c)); }        /* FMNMX3 */
        if (OP == 2) { asm volatile("min.s32 %0, %0, %1;" : "+r"(ia[i]) : "r"(ib)); }                  /* IMNMX / VIMNMX */
        if (OP == 3) { ia[i] = min(min(ia[i], ib), ic); OPAQUE_I(ia[i]); }                             /* VIMNMX3 if fused */
        if (OP == 4) { asm volatile("min.u32 %0, %0, %1;" : "+r"(ia[i]) : "r"(ib)); }
        if (OP == 5) { asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; @p add.s32 %2, %2, 1;}" : "+f"(a[i]), "+f"(b), "+r"(ia[i])); }   /* FSETP + predicated IADD */
        if (OP == 7) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c)); ia[i] = min(min(ia[i], ib), ic); OPAQUE_I(ia[i]); }   /* FMNMX3 + VIMNMX3 interleaved */
        if (OP == 6) { asm volatile("{.reg .pred p; setp.lt.s32 p, %0, %1; @p add.s32 %0, %0, 1;}" : "+r"(ia[i]) : "r"(ib)); }         /* ISETP + predicated IADD */
    }
}
template <int OP> __global__ void __launch_bounds__(1024) bench(int iters, float b, float c, int ib, int ic, float* out, long long* cyc)
{
    float a[8]; int ia[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x + i; ia[i] = threadIdx.x * 3 + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) { step<OP>(a, ia, b, c, ib, ic); step<OP>(a, ia, c, b, ic, ib); }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; i++) s += a[i] + ia[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, int threads, int per_step)
{
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int rep = 0; rep < 2; rep++) bench<OP><<<148, threads>>>(iters, 1.5f, 2.5f, 7, 9, out, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double winst = (double)iters * 16 * per_step * (threads / 32);
    printf("%-26s %4d threads/SM: %6.3f warp-instr/clk/SM (%5.2f per SMSP)\n", name, threads, winst / h, winst / h / 4);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int threads : {128, 512}) {
        if (threads == 128) { run<0>("min.f32 (2 in)", 128, 1); run<1>("min.f32 (3 in)", 128, 1); run<2>("min.s32 (2 in)", 128, 1); run<3>("min(min()) s32 (3 in)", 128, 1); run<4>("min.u32 (2 in)", 128, 1); run<5>("setp.f32 + @p add", 128, 2); run<6>("setp.s32 + @p add", 128, 2); run<7>("min3.f32 + min3.s32 mixed", 128, 2); }
        else { run<0>("min.f32 (2 in)", 512, 1); run<1>("min.f32 (3 in)", 512, 1); run<2>("min.s32 (2 in)", 512, 1); run<3>("min(min()) s32 (3 in)", 512, 1); run<4>("min.u32 (2 in)", 512, 1); run<5>("setp.f32 + @p add", 512, 2); run<6>("setp.s32 + @p add", 512, 2); run<7>("min3.f32 + min3.s32 mixed", 512, 2); }
    }
    return 0;
}
