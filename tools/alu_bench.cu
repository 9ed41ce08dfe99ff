// alu_bench.cu — microbenchmark: warp-instructions per clock per SM of the min/compare flavours the kNN epilogue could use.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int OP> __device__ __forceinline__ void step(float (&a)[8], const float b, const float c)
{
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (OP == 0) a[i] = fminf(a[i], b + i);                                                     /* FMNMX */
        if (OP == 1) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));          /* FMNMX3 */
        if (OP == 2) a[i] = fmaf(a[i], b, c);                                                        /* FFMA */
        if (OP == 3) { int x = __float_as_int(a[i]); x = min(x, __float_as_int(b) + i); a[i] = __int_as_float(x); }   /* IMNMX */
        if (OP == 4) { int x = __float_as_int(a[i]); x = min(min(x, __float_as_int(b) + i), __float_as_int(c) - i); a[i] = __int_as_float(x); } /* VIMNMX3 if the compiler fuses */
        if (OP == 5) { __half2 h = *reinterpret_cast<__half2*>(&a[i]); __half2 hb = *reinterpret_cast<const __half2*>(&b); h = __hmin2(h, hb); a[i] = *reinterpret_cast<float*>(&h); } /* HMNMX2 */
        if (OP == 6) a[i] = a[i] + b;                                                                /* FADD */
        if (OP == 7) { asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %0, %2, p;}" : "+f"(a[i]) : "f"(b), "f"(c)); }   /* FSETP + SEL */
    }
}
template <int OP> __global__ void __launch_bounds__(1024) bench(int iters, float b, float c, float* out, long long* cyc)
{
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) { step<OP>(a, b, c); step<OP>(a, c, b); }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, int threads)
{
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int rep = 0; rep < 2; rep++) bench<OP><<<148, threads>>>(iters, 1.5f, 2.5f, out, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double winst = (double)iters * 16 * (threads / 32);
    printf("%-14s %4d threads/SM: %6.3f warp-instr/clk/SM (%5.2f per SMSP)\n", name, threads, winst / h, winst / h / 4);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int threads : {256, 1024}) {
        if (threads == 256) {
            run<0>("FMNMX", 256); run<1>("FMNMX3", 256); run<2>("FFMA", 256); run<3>("IMNMX", 256); run<4>("VIMNMX3", 256);
            run<5>("HMNMX2", 256); run<6>("FADD", 256); run<7>("FSETP+SEL", 256);
        } else {
            run<0>("FMNMX", 1024); run<1>("FMNMX3", 1024); run<2>("FFMA", 1024); run<3>("IMNMX", 1024); run<4>("VIMNMX3", 1024);
            run<5>("HMNMX2", 1024); run<6>("FADD", 1024); run<7>("FSETP+SEL", 1024);
        }
    }
    return 0;
}
