// fp64_bench.cu — microbenchmark behind K4's design: dependent-chain latency and throughput of the FP64
// operations and conversions the exact evaluation uses (DADD, DFMA, DMUL, F2F.F64.F32, F2F.F32.F64, LDS.64 + DADD).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP, int ILP> __global__ void __launch_bounds__(1024) bench(int iters, double b, double c, double* out, long long* cyc)
{
    __shared__ double sh[1024];
    sh[threadIdx.x] = b + threadIdx.x;
    double a[ILP]; float f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x + i; f[i] = threadIdx.x + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) a[i] = __dadd_rn(a[i], b);
            if (OP == 1) a[i] = __fma_rn(a[i], b, c);
            if (OP == 2) a[i] = __dmul_rn(a[i], b);
            if (OP == 3) { f[i] = (float)((double)f[i]) + 1.0f; }                       /* F2F.F64.F32 + F2F.F32.F64 + FADD */
            if (OP == 4) a[i] = __dadd_rn(a[i], sh[(threadIdx.x + it + i) & 1023]);       /* LDS.64 feeding a DADD chain */
            if (OP == 5) f[i] = fmaf(f[i], 1.0001f, 0.5f);                                /* FFMA for reference */
            if (OP == 6) a[i] = __ddiv_rn(a[i], b);
            if (OP == 7) a[i] = __dsqrt_rn(a[i]) + b;
        }
    }
    const long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; i++) s += a[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP, int ILP> void run(const char* name, int threads)
{
    double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 1000;
    for (int rep = 0; rep < 2; rep++) bench<OP, ILP><<<148, threads>>>(iters, 1.0000001, 2.5, out, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-22s ILP %d, %4d threads/SM: %7.2f cycles per dependent step, %6.3f warp-ops/clk/SM\n", name, ILP, threads,
           (double)h / iters, (double)iters * ILP * (threads / 32) / h);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0, 1>("DADD", 32); run<0, 1>("DADD", 128); run<0, 4>("DADD", 128); run<0, 4>("DADD", 1024);
    run<1, 1>("DFMA", 32); run<1, 1>("DFMA", 128); run<1, 4>("DFMA", 128); run<1, 4>("DFMA", 1024);
    run<2, 1>("DMUL", 32); run<2, 4>("DMUL", 1024);
    run<3, 1>("F2F 32->64->32 + FADD", 32); run<3, 4>("F2F 32->64->32 + FADD", 128); run<3, 4>("F2F 32->64->32 + FADD", 1024);
    run<4, 1>("LDS.64 + DADD", 32); run<4, 4>("LDS.64 + DADD", 128); run<4, 4>("LDS.64 + DADD", 1024);
    run<5, 1>("FFMA", 32); run<5, 4>("FFMA", 1024);
    run<6, 1>("DDIV", 32); run<6, 4>("DDIV", 1024);
    run<7, 1>("DSQRT + DADD", 32); run<7, 4>("DSQRT + DADD", 1024);
    return 0;
}
