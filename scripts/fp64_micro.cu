// fp64 pipe micro-benchmarks on B200: throughput per SM, single-warp issue rate, dependent latency.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dfma_kernel(double *out, int iters, long long *clk) {
    double acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
    const double a = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int ILP>
__global__ void ffma_kernel(float *out, int iters, long long *clk) {
    float acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3f + i;
    const float a = 1.0000001f, b = 1e-9f;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fmaf(acc[i], a, b);
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
    double *out; long long *clk; cudaMalloc(&out, 1 << 24); cudaMalloc(&clk, 8);
    long long h;
    const int iters = 4096;
    auto run = [&](const char *name, auto kern, int grid, int block, int ilp) {
        kern<<<grid, block>>>((decltype(out))out, iters, clk); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); kern<<<grid, block>>>((decltype(out))out, iters, clk); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        double fmas = (double)grid * block * ilp * iters;
        printf("%-34s grid %4d block %4d ilp %2d: %8lld clk  -> %.2f clk per warp-instr (per warp), %.1f FMA/clk/SM (block0), %.2f TFMA/s total\n", name, grid, block, ilp, h,
               (double)h / (iters * ilp), (double)block * ilp * iters / h * (grid >= 148 ? (double)((grid + 147) / 148) : 1.0), fmas / (ms * 1e-3) / 1e12);
    };
    run("dfma 1 warp dependent (ilp1)", dfma_kernel<1>, 1, 32, 1);
    run("dfma 1 warp ilp8", dfma_kernel<8>, 1, 32, 8);
    run("dfma 1 warp ilp16", dfma_kernel<16>, 1, 32, 16);
    run("dfma 4 warps ilp8", dfma_kernel<8>, 1, 128, 8);
    run("dfma 8 warps ilp8", dfma_kernel<8>, 1, 256, 8);
    run("dfma 16 warps ilp8", dfma_kernel<8>, 1, 512, 8);
    run("dfma full chip 148x2x256 ilp8", dfma_kernel<8>, 296, 256, 8);
    auto runf = [&](const char *name, auto kern, int grid, int block, int ilp) {
        kern<<<grid, block>>>((float *)out, iters, clk); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); kern<<<grid, block>>>((float *)out, iters, clk); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        double fmas = (double)grid * block * ilp * iters;
        printf("%-34s grid %4d block %4d ilp %2d: %8lld clk  -> %.2f clk per warp-instr (per warp), %.2f TFMA/s total\n", name, grid, block, ilp, h, (double)h / (iters * ilp), fmas / (ms * 1e-3) / 1e12);
    };
    runf("ffma 1 warp dependent", ffma_kernel<1>, 1, 32, 1);
    runf("ffma 1 warp ilp8", ffma_kernel<8>, 1, 32, 8);
    runf("ffma 8 warps ilp8", ffma_kernel<8>, 1, 256, 8);
    runf("ffma full chip 148x4x256 ilp8", ffma_kernel<8>, 592, 256, 8);
    return 0;
}
