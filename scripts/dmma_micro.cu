#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k(double *out, int iters, long long *clk) {
    double d[ILP][2];
    double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
    for (int i = 0; i < ILP; i++) d[i][0] = d[i][1] = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) dmma(d[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += d[i][0] + d[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
    double *out; long long *clk, h; cudaMalloc(&out, 1 << 22); cudaMalloc(&clk, 8);
    const int iters = 2048;
    for (int block : {32, 128, 256, 512}) {
        k<8><<<1, block>>>(out, iters, clk);
        cudaDeviceSynchronize(); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        const double macs = (double)(block / 32) * 8 * iters * 8 * 8 * 4;
        printf("dmma m8n8k4 block %3d: %lld clk -> %.1f DFMA/clk/SM, %.2f clk per mma per warp (%s)\n", block, h, macs / h, (double)h / (8 * iters), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
