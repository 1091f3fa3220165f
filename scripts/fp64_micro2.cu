// does an fp64 outer-product register tile sustain the DFMA rate?  (no shared memory involved)
#include <cstdio>
#include <cuda_runtime.h>
template <int M, int N>
__global__ void outer_kernel(double *out, int iters, long long *clk) {
    double acc[M][N], a[M], b[N];
    for (int i = 0; i < M; i++) { a[i] = 1.0 + 1e-9 * (threadIdx.x + i); for (int j = 0; j < N; j++) acc[i][j] = i * j; }
    for (int j = 0; j < N; j++) b[j] = 1.0 - 1e-9 * (threadIdx.x + j);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < M; i++)
#pragma unroll
            for (int j = 0; j < N; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        // perturb operands so the compiler cannot hoist anything
#pragma unroll
        for (int i = 0; i < M; i++) a[i] += 1e-12;
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int M, int N>
__global__ void outer_lds_kernel(double *out, int iters, long long *clk) {
    __shared__ __align__(16) double As[32 * 66], Bs[32 * 66];
    for (int i = threadIdx.x; i < 32 * 66; i += blockDim.x) { As[i] = 1.0 + 1e-9 * i; Bs[i] = 1.0 - 1e-9 * i; }
    __syncthreads();
    double acc[M][N];
    for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) acc[i][j] = i * j;
    const int r0 = (threadIdx.x % (64 / M)) * M, c0 = ((threadIdx.x / (64 / M)) % (64 / N)) * N;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 4
        for (int k = 0; k < 64; k++) {
            double a[M], b[N];
#pragma unroll
            for (int q = 0; q < M / 2; q++) { double2 x = *reinterpret_cast<const double2 *>(As + (k & 31) * 66 + r0 + 2 * q); a[2 * q] = x.x; a[2 * q + 1] = x.y; }
#pragma unroll
            for (int q = 0; q < N / 2; q++) { double2 y = *reinterpret_cast<const double2 *>(Bs + (k & 31) * 66 + c0 + 2 * q); b[2 * q] = y.x; b[2 * q + 1] = y.y; }
#pragma unroll
            for (int i = 0; i < M; i++)
#pragma unroll
                for (int j = 0; j < N; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
    double *out; long long *clk, h; cudaMalloc(&out, 1 << 22); cudaMalloc(&clk, 8);
    auto rep = [&](const char *name, int block, double fmas_per_thread) {
        cudaDeviceSynchronize(); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        printf("%-40s block %3d: %9lld clk, %.1f DFMA/clk/SM  (%s)\n", name, block, h, fmas_per_thread * block / h, cudaGetErrorString(cudaGetLastError()));
    };
    const int it = 512;
    outer_kernel<4, 4><<<1, 256>>>(out, it, clk); rep("regs only 4x4, 8 warps", 256, 16.0 * it);
    outer_kernel<8, 4><<<1, 256>>>(out, it, clk); rep("regs only 8x4, 8 warps", 256, 32.0 * it);
    outer_kernel<8, 8><<<1, 256>>>(out, it, clk); rep("regs only 8x8, 8 warps", 256, 64.0 * it);
    outer_kernel<8, 8><<<1, 128>>>(out, it, clk); rep("regs only 8x8, 4 warps", 128, 64.0 * it);
    const int it2 = 16;
    outer_lds_kernel<4, 4><<<1, 256>>>(out, it2, clk); rep("smem operands 4x4, 8 warps", 256, 16.0 * 64 * it2);
    outer_lds_kernel<8, 4><<<1, 256>>>(out, it2, clk); rep("smem operands 8x4, 8 warps", 256, 32.0 * 64 * it2);
    outer_lds_kernel<8, 8><<<1, 256>>>(out, it2, clk); rep("smem operands 8x8, 8 warps", 256, 64.0 * 64 * it2);
    outer_lds_kernel<8, 8><<<1, 128>>>(out, it2, clk); rep("smem operands 8x8, 4 warps", 128, 64.0 * 64 * it2);
    outer_lds_kernel<8, 8><<<1, 64>>>(out, it2, clk); rep("smem operands 8x8, 2 warps", 64, 64.0 * 64 * it2);
    return 0;
}
