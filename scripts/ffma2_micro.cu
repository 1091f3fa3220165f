#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
template <int ILP, bool PACKED>
__global__ void k(float2* out, const float2* in, int iters, long long* clk) {
    float2 acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = in[threadIdx.x + i];
    float2 a = in[100 + threadIdx.x], b = in[200 + threadIdx.x];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (PACKED) acc[i] = ffma2(acc[i], a, b);
            else { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaf(acc[i].y, a.y, b.y); }
        }
    }
    long long t1 = clock64();
    float2 s = acc[0];
    for (int i = 1; i < ILP; i++) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
    float2 *out, *in; long long *clk, h; cudaMalloc(&out, 1 << 24); cudaMalloc(&in, 1 << 16); cudaMemset(in, 0, 1 << 16); cudaMalloc(&clk, 8);
    const int iters = 4096;
    for (int packed = 0; packed < 2; packed++) for (int block : {32, 128, 256, 512}) {
        if (packed) k<8, true><<<1, block>>>(out, in, iters, clk); else k<8, false><<<1, block>>>(out, in, iters, clk);
        cudaDeviceSynchronize(); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        printf("%s block %3d: %lld clk -> %.1f scalar FMA/clk/SM\n", packed ? "FFMA2" : "FFMA ", block, h, 2.0 * 8 * iters * block / h);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int packed = 0; packed < 2; packed++) {
        cudaEventRecord(e0);
        if (packed) k<8, true><<<592, 256>>>(out, in, iters, clk); else k<8, false><<<592, 256>>>(out, in, iters, clk);
        cudaEventRecord(e1); cudaDeviceSynchronize(); float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s full chip: %.2f TFMA/s (scalar FMAs)\n", packed ? "FFMA2" : "FFMA ", 2.0 * 8 * iters * 256 * 592 / (ms * 1e-3) / 1e12);
    }
    return 0;
}
