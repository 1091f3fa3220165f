// throughput of legacy mma.sync TF32 / BF16 on B200 (register operands only)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int ILP, bool BF>
__global__ void k(float *out, int iters, long long *clk) {
    float d[ILP][4];
    unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, threadIdx.x * 5};
    for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) d[i][j] = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) { if (BF) mma_bf16(d[i], a, b); else mma_tf32(d[i], a, b); }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
    float *out; long long *clk, h; cudaMalloc(&out, 1 << 22); cudaMalloc(&clk, 8);
    const int iters = 2048;
    for (int bf = 0; bf < 2; bf++) for (int block : {32, 128, 256, 512}) {
        if (bf) k<8, true><<<1, block>>>(out, iters, clk); else k<8, false><<<1, block>>>(out, iters, clk);
        cudaDeviceSynchronize(); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        const double macs = (double)(block / 32) * 8 * iters * 16 * 8 * (bf ? 16 : 8);
        printf("%s block %3d: %lld clk -> %.0f MAC/clk/SM, %.2f clk per mma per warp (%s)\n", bf ? "bf16 m16n8k16" : "tf32 m16n8k8 ", block, h, macs / h, (double)h / (8 * iters), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
