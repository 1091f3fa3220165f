// Latency micro-benchmarks behind the tile solve design (single warp): dependent DMMA chain, shared-memory round trip
// through __syncwarp, and the C-fragment -> A-fragment conversion step of the tile solve.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/dmma_lat.cu -o scripts/dmma_lat
#include <cstdio>
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__global__ void k(double *out, long long *clk) {
    __shared__ double sm[512];
    const int lane = threadIdx.x & 31, nw = blockDim.x >> 5, warp = threadIdx.x >> 5;
    double acc[2] = {1.0 + lane, 2.0};
    double a = 1e-3 * lane, b = 1e-3;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) { dmma884(acc, a, b); }   // dependent through acc
    long long t1 = clock64();
    // dependent through the A operand (C -> A needs smem in reality; here just feeds acc[0] back as a)
#pragma unroll 1
    for (int i = 0; i < 64; i++) { dmma884(acc, acc[0] * 1e-9, b); }
    long long t2 = clock64();
    double v = lane;
    double *my = sm + warp * 64;
#pragma unroll 1
    for (int i = 0; i < 64; i++) { my[lane] = v; __syncwarp(); v += my[(lane + 1) & 31]; __syncwarp(); }
    long long t3 = clock64();
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll 1
    for (int i = 0; i < 32; i++) {
        *reinterpret_cast<double2 *>(my + fr * 8 + 2 * fk) = make_double2(acc[0], acc[1]);
        __syncwarp();
        double x0[2] = {0, 0};
        dmma884(x0, my[fr * 8 + fk] * 1e-9, b);
        acc[0] = x0[0]; acc[1] = x0[1];
        __syncwarp();
    }
    long long t4 = clock64();
    // 4 independent chains, 16 deep
    double c4[4][2] = {{1,2},{3,4},{5,6},{7,8}};
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
#pragma unroll
        for (int q = 0; q < 4; q++) dmma884(c4[q], a, b);
    }
    long long t5 = clock64();
    if (lane == 0) { clk[warp * 8 + 0] = (t1 - t0); clk[warp * 8 + 1] = (t2 - t1); clk[warp * 8 + 2] = (t3 - t2); clk[warp * 8 + 3] = (t4 - t3); clk[warp * 8 + 4] = t5 - t4; }
    out[threadIdx.x] = acc[0] + acc[1] + v + c4[0][0] + c4[1][1] + c4[2][0] + c4[3][1];
    (void)nw;
}
int main() {
    double *out; long long *clk;
    cudaMalloc(&out, 8192); cudaMalloc(&clk, 1024);
    for (int threads : {32, 128, 256}) {
        for (int r = 0; r < 2; r++) {
            k<<<1, threads>>>(out, clk);
            cudaDeviceSynchronize();
        }
        long long h[64];
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%3d threads (warp 0): dependent DMMA via acc %.1f clk | via operand %.1f clk | smem round trip (STS, syncwarp, LDS, DADD, syncwarp) %.1f clk | C->A conversion + DMMA %.1f clk | 4 chains x16: %.1f clk per level  (%s)\n",
               threads, h[0] / 64.0, h[1] / 64.0, h[2] / 64.0, h[3] / 32.0, h[4] / 16.0, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
