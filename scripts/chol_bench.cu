// Developer benchmark of the dense Cholesky solve (chol.cu) on a random SPD system, with per-tile phase stamps.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DVBA_CHOL_TRACE -Ivipe_b200/csrc scripts/chol_bench.cu -o scripts/chol_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <algorithm>
#include "../vipe_b200/csrc/chol.cu"

using namespace vba;


// timing-only variants of warp_potrf32 (results are wrong on purpose): where do the ~230 clk per pivot go?
template <int VARIANT>
__device__ __forceinline__ void potrf32_variant(double (&a)[32], double *colbuf, int lane) {
    double diag = 0.0;
#pragma unroll
    for (int c = 0; c < 32; c++)
        if (lane == c) diag = a[c];
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const double piv = (VARIANT == 3) ? diag : __shfl_sync(0xffffffffu, diag, j);  // 3: no shuffle
        double inv;
        if (VARIANT == 1) { asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(piv)); }   // 1: seed only
        else if (VARIANT == 2) inv = piv * 0.001;                                           // 2: no rsqrt at all
        else inv = fast_rsqrt(piv);
        double l = ((lane == j) ? piv : a[j]) * inv;
        if (lane < j) l = 0.0;
        a[j] = l;
        diag = fma(-l, l, diag);
        if (VARIANT == 4) {  // 4: no column exchange (use own value)
#pragma unroll
            for (int k = j + 1; k < 32; k++) a[k] = fma(-l, l, a[k]);
        } else {
            double *cb = colbuf + (j & 1) * 32;
            cb[lane] = l;
            __syncwarp();
#pragma unroll
            for (int k = j + 1; k < 32; k++) a[k] = fma(-l, cb[k], a[k]);
        }
    }
    __syncwarp();
}

// ---- micro timing of the tile primitives (single CTA), cycles via clock64
__global__ void __launch_bounds__(CT) micro_kernel(double *out, long long *clk) {
    extern __shared__ __align__(16) double sm[];
    double *As = sm, *Bs = sm + TB * RS, *col = Bs + TB * RS, *dinv = col + TB, *linv8 = dinv + TB, *tmpw = linv8 + 8 * 96;
    __shared__ int sh_ok;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto fill = [&]() {
        for (int idx = tid; idx < TB * TB; idx += CT) {
            const int r = idx >> 6, c = idx & 63;
            As[r * RS + c] = (r == c) ? 100.0 + r : 1.0 / (1 + abs(r - c));
            Bs[r * RS + c] = 0.5 + 0.001 * idx;
        }
        if (tid == 0) sh_ok = 1;
        __syncthreads();
    };
    fill();
    tile_potrf_mma(As, dinv, col, linv8, tmpw, &sh_ok);  // warm the instruction cache
    fill();
    long long t0 = clock64();
    tile_potrf_mma(As, dinv, col, linv8, tmpw, &sh_ok);
    long long t1 = clock64();
    // warp-level pieces alone
    double a[32];
    if (warp == 0) {
        double inv;
#pragma unroll 1
        for (int rep = 0; rep < 4; rep++) {
            for (int c = 0; c < 32; c++) a[c] = (lane == c) ? 100.0 + c + rep : 1.0 / (1 + abs(lane - c));
            const long long q0 = clock64();
            warp_potrf32(a, col, lane, inv);
            const long long q1 = clock64();
            if (lane == 0) clk[8 + rep] = q1 - q0;
        }
#define TIME_VARIANT(V)                                                                        \
        for (int rep = 0; rep < 3; rep++) {                                                    \
            for (int c = 0; c < 32; c++) a[c] = (lane == c) ? 100.0 + c + rep : 1.0 / (1 + abs(lane - c)); \
            long long q0 = clock64();                                                          \
            potrf32_variant<V>(a, col, lane);                                                  \
            long long q1 = clock64();                                                          \
            if (lane == 0 && rep == 2) clk[16 + V] = q1 - q0;                                  \
            out[lane] += a[lane & 31];                                                         \
        }
        TIME_VARIANT(0) TIME_VARIANT(1) TIME_VARIANT(2) TIME_VARIANT(3) TIME_VARIANT(4)
        out[lane] = a[lane & 31] + inv;
        // rolled loops (fit the instruction cache): (5) the pivot dependency chain alone, (6) the whole factorisation with
        // the block in shared memory, column-major Am[k][lane]: LDS + DFMA + STS per trailing entry
        {
            double diag = 100.0 + lane, a0 = 1.0 / (1 + lane);
            long long q0 = clock64();
#pragma unroll 1
            for (int j = 0; j < 32; j++) {
                const double piv = __shfl_sync(0xffffffffu, diag, j);
                const double iv = fast_rsqrt(piv);
                const double l = ((lane == j) ? piv : a0) * iv;
                diag = fma(-l, l, diag);
                a0 = fma(l, 1e-3, a0);
            }
            long long q1 = clock64();
            if (lane == 0) clk[21] = q1 - q0;
            out[lane] += diag + a0;
        }
        {
            double *Am = Bs;  // [32][33] column-major scratch
            for (int c = 0; c < 32; c++) Am[c * 33 + lane] = (lane == c) ? 100.0 + c : 1.0 / (1 + abs(lane - c));
            __syncwarp();
            double diag = Am[lane * 33 + lane];
            long long q0 = clock64();
#pragma unroll 1
            for (int j = 0; j < 32; j++) {
                const double piv = __shfl_sync(0xffffffffu, diag, j);
                const double iv = fast_rsqrt(piv);
                double l = ((lane == j) ? piv : Am[j * 33 + lane]) * iv;
                if (lane < j) l = 0.0;
                Am[j * 33 + lane] = l;
                diag = fma(-l, l, diag);
                col[lane] = l;
                __syncwarp();
#pragma unroll 4
                for (int k = j + 1; k < 32; k++) Am[k * 33 + lane] = fma(-l, col[k], Am[k * 33 + lane]);
                __syncwarp();
            }
            long long q1 = clock64();
            if (lane == 0) clk[22] = q1 - q0;
            out[lane] += Am[lane * 33 + (lane & 7)];
        }
    }
    __syncthreads();
    long long t2 = clock64();
    tile_trsm_mma(Bs, As, dinv, linv8, tmpw);
    __syncthreads();
    long long t3 = clock64();
    double acc[8][2];
    for (int r = 0; r < 8; r++) for (int c = 0; c < 2; c++) acc[r][c] = r + c;
    __syncthreads();
    long long t4 = clock64();
    tile_gemm_sub(acc, As, Bs);
    __syncthreads();
    long long t5 = clock64();
    if (tid == 0) {
        clk[0] = t1 - t0; clk[1] = t3 - t2; clk[2] = t5 - t4;
    }
    out[64 + tid] = acc[0][0] + acc[7][1] + Bs[tid];
}

int main(int argc, char **argv) {
    if (argc > 1 && atoi(argv[1]) == 0) {
        double *out; long long *clk;
        cudaMalloc(&out, 4096 * 8); cudaMalloc(&clk, 256);
        cudaMemset(out, 0, 4096 * 8);
        const size_t smb = (size_t)(2 * TB * RS + 2 * TB + 8 * 96 + 8 * 160) * sizeof(double);
        cudaFuncSetAttribute(micro_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
        for (int r = 0; r < 3; r++) {
            micro_kernel<<<1, CT, smb>>>(out, clk);
            cudaDeviceSynchronize();
            long long h[24];
            cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
            printf("warp_potrf32 repeated in a loop: %lld %lld %lld %lld\n", h[8], h[9], h[10], h[11]);
            printf("variants (warm): full %lld | rsqrt seed only %lld | no rsqrt %lld | no pivot shuffle %lld | no column exchange %lld\n", h[16], h[17], h[18], h[19], h[20]);
            printf("rolled loops: pivot chain alone %lld | whole potrf32 with the block in shared memory %lld\n", h[21], h[22]);
            printf("cycles: tile_potrf_mma (64x64) %lld | tile_trsm_mma %lld | tile_gemm_sub %lld   (%s)\n", h[0], h[1], h[2], cudaGetErrorString(cudaGetLastError()));
        }
        return 0;
    }
    const int n = argc > 1 ? atoi(argv[1]) : 1794;
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    const int npad = std::max(64, (n + 63) / 64 * 64);
    const int T = npad / 64;
    // SPD: banded + random, diagonally dominant-ish
    std::vector<double> A((size_t)npad * npad, 0.0), b(npad, 0.0);
    srand(1);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j <= i; j++) {
            double v = ((rand() % 2001) - 1000) / 1000.0;
            if (i - j > 72 && (rand() % 10)) v = 0;
            A[(size_t)i * npad + j] = v;
        }
        A[(size_t)i * npad + i] = 200.0 + (rand() % 100);
        b[i] = ((rand() % 2001) - 1000) / 1000.0;
    }
    for (int i = n; i < npad; i++) A[(size_t)i * npad + i] = 1.0;
    double *dH, *dsys, *ddinv;
    float *ddx;
    int *dscr;
    long long *dtrace;
    const size_t sysn = (size_t)npad * npad + npad;
    cudaMalloc(&dsys, sysn * 8);
    cudaMalloc(&dH, sysn * 8);
    cudaMalloc(&ddinv, ((size_t)npad + (size_t)npad * 64) * 8);
    cudaMalloc(&ddx, npad * 4);
    const size_t nscr = chol_scratch_ints(npad);
    cudaMalloc(&dscr, nscr * 4);
    cudaMemset(dscr, 0, nscr * 4);
    const int total = T * (T + 1) / 2 + T;
    cudaMalloc(&dtrace, (size_t)total * 8 * 8);
    cudaMemset(dtrace, 0, (size_t)total * 8 * 8);
    cudaMemcpyToSymbol(g_trace, &dtrace, sizeof(dtrace));
    std::vector<double> sys(sysn);
    std::copy(A.begin(), A.end(), sys.begin());
    std::copy(b.begin(), b.end(), sys.begin() + (size_t)npad * npad);
    cudaMemcpy(dsys, sys.data(), sysn * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    float best = 1e9;
    for (int r = 0; r < reps; r++) {
        cudaMemcpy(dH, dsys, sysn * 8, cudaMemcpyDeviceToDevice);
        cudaDeviceSynchronize();
        int cnt = 0;
        cudaEventRecord(e0);
        cudaError_t err = launch_damped_solve(dH, dH + (size_t)npad * npad, n, npad, 0.0f, 0.0f, ddx, dscr, ddinv, ddinv + npad, nullptr, nullptr, r + 1, nullptr, nullptr, 0, &cnt);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(err)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
        printf("rep %d: %.3f ms\n", r, ms);
    }
    // residual check
    std::vector<float> x(npad);
    cudaMemcpy(x.data(), ddx, n * 4, cudaMemcpyDeviceToHost);
    double rmax = 0, bmax = 0;
    for (int i = 0; i < n; i++) {
        double s = 0;
        for (int j = 0; j < n; j++) s += (j <= i ? A[(size_t)i * npad + j] : A[(size_t)j * npad + i]) * x[j];
        rmax = std::max(rmax, std::fabs(s - b[i]));
        bmax = std::max(bmax, std::fabs(b[i]));
    }
    printf("n=%d T=%d best %.3f ms  residual max %.3e (|b| max %.3e)\n", n, T, best, rmax, bmax);
    // trace analysis of the last rep
    std::vector<long long> tr((size_t)total * 8);
    cudaMemcpy(tr.data(), dtrace, tr.size() * 8, cudaMemcpyDeviceToHost);
    long long t0 = tr[0];
    double sum[8] = {0};
    int cntd = 0, cnto = 0;
    double dsum[8] = {0}, osum[8] = {0};
    int t = 0;
    for (int j = 0; j < T; j++) {
        for (int i = j; i <= T; i++, t++) {
            if (i == T) continue;
            long long *s = &tr[(size_t)t * 8];
            if (i == j) {
                cntd++;
                dsum[0] += (s[1] - s[0]); dsum[1] += (s[2] - s[1]); dsum[2] += (s[4] - s[2]); dsum[3] += (s[5] - s[4]); dsum[4] += (s[6] - s[5]);
                if (j < 6 || j == T - 1) printf("diag %2d: start %8.1f us  load %5.1f  kloop %6.1f (last step %5.1f)  potrf %5.1f  store %5.1f  publish %5.1f  end %8.1f\n", j, (s[0] - t0) / 1e3,
                    (s[1] - s[0]) / 1e3, (s[2] - s[1]) / 1e3, j ? (s[2] - s[7]) / 1e3 : 0.0, (s[4] - s[2]) / 1e3, (s[5] - s[4]) / 1e3, (s[6] - s[5]) / 1e3, (s[6] - t0) / 1e3);
            } else {
                cnto++;
                osum[0] += (s[1] - s[0]); osum[1] += (s[2] - s[1]); osum[2] += (s[3] - s[2]); osum[3] += (s[4] - s[3]); osum[4] += (s[5] - s[4]); osum[5] += (s[6] - s[5]);
                if (i == j + 1 && (j < 6)) printf("sub  %2d: start %8.1f us  load %5.1f  kloop %6.1f  waitLjj+load %5.1f  trsm %5.1f  store %5.1f  publish %5.1f  end %8.1f\n", j, (s[0] - t0) / 1e3,
                    (s[1] - s[0]) / 1e3, (s[2] - s[1]) / 1e3, (s[3] - s[2]) / 1e3, (s[4] - s[3]) / 1e3, (s[5] - s[4]) / 1e3, (s[6] - s[5]) / 1e3, (s[6] - t0) / 1e3);
            }
        }
    }
    (void)sum;
    printf("diag tiles avg (us): load %.2f kloop %.2f potrf %.2f store %.2f publish %.2f\n", dsum[0] / cntd / 1e3, dsum[1] / cntd / 1e3, dsum[2] / cntd / 1e3, dsum[3] / cntd / 1e3, dsum[4] / cntd / 1e3);
    if (cnto) printf("off  tiles avg (us): load %.2f kloop %.2f wait+loadLjj %.2f trsm %.2f store %.2f publish %.2f\n", osum[0] / cnto / 1e3, osum[1] / cnto / 1e3, osum[2] / cnto / 1e3, osum[3] / cnto / 1e3, osum[4] / cnto / 1e3, osum[5] / cnto / 1e3);
    return 0;
}
