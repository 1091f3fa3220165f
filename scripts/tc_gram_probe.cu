// Probe for the tensor-core Schur Gram (tcgen05.mma kind::tf32, error-compensated hi/lo operands, TMEM accumulators):
//   1. are the hand-built shared-memory / instruction descriptors right (D compared element-wise with fp64 products of
//      the very hi/lo values the kernel staged)?
//   2. how accurate is Y = hh + lh + lh^T against the exact Gram, as a function of the in-unit accumulation chain?
//   3. what does one K = 8 step cost for the (M, N) shapes the linearise kernel uses?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_gram_probe scripts/tc_gram_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../vipe_b200/csrc/sm100_async.cuh"

using namespace vba::sm100;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

// X: [R][K] fp32 row-major.  stacked: R <= 64, operand rows = [hi(64); lo(64)], D = 128 x N with N = R16.
// unstacked: R <= 128, planes h[128], l[128]; D = [h h^T | l h^T] = 128 x 2N.
// TILE pixels (K) per sub-tile; `chain` sub-tiles accumulate inside the tensor core before a flush (fp32 RN add outside).
// hl_out: the staged hi/lo values [2][R][K] for the exact check.  dacc: [128][2*N] fp32 accumulated outside.
__global__ void __launch_bounds__(128) probe_kernel(const float *X, int R, int K, int stacked, int TILE, int chain, int N,
                                                    float *hl_out, float *dacc, long long *cycles, int time_reps) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int natom = TILE / 32;                      // 32 tf32 per 128-byte swizzle row
    const int plane_rows = 128;                       // rows per A operand
    const uint32_t atom_bytes = plane_rows * 128;     // one [128 x 32] block
    const uint32_t plane_bytes = atom_bytes * natom;  // one operand plane for a sub-tile
    // stacked: one plane (rows 0..63 hi, 64..127 lo); unstacked: plane 0 = hi, plane 1 = lo
    const int nsub = K / TILE;
    uint32_t parity = 0;
    const int ncols = stacked ? N : 2 * N;
    for (int i = tid; i < 128 * ncols; i += 128) dacc[i] = 0.0f;
    __syncthreads();
    const uint32_t idesc = make_idesc_tf32(128, N);
    long long t_total = 0;
    for (int s = 0; s < nsub; s++) {
        // stage the sub-tile
        for (int idx = tid; idx < 128 * TILE; idx += 128) {
            const int r = idx / TILE, k = idx - r * TILE;
            float hi = 0.0f, lo = 0.0f;
            if (r < R) {
                split_tf32(X[(size_t)r * K + s * TILE + k], hi, lo);
                hl_out[((size_t)0 * R + r) * K + s * TILE + k] = hi;
                hl_out[((size_t)1 * R + r) * K + s * TILE + k] = lo;
            }
            const int atom = k >> 5, kk = k & 31;
            if (stacked) {
                if (r < 64) {
                    *(float *)(smem + atom * atom_bytes + sw128_offset(r, kk)) = hi;
                    *(float *)(smem + atom * atom_bytes + sw128_offset(r + 64, kk)) = lo;
                }
            } else {
                *(float *)(smem + atom * atom_bytes + sw128_offset(r, kk)) = hi;
                *(float *)(smem + plane_bytes + atom * atom_bytes + sw128_offset(r, kk)) = lo;
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const bool fresh = (s % chain) == 0;
            const long long t0 = clock64();
            for (int rep = 0; rep < (time_reps > 0 ? time_reps : 1); rep++) {
                for (int ks = 0; ks < TILE / 8; ks++) {
                    const uint32_t off = (ks >> 2) * atom_bytes + (ks & 3) * 32;
                    const uint64_t adesc = make_desc_k_sw128(smem_u32(smem) + off);
                    const uint64_t bdesc = adesc;  // B = the hi rows [0, N) of the same block
                    const bool acc = !(fresh && ks == 0 && rep == 0);
                    mma_tf32(tmem, adesc, bdesc, idesc, acc);
                    if (!stacked) {
                        const uint64_t a2 = make_desc_k_sw128(smem_u32(smem) + plane_bytes + off);
                        mma_tf32(tmem + N, a2, bdesc, idesc, acc);
                    }
                }
            }
            mma_commit(&bar);
            mbar_wait(&bar, parity);
            t_total += clock64() - t0;
        }
        __syncthreads();  // everybody else waits for thread 0, which waited for the MMAs
        parity ^= 1;
        tc_fence_after();
        if ((s % chain) == chain - 1 || s == nsub - 1) {
            // flush: thread (warp, lane) owns TMEM lane 32*warp + lane
            const int row = 32 * warp + lane;
            for (int c0 = 0; c0 < ncols; c0 += 16) {
                float v[16];
                tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) dacc[(size_t)row * ncols + c0 + i] += v[i];
            }
            tc_fence_before();
        }
        __syncthreads();
    }
    if (tid == 0) *cycles = t_total;
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

static double frand() { return (double)rand() / RAND_MAX; }

static void run_case(const char *name, int R, int K, int stacked, int TILE, int chain, bool positive, int time_reps = 0) {
    const int N = ((R + 15) / 16) * 16;
    std::vector<float> X((size_t)R * K);
    for (auto &v : X) {
        const double u = frand();
        v = (float)(positive ? (0.25 + u) : (2.0 * u - 1.0)) * (float)std::exp(3.0 * (frand() - 0.5));
    }
    float *dX, *dhl, *dacc;
    long long *dcyc;
    const int ncols = stacked ? N : 2 * N;
    CK(cudaMalloc(&dX, X.size() * 4));
    CK(cudaMalloc(&dhl, X.size() * 8));
    CK(cudaMalloc(&dacc, 128 * ncols * 4));
    CK(cudaMalloc(&dcyc, 8));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dhl, 0, X.size() * 8));
    const size_t smem = (size_t)(stacked ? 1 : 2) * 128 * 128 * (TILE / 32) + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(dX, R, K, stacked, TILE, chain, N, dhl, dacc, dcyc, time_reps);
    CK(cudaDeviceSynchronize());
    std::vector<float> hl(X.size() * 2), D((size_t)128 * ncols);
    long long cyc = 0;
    CK(cudaMemcpy(hl.data(), dhl, hl.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D.data(), dacc, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    if (time_reps > 0) {
        const double nmma = (double)(K / TILE) * time_reps * (TILE / 8) * (stacked ? 1 : 2);
        printf("%-34s M=128 N=%3d: %.1f cycles per K=8 step (%s), %lld cycles total\n", name, N, cyc / nmma * (stacked ? 1 : 2),
               stacked ? "1 MMA" : "2 MMAs", cyc);
    } else {
        // element-wise check of hh and lh against fp64 products of the staged values; then Y vs the exact Gram
        const float *hi = hl.data(), *lo = hl.data() + (size_t)R * K;
        double max_el = 0.0, ynum = 0.0, yden = 0.0, bias = 0.0, diag_bias = 0.0;
        for (int i = 0; i < R; i++)
            for (int j = 0; j < R; j++) {
                double hh = 0, lh = 0, hlv = 0, ex = 0, scale = 0;
                for (int k = 0; k < K; k++) {
                    hh += (double)hi[(size_t)i * K + k] * hi[(size_t)j * K + k];
                    lh += (double)lo[(size_t)i * K + k] * hi[(size_t)j * K + k];
                    hlv += (double)hi[(size_t)i * K + k] * lo[(size_t)j * K + k];
                    ex += (double)X[(size_t)i * K + k] * X[(size_t)j * K + k];
                    scale += std::fabs((double)X[(size_t)i * K + k] * X[(size_t)j * K + k]);
                }
                double dhh, dlh, dlhT;
                if (stacked) {
                    dhh = D[(size_t)i * ncols + j];
                    dlh = D[(size_t)(64 + i) * ncols + j];
                    dlhT = D[(size_t)(64 + j) * ncols + i];
                } else {
                    dhh = D[(size_t)i * ncols + j];
                    dlh = D[(size_t)i * ncols + N + j];
                    dlhT = D[(size_t)j * ncols + N + i];
                }
                max_el = std::fmax(max_el, std::fabs(dhh - hh) / scale);
                max_el = std::fmax(max_el, std::fabs(dlh - lh) / scale);
                const double y = dhh + dlh + dlhT;
                ynum += (y - ex) * (y - ex);
                yden += ex * ex;
                bias += (y - ex) / scale;
                if (i == j) diag_bias += (y - ex) / ex;
            }
        printf("%-34s R=%3d K=%5d chain=%2d: max |D-ref|/sum|.| = %.2e   Y rel Frobenius err = %.2e   mean signed err = %+.2e   "
               "mean signed diag err = %+.2e\n",
               name, R, K, chain, max_el, std::sqrt(ynum / yden), bias / (R * R), diag_bias / R);
    }
    cudaFree(dX);
    cudaFree(dhl);
    cudaFree(dacc);
    cudaFree(dcyc);
}

int main() {
    srand(7);
    run_case("stacked  signed  flush/tile", 61, 512, 1, 128, 1, false);
    run_case("stacked  positive flush/tile", 61, 512, 1, 128, 1, true);
    run_case("stacked  positive chain 4", 61, 512, 1, 128, 4, true);
    run_case("stacked  positive chain 24", 61, 3072, 1, 128, 24, true);
    run_case("stacked  positive flush/tile 3072", 61, 3072, 1, 128, 1, true);
    run_case("stacked  N=48", 43, 512, 1, 128, 1, false);
    run_case("unstacked signed flush/tile", 127, 512, 0, 64, 1, false);
    run_case("unstacked positive chain 8", 127, 512, 0, 64, 8, true);
    run_case("unstacked N=80", 73, 512, 0, 64, 1, false);
    run_case("time stacked N=64", 61, 512, 1, 128, 4, false, 16);
    run_case("time stacked N=48", 43, 512, 1, 128, 4, false, 16);
    run_case("time unstacked N=128", 127, 512, 0, 64, 8, false, 16);
    run_case("time unstacked N=80", 73, 512, 0, 64, 8, false, 16);
    return 0;
}
