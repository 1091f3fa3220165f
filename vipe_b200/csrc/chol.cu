// Damped dense Cholesky solve of the reduced camera system, entirely on the device, fp64.
//
// Replaces SparseBlock::solve (csrc/slam_ext/geom_kernels.cu:1172-1191): the reference copies the system to
// the host and runs Eigen::SimplicialLLT<double>; here the factorisation never leaves the GPU.
//   n <= kSmallMax : one CTA, matrix resident in shared memory (frontend windows, 6*P <= 160)
//   otherwise      : right-looking blocked factorisation, 64-wide panels
// A non-positive (or NaN) pivot raises *flag and the solve writes dx = 0, like the reference (:1186-1188).
#include "ba_launch.h"

namespace vba {

constexpr int kSmallMax = 160;
constexpr int NB = kCholBlock;  // 64
constexpr int LDS = NB + 1;

// ------------------------------------------------------------------------------------------------
__global__ void pad_identity_kernel(double *H, int n, int npad) {
    const int i = n + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npad) H[(size_t)i * npad + i] = 1.0;
}

cudaError_t launch_system_clear(double *H, double *b, int n, int npad, cudaStream_t st) {
    // H and b are contiguous ([npad*npad] then [npad])
    cudaError_t err = cudaMemsetAsync(H, 0, ((size_t)npad * npad + npad) * sizeof(double), st);
    if (err != cudaSuccess) return err;
    (void)b;
    if (npad > n) {
        pad_identity_kernel<<<1, 64, 0, st>>>(H, n, npad);
        return cudaGetLastError();
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// factor an m x m lower-triangular block held in shared memory (row stride lds); returns false on failure.
// All threads of the CTA must call it.
__device__ bool smem_potrf(double *A, int m, int lds) {
    const int tid = threadIdx.x, NT = blockDim.x;
    for (int j = 0; j < m; j++) {
        const double ajj = A[j * lds + j];
        if (!(ajj > 0.0)) return false;  // uniform: every thread reads the same value
        const double dj = sqrt(ajj);
        const double inv = 1.0 / dj;
        __syncthreads();
        for (int i = j + 1 + tid; i < m; i += NT) A[i * lds + j] *= inv;
        if (tid == 0) A[j * lds + j] = dj;
        __syncthreads();
        const int r = m - j - 1;
        for (int idx = tid; idx < r * r; idx += NT) {
            const int a = idx / r, c = idx - a * r;
            if (c <= a) {
                const int i = j + 1 + a, k = j + 1 + c;
                A[i * lds + k] -= A[i * lds + j] * A[k * lds + j];
            }
        }
        __syncthreads();
    }
    return true;
}

__global__ void __launch_bounds__(512) chol_small_kernel(const double *__restrict__ H, const double *__restrict__ b,
                                                        int n, int ld, float lm, float ep, float *__restrict__ dx,
                                                        int *flag) {
    extern __shared__ __align__(16) double sm[];
    const int lds = n | 1;
    double *A = sm;
    double *bs = A + (size_t)n * lds;
    const int tid = threadIdx.x, NT = blockDim.x;
    for (int idx = tid; idx < n * n; idx += NT) {
        const int i = idx / n, j = idx - i * n;
        if (j <= i) {
            double v = H[(size_t)i * ld + j];
            if (i == j) v += (double)ep + (double)lm * v;  // geom_kernels.cu:1176
            A[i * lds + j] = v;
        }
    }
    for (int i = tid; i < n; i += NT) bs[i] = b[i];
    __syncthreads();
    const bool ok = smem_potrf(A, n, lds);
    if (!ok) {
        for (int i = tid; i < n; i += NT) dx[i] = 0.0f;
        if (tid == 0) *flag = 1;
        return;
    }
    // forward substitution (column oriented)
    for (int j = 0; j < n; j++) {
        const double yj = bs[j] / A[j * lds + j];
        __syncthreads();
        if (tid == 0) bs[j] = yj;
        for (int i = j + 1 + tid; i < n; i += NT) bs[i] -= A[i * lds + j] * yj;
        __syncthreads();
    }
    // backward substitution with L^T
    for (int j = n - 1; j >= 0; j--) {
        const double xj = bs[j] / A[j * lds + j];
        __syncthreads();
        if (tid == 0) bs[j] = xj;
        for (int i = tid; i < j; i += NT) bs[i] -= A[j * lds + i] * xj;
        __syncthreads();
    }
    for (int i = tid; i < n; i += NT) dx[i] = (float)bs[i];
    if (tid == 0) *flag = 0;
}

// ------------------------------------------------------------------------------------------------
// blocked path
__global__ void chol_damp_kernel(double *H, int n, int ld, float lm, float ep, int *flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *flag = 0;
    if (i < n) {
        const double v = H[(size_t)i * ld + i];
        H[(size_t)i * ld + i] = v + (double)ep + (double)lm * v;
    }
}

// CTA 0 factors the diagonal block and writes it back; CTA c >= 1 re-factors it locally (identical
// arithmetic) and solves its 64-row slab of the panel: X = A_rk L_kk^-T.
__global__ void __launch_bounds__(256) chol_panel_kernel(double *H, int ld, int k0, int *flag) {
    extern __shared__ __align__(16) double dyn_sm[];
    double *D = dyn_sm;
    double *X = dyn_sm + NB * LDS;
    if (*flag) return;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx / NB, c = idx - r * NB;
        D[r * LDS + c] = (c <= r) ? H[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
    }
    const int r0 = k0 + blockIdx.x * NB;
    if (blockIdx.x > 0) {
        for (int idx = tid; idx < NB * NB; idx += 256) {
            const int r = idx / NB, c = idx - r * NB;
            X[r * LDS + c] = H[(size_t)(r0 + r) * ld + k0 + c];
        }
    }
    __syncthreads();
    const bool ok = smem_potrf(D, NB, LDS);
    if (!ok) {
        if (blockIdx.x == 0 && tid == 0) *flag = 1;
        return;
    }
    if (blockIdx.x == 0) {
        for (int idx = tid; idx < NB * NB; idx += 256) {
            const int r = idx / NB, c = idx - r * NB;
            if (c <= r) H[(size_t)(k0 + r) * ld + k0 + c] = D[r * LDS + c];
        }
        return;
    }
    // 4 threads per row: thread (row, part) accumulates the terms p = part, part+4, ... of the dot product
    const int row = tid >> 2, part = tid & 3;
    double *xr = X + row * LDS;
    for (int c = 0; c < NB; c++) {
        const double *dc = D + c * LDS;
        double s = 0.0;
        for (int p = part; p < c; p += 4) s += xr[p] * dc[p];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const double v = (xr[c] - s) / dc[c];
        __syncwarp();
        if (part == 0) xr[c] = v;
        __syncwarp();
    }
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx / NB, c = idx - r * NB;
        H[(size_t)(r0 + r) * ld + k0 + c] = X[r * LDS + c];
    }
}

// trailing update: for every lower tile (bi >= bj > kb): A_ij -= L_ik L_jk^T   (64x64x64, fp64 FMA)
__global__ void __launch_bounds__(256) chol_update_kernel(double *H, int ld, int k0, int nb_rem, const int *flag) {
    extern __shared__ __align__(16) double dyn_sm[];
    double *As = dyn_sm;             // [k][i]
    double *Bs = dyn_sm + NB * LDS;  // [k][j]
    if (*flag) return;
    // decode tile index -> (ti >= tj) within the remaining nb_rem x nb_rem block grid
    const int t = blockIdx.x;
    int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
    while (ti * (ti + 1) / 2 > t) ti--;
    const int tj = t - ti * (ti + 1) / 2;
    const int i0 = k0 + NB + ti * NB, j0 = k0 + NB + tj * NB;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx / NB, c = idx - r * NB;  // r: row within tile, c: k index (contiguous in memory)
        As[c * LDS + r] = H[(size_t)(i0 + r) * ld + k0 + c];
        Bs[c * LDS + r] = H[(size_t)(j0 + r) * ld + k0 + c];
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 8
    for (int k = 0; k < NB; k++) {
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; a++) av[a] = As[k * LDS + ty * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; b++) bv[b] = Bs[k * LDS + tx * 4 + b];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        double *row = H + (size_t)(i0 + ty * 4 + a) * ld + j0 + tx * 4;
#pragma unroll
        for (int b = 0; b < 4; b++) row[b] -= acc[a][b];
    }
}

// forward + backward substitution on the factor, one CTA, rhs resident in shared memory
__global__ void __launch_bounds__(1024) chol_solve_kernel(const double *__restrict__ L, const double *__restrict__ b,
                                                         int n, int npad, float *__restrict__ dx, const int *flag) {
    extern __shared__ __align__(16) double sm[];
    double *bs = sm;          // [npad]
    double *D = bs + npad;    // [NB][LDS]
    double *ys = D + NB * LDS;  // [NB]
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    if (*flag) {
        for (int i = tid; i < n; i += NT) dx[i] = 0.0f;
        return;
    }
    for (int i = tid; i < npad; i += NT) bs[i] = b[i];
    const int nblk = npad / NB;
    for (int kb = 0; kb < nblk; kb++) {
        const int k0 = kb * NB;
        __syncthreads();
        for (int idx = tid; idx < NB * NB; idx += NT) {
            const int r = idx / NB, c = idx - r * NB;
            D[r * LDS + c] = L[(size_t)(k0 + r) * npad + k0 + c];
        }
        __syncthreads();
        if (warp == 0) {
            for (int c = 0; c < NB; c++) {
                const double yc = bs[k0 + c] / D[c * LDS + c];
                __syncwarp();
                for (int r = lane; r < NB; r += 32) {
                    if (r > c) bs[k0 + r] -= D[r * LDS + c] * yc;
                    if (r == c) {
                        bs[k0 + r] = yc;
                        ys[c] = yc;
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = k0 + NB + tid; i < npad; i += NT) {
            const double *row = L + (size_t)i * npad + k0;
            double s = 0.0;
#pragma unroll 8
            for (int c = 0; c < NB; c++) s += row[c] * ys[c];
            bs[i] -= s;
        }
    }
    for (int kb = nblk - 1; kb >= 0; kb--) {
        const int k0 = kb * NB;
        __syncthreads();
        for (int idx = tid; idx < NB * NB; idx += NT) {
            const int r = idx / NB, c = idx - r * NB;
            D[r * LDS + c] = L[(size_t)(k0 + r) * npad + k0 + c];
        }
        __syncthreads();
        if (warp == 0) {
            for (int c = NB - 1; c >= 0; c--) {
                const double xc = bs[k0 + c] / D[c * LDS + c];
                __syncwarp();
                for (int r = lane; r < NB; r += 32) {
                    if (r < c) bs[k0 + r] -= D[c * LDS + r] * xc;
                    if (r == c) {
                        bs[k0 + r] = xc;
                        ys[c] = xc;
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = tid; i < k0; i += NT) {
            double s = 0.0;
#pragma unroll 8
            for (int c = 0; c < NB; c++) s += L[(size_t)(k0 + c) * npad + i] * ys[c];
            bs[i] -= s;
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += NT) dx[i] = (float)bs[i];
}

cudaError_t launch_damped_solve(double *H, double *b, int n, int npad, float lm, float ep, float *dx, int *flag,
                                cudaStream_t st, int *launches) {
    cudaError_t err;
    int cnt = 0;
    if (n <= kSmallMax) {
        const size_t sm = ((size_t)n * (n | 1) + n) * sizeof(double);
        err = cudaFuncSetAttribute(chol_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (err != cudaSuccess) return err;
        chol_small_kernel<<<1, 512, sm, st>>>(H, b, n, npad, lm, ep, dx, flag);
        cnt++;
    } else {
        chol_damp_kernel<<<(n + 255) / 256, 256, 0, st>>>(H, n, npad, lm, ep, flag);
        cnt++;
        const int nblk = npad / NB;
        const size_t sm2 = 2 * NB * LDS * sizeof(double);
        err = cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        if (err != cudaSuccess) return err;
        for (int kb = 0; kb < nblk; kb++) {
            const int k0 = kb * NB;
            const int rem = nblk - kb - 1;
            chol_panel_kernel<<<1 + rem, 256, sm2, st>>>(H, npad, k0, flag);
            cnt++;
            if (rem > 0) {
                chol_update_kernel<<<rem * (rem + 1) / 2, 256, sm2, st>>>(H, npad, k0, rem, flag);
                cnt++;
            }
        }
        const size_t sm = ((size_t)npad + NB * LDS + NB) * sizeof(double);
        err = cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (err != cudaSuccess) return err;
        chol_solve_kernel<<<1, 1024, sm, st>>>(H, b, n, npad, dx, flag);
        cnt++;
    }
    if (launches) *launches += cnt;
    return cudaGetLastError();
}

}  // namespace vba
