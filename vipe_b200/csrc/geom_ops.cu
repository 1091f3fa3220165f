// The other four operators of the reference's `slam_ext` module ("next" rows 2-3):
// projmap, frame_distance, depth_filter, iproj  (csrc/slam_ext/geom_kernels.cu:434-861, hosts at :1406-1507;
// bound at csrc/slam_ext/slam.cpp:33-36).  All four are streaming, HBM-bound kernels around the same per-edge
// projective transform as the BA; they share its relative-pose code (ba_common.cuh).
//
// Differences in organisation from the reference (results identical up to fp32 rounding):
//   * one thread per pixel over a 2-D grid instead of one 256-thread CTA striding over a whole frame;
//   * depth_filter loops over its six neighbour frames inside the thread and writes the count once: no atomics
//     (the reference launches num x 6 x tiles CTAs that atomicAdd into the counter, :783-790);
//   * frame_distance reduces with warp shuffles, and the relative pose is computed once per CTA.
#include "../../include/vipe_ba.h"
#include <cstdio>

#include "ba_common.cuh"
#include "ba_launch.h"

namespace vba {

constexpr int GT = 256;

// Tij = Tj * Ti^-1 WITHOUT the stereo convention (these ops use relSE3 as is, e.g. geom_kernels.cu:484)
__device__ __forceinline__ void rel_pose_noconv(const float *__restrict__ poses, int i, int j, float *R, float *t) {
    float ti[3], qi[4], tj[3], qj[4], qij[4];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        ti[k] = poses[7 * i + k];
        tj[k] = poses[7 * j + k];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        qi[k] = poses[7 * i + 3 + k];
        qj[k] = poses[7 * j + 3 + k];
    }
    qij[0] = -qj[3] * qi[0] + qj[0] * qi[3] - qj[1] * qi[2] + qj[2] * qi[1];  // :105-108
    qij[1] = -qj[3] * qi[1] + qj[1] * qi[3] - qj[2] * qi[0] + qj[0] * qi[2];
    qij[2] = -qj[3] * qi[2] + qj[2] * qi[3] - qj[0] * qi[1] + qj[1] * qi[0];
    qij[3] = qj[3] * qi[3] + qj[0] * qi[0] + qj[1] * qi[1] + qj[2] * qi[2];
    quat_to_mat(qij, R);
#pragma unroll
    for (int k = 0; k < 3; k++) t[k] = tj[k] - (R[3 * k] * ti[0] + R[3 * k + 1] * ti[1] + R[3 * k + 2] * ti[2]);
}

__device__ __forceinline__ void act(const float *R, const float *t, float xn, float yn, float h, float &x, float &y, float &z) {
    x = fmaf(R[0], xn, fmaf(R[1], yn, fmaf(h, t[0], R[2])));
    y = fmaf(R[3], xn, fmaf(R[4], yn, fmaf(h, t[1], R[5])));
    z = fmaf(R[6], xn, fmaf(R[7], yn, fmaf(h, t[2], R[8])));
}

// ------------------------------------------------------------------------------------------------
// projmap (:434-519): coords[e][px] = (u', v', 0), valid[e][px] = z > MIN_DEPTH
__global__ void __launch_bounds__(GT) projmap_kernel(const float *__restrict__ poses, const float *__restrict__ disps,
                                                     const float *__restrict__ intr, const long long *__restrict__ ii,
                                                     const long long *__restrict__ jj, int HW, int wd,
                                                     float *__restrict__ coords, float *__restrict__ valid) {
    __shared__ float sR[9], st[3];
    const int e = blockIdx.y;
    const int ix = (int)ii[e], jx = (int)jj[e];
    if (threadIdx.x == 0) rel_pose_noconv(poses, ix, jx, sR, st);
    __syncthreads();
    const int px = blockIdx.x * GT + threadIdx.x;
    if (px >= HW) return;
    const float fx = __ldg(intr), fy = __ldg(intr + 1), cx = __ldg(intr + 2), cy = __ldg(intr + 3);
    const int row = px / wd, col = px - row * wd;
    const float u = (float)col, v = (float)row;
    const float h = __ldg(disps + (size_t)ix * HW + px);
    float x, y, z;
    act(sR, st, __fdiv_rn(u - cx, fx), __fdiv_rn(v - cy, fy), h, x, y, z);
    float cu = u, cv = v;
    if (z > 0.01f) {  // :512-515
        cu = fmaf(fx, __fdiv_rn(x, z), cx);
        cv = fmaf(fy, __fdiv_rn(y, z), cy);
    }
    float *c = coords + ((size_t)e * HW + px) * 3;
    c[0] = cu;
    c[1] = cv;
    c[2] = 0.0f;  // the reference allocates zeros and never writes the third channel (:1448)
    valid[(size_t)e * HW + px] = (z > kMinDepth) ? 1.0f : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// frame_distance (:521-676): mean induced flow, beta-blend of full motion and translation-only motion
__global__ void __launch_bounds__(GT) frame_distance_kernel(const float *__restrict__ poses, const float *__restrict__ disps,
                                                            const float *__restrict__ intr, const long long *__restrict__ pi,
                                                            const long long *__restrict__ pj, const long long *__restrict__ ri,
                                                            const long long *__restrict__ rj, const long long *__restrict__ di,
                                                            int HW, int wd, float beta, float *__restrict__ dist) {
    __shared__ float sR[9], st[3];
    __shared__ float red[3][GT / 32];
    const int b = blockIdx.x;
    const int ix = (int)pi[b], jx = (int)pj[b], dix = (int)di[b];
    const int rix = (int)ri[b], rjx = (int)rj[b];
    if (threadIdx.x == 0) rel_pose_noconv(poses, ix, jx, sR, st);
    __syncthreads();
    const float fxi = __ldg(intr + 4 * rix), fyi = __ldg(intr + 4 * rix + 1), cxi = __ldg(intr + 4 * rix + 2), cyi = __ldg(intr + 4 * rix + 3);
    const float fxj = __ldg(intr + 4 * rjx), fyj = __ldg(intr + 4 * rjx + 1), cxj = __ldg(intr + 4 * rjx + 2), cyj = __ldg(intr + 4 * rjx + 3);
    float accum = 0.0f, valid = 0.0f, total = 0.0f;
    for (int px = threadIdx.x; px < HW; px += GT) {
        const int row = px / wd, col = px - row * wd;
        const float u = (float)col, v = (float)row;
        const float xn = __fdiv_rn(u - cxi, fxi), yn = __fdiv_rn(v - cyi, fyi);
        const float h = __ldg(disps + (size_t)dix * HW + px);
        float x, y, z;
        act(sR, st, xn, yn, h, x, y, z);  // full motion (:622-633)
        float du = fmaf(fxj, __fdiv_rn(x, z), cxj) - u, dv = fmaf(fyj, __fdiv_rn(y, z), cyj) - v;
        float d = sqrtf(du * du + dv * dv);
        total += beta;
        if (z > kMinDepth) {
            accum += beta * d;
            valid += beta;
        }
        x = fmaf(h, st[0], xn);  // translation only (:640-653)
        y = fmaf(h, st[1], yn);
        z = fmaf(h, st[2], 1.0f);
        du = fmaf(fxj, __fdiv_rn(x, z), cxj) - u;
        dv = fmaf(fyj, __fdiv_rn(y, z), cyj) - v;
        d = sqrtf(du * du + dv * dv);
        total += (1.0f - beta);
        if (z > kMinDepth) {
            accum += (1.0f - beta) * d;
            valid += (1.0f - beta);
        }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        accum += __shfl_xor_sync(0xffffffffu, accum, s);
        valid += __shfl_xor_sync(0xffffffffu, valid, s);
        total += __shfl_xor_sync(0xffffffffu, total, s);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = accum;
        red[1][warp] = valid;
        red[2][warp] = total;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.0f, vv = 0.0f, tt = 0.0f;
        for (int w = 0; w < GT / 32; w++) {
            a += red[0][w];
            vv += red[1][w];
            tt += red[2][w];
        }
        dist[b] = (vv / (tt + 1e-8f) < 0.75f) ? 1000.0f : a / vv;  // :674
    }
}

// ------------------------------------------------------------------------------------------------
// depth_filter (:678-793): for every pixel of frame ix[b], how many of the six temporal neighbours
// (ix-3..ix-1, ix+1..ix+3) see a consistent inverse depth at the reprojected location
__global__ void __launch_bounds__(GT) depth_filter_kernel(const float *__restrict__ poses, const float *__restrict__ disps,
                                                          const float *__restrict__ intr, const long long *__restrict__ inds,
                                                          const float *__restrict__ thresh, int num, int HW, int ht, int wd,
                                                          float *__restrict__ counter) {
    __shared__ float sR[6][9], st[6][3];
    __shared__ int sj[6];
    const int b = blockIdx.y;
    const int ix = (int)inds[b];
    if (threadIdx.x < 6) {
        const int n = threadIdx.x;
        const int jx = (n < 3) ? ix - n - 1 : ix + n - 2;  // :709
        const bool ok = jx >= 0 && jx < num;               // :718
        sj[n] = ok ? jx : -1;
        if (ok) rel_pose_noconv(poses, ix, jx, sR[n], st[n]);
    }
    __syncthreads();
    const int px = blockIdx.x * GT + threadIdx.x;
    if (px >= HW) return;
    const float fx = __ldg(intr), fy = __ldg(intr + 1), cx = __ldg(intr + 2), cy = __ldg(intr + 3);
    const float t = __ldg(thresh + b);
    const int row = px / wd, col = px - row * wd;
    const float xn = __fdiv_rn((float)col - cx, fx), yn = __fdiv_rn((float)row - cy, fy);
    const float di = __ldg(disps + (size_t)ix * HW + px);
    float count = 0.0f;
#pragma unroll
    for (int n = 0; n < 6; n++) {
        const int jx = sj[n];
        if (jx < 0) continue;
        float x, y, z;
        act(sR[n], st[n], xn, yn, di, x, y, z);
        const float uj = fmaf(fx, __fdiv_rn(x, z), cx), vj = fmaf(fy, __fdiv_rn(y, z), cy);
        const float dj = __fdiv_rn(di, z);  // :766 (Xj[3] = Xi[3])
        const int u0 = (int)floorf(uj), v0 = (int)floorf(vj);
        if (u0 >= 0 && v0 >= 0 && u0 < wd - 1 && v0 < ht - 1) {
            const float *dp = disps + (size_t)jx * HW + v0 * wd + u0;
            const double idj = 1.0 / (double)dj;  // the reference compares in double (:782-789)
            const double tt = (double)t;
            if (fabs(idj - 1.0 / (double)__ldg(dp)) < tt || fabs(idj - 1.0 / (double)__ldg(dp + 1)) < tt ||
                fabs(idj - 1.0 / (double)__ldg(dp + wd)) < tt || fabs(idj - 1.0 / (double)__ldg(dp + wd + 1)) < tt)
                count += 1.0f;
        }
    }
    counter[(size_t)b * HW + px] = count;
}

// ------------------------------------------------------------------------------------------------
// iproj (:795-861): points = (R X + d t) / d with the frame's own pose
__global__ void __launch_bounds__(GT) iproj_kernel(const float *__restrict__ poses, const float *__restrict__ disps,
                                                   const float *__restrict__ intr, int HW, int wd, float *__restrict__ points) {
    __shared__ float sR[9], st[3];
    const int n = blockIdx.y;
    if (threadIdx.x == 0) {
        float q[4];
#pragma unroll
        for (int k = 0; k < 4; k++) q[k] = poses[7 * n + 3 + k];
        quat_to_mat(q, sR);
#pragma unroll
        for (int k = 0; k < 3; k++) st[k] = poses[7 * n + k];
    }
    __syncthreads();
    const int px = blockIdx.x * GT + threadIdx.x;
    if (px >= HW) return;
    const float fx = __ldg(intr), fy = __ldg(intr + 1), cx = __ldg(intr + 2), cy = __ldg(intr + 3);
    const int row = px / wd, col = px - row * wd;
    const float d = __ldg(disps + (size_t)n * HW + px);
    float x, y, z;
    act(sR, st, __fdiv_rn((float)col - cx, fx), __fdiv_rn((float)row - cy, fy), d, x, y, z);
    float *p = points + ((size_t)n * HW + px) * 3;
    p[0] = __fdiv_rn(x, d);
    p[1] = __fdiv_rn(y, d);
    p[2] = __fdiv_rn(z, d);
}

}  // namespace vba

using namespace vba;

static int geom_fail(const char *what, cudaError_t e);
#define GEOM_LAUNCH_CHECK(what)                           \
    do {                                                  \
        cudaError_t _e = cudaGetLastError();              \
        if (_e != cudaSuccess) return geom_fail(what, _e); \
    } while (0)

static int geom_fail(const char *what, cudaError_t e) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    return vba::set_last_error(buf);
}

extern "C" int vipe_projmap(const float *poses, const float *disps, const float *intrinsics, const int64_t *ii,
                            const int64_t *jj, int64_t n_edges, int ht, int wd, float *coords, float *valid, void *stream) {
    if (n_edges <= 0) return 0;
    const int HW = ht * wd;
    dim3 grid((HW + GT - 1) / GT, (unsigned)n_edges);
    projmap_kernel<<<grid, GT, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics, (const long long *)ii, (const long long *)jj,
                                                          HW, wd, coords, valid);
    GEOM_LAUNCH_CHECK("vipe_projmap");
    return 0;
}

extern "C" int vipe_frame_distance(const float *poses, const float *disps, const float *intrinsics, const int64_t *pi,
                                   const int64_t *pj, const int64_t *qi, const int64_t *qj, const int64_t *di,
                                   int64_t n_pairs, int ht, int wd, float beta, float *dist, void *stream) {
    if (n_pairs <= 0) return 0;
    frame_distance_kernel<<<(unsigned)n_pairs, GT, 0, (cudaStream_t)stream>>>(
        poses, disps, intrinsics, (const long long *)pi, (const long long *)pj, (const long long *)qi, (const long long *)qj,
        (const long long *)di, ht * wd, wd, beta, dist);
    GEOM_LAUNCH_CHECK("vipe_frame_distance");
    return 0;
}

extern "C" int vipe_depth_filter(const float *poses, const float *disps, const float *intrinsics, const int64_t *ix,
                                 const float *thresh, int64_t n_ix, int64_t n_frames, int ht, int wd, float *counter,
                                 void *stream) {
    if (n_ix <= 0) return 0;
    const int HW = ht * wd;
    dim3 grid((HW + GT - 1) / GT, (unsigned)n_ix);
    depth_filter_kernel<<<grid, GT, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics, (const long long *)ix, thresh,
                                                               (int)n_frames, HW, ht, wd, counter);
    GEOM_LAUNCH_CHECK("vipe_depth_filter");
    return 0;
}

extern "C" int vipe_iproj(const float *poses, const float *disps, const float *intrinsics, int64_t n_frames, int ht, int wd,
                          float *points, void *stream) {
    if (n_frames <= 0) return 0;
    const int HW = ht * wd;
    dim3 grid((HW + GT - 1) / GT, (unsigned)n_frames);
    iproj_kernel<<<grid, GT, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics, HW, wd, points);
    GEOM_LAUNCH_CHECK("vipe_iproj");
    return 0;
}
