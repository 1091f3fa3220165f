// The per-(edge, pixel pair) Jacobian math shared by the Blackwell linearisation pipelines (ba_lin3.cu, ba_lin4.cu):
// projective transform, residuals, J_j, J_z in normalised image coordinates with sign-stripped rows, packed FFMA2.
// Reference: projective_transform_kernel, csrc/slam_ext/geom_kernels.cu:284-387.
#pragma once

#include "ba_common.cuh"
#include "sm100_async.cuh"

namespace vba {
namespace jmath {
using namespace sm100;

constexpr int kSubTile = 64;  // pixels one warp handles per step (one pixel pair per lane); row stride of the raw staging rows

__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ float2 splat2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float rcp_apx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// one elected arrival per warp: the warp's earlier shared-memory writes are ordered before it by the warp barrier
__device__ __forceinline__ void warp_arrive(uint64_t *bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// The J warps accumulate sign-stripped quantities (see the J step): with row signs s = (+,+,-,-,+,-), entry (i,j) of H_jj
// carries s_i s_j, v_i carries s_i, and the staged u_hat = s * u.
__device__ __forceinline__ float hsign(int slot) {
    // slot order: (0,0) (1,1) (2,0) (2,1) (2,2) (3,0) (3,1) (3,2) (3,3) (4,0) (4,1) (4,2) (4,3) (4,4) (5,0) (5,1) (5,2) (5,3) (5,4) (5,5), v0..v5, energy
    const unsigned neg = (1u << 2) | (1u << 3) | (1u << 5) | (1u << 6) | (1u << 11) | (1u << 12) | (1u << 14) | (1u << 15) | (1u << 18) |
                         (1u << 22) | (1u << 23) | (1u << 25);
    return ((neg >> slot) & 1u) ? -1.0f : 1.0f;
}
__device__ __forceinline__ float usign(int r) { return (r == 2 || r == 3 || r == 5) ? -1.0f : 1.0f; }

// ---------------------------------------------------------------------------------------------------------------- J step
// Per-kernel constants of the J math, in normalised image coordinates: weights w' = 0.001 w fx^2 (geom_kernels.cu:304-305),
// residuals r' = r / fx, so that every J_j row loses its fx / fy factor and the sums come out in image units again.
struct JConst {
    float2 ifx2, ify2, ncx2, ncy2, m1, one2;
    float ifx, ify, ncx, ncy, wsx, wsy, md, wdf;
    bool strict;
};
__device__ __forceinline__ JConst make_jconst(const LinArgs &a) {
    const float fx = __ldg(a.intr + 0), fy = __ldg(a.intr + 1), cx = __ldg(a.intr + 2), cy = __ldg(a.intr + 3);
    JConst k;
    k.ifx = __fdiv_rn(1.0f, fx), k.ify = __fdiv_rn(1.0f, fy);
    k.ncx = -cx * k.ifx, k.ncy = -cy * k.ify;
    k.ifx2 = splat2(k.ifx), k.ify2 = splat2(k.ify), k.ncx2 = splat2(k.ncx), k.ncy2 = splat2(k.ncy);
    k.m1 = splat2(-1.0f), k.one2 = splat2(1.0f);
    k.wsx = kWeightScale * fx * fx, k.wsy = kWeightScale * fy * fy;
    k.md = a.opt.min_depth, k.strict = a.opt.depth_strict != 0;
    k.wdf = (float)a.tb.wd;
    return k;
}
// image coordinates of a lane's pixel pair, kept as floats (exact for these small integers)
struct PxPos {
    float col, row;
    __device__ __forceinline__ void set(int px, int wd) {
        const int r = px / wd;
        row = (float)r, col = (float)(px - r * wd);
    }
    __device__ __forceinline__ void advance(float n, float wdf) {
        col += n;
        while (col >= wdf) col -= wdf, row += 1.0f;
    }
    // normalised coordinates (col - cx) / fx, (row - cy) / fy (geom_kernels.cu:289-290) of the pair (px, px + 1)
    __device__ __forceinline__ void normalised(const JConst &k, float2 &xn, float2 &yn) const {
        float c1 = col + 1.0f, r1 = row;
        if (c1 >= k.wdf) c1 -= k.wdf, r1 += 1.0f;
        xn = make_float2(fmaf(col, k.ifx, k.ncx), fmaf(c1, k.ifx, k.ncx));
        yn = make_float2(fmaf(row, k.ify, k.ncy), fmaf(r1, k.ify, k.ncy));
    }
};

// One pixel pair of one edge: projective transform, residuals, J_j, J_z (projective_transform_kernel, :284-387), accumulated
// sign-stripped into hh (20 unique entries of H_jj, 6 of v_j, energy).  FULL: also the edge's C, w contributions and the
// pose-disparity vector u_hat -> rawp[row * kSubTile] (rows 0..5 u_hat, 6 C, 7 w).
//   A' = (a, 0, -P2, -P3, P4, -P5),  B' = (0, a, -Q2, -Q3, Q4, -Q5),  Q5 = -X;  u_hat = (u0, u1, -u2, -u3, u4, -u5).
template <bool FULL>
__device__ __forceinline__ void j_step(const JConst &k, const float2 *__restrict__ c, bool stereo, float2 xn, float2 yn, float2 h, float2 tu,
                                       float2 tv, float2 wu, float2 wv, float2 (&hh)[kEdgeVals], float *rawp) {
    const float2 t0 = c[9], t1 = c[10], t2 = c[11];
    // X_j = R X_i + h t (actSE3, :295), X_i = (xn, yn, 1, h)
    const float2 x = ffma2(c[0], xn, ffma2(c[1], yn, ffma2(h, t0, c[2])));
    const float2 y = ffma2(c[3], xn, ffma2(c[4], yn, ffma2(h, t1, c[5])));
    const float2 z = ffma2(c[6], xn, ffma2(c[7], yn, ffma2(h, t2, c[8])));
    const bool v0 = k.strict ? (z.x > k.md) : !(z.x < k.md), v1 = k.strict ? (z.y > k.md) : !(z.y < k.md);  // :301
    float2 r = make_float2(rcp_apx(z.x), rcp_apx(z.y));
    r = ffma2(r, ffma2(z, fmul2(r, k.m1), k.one2), r);  // one Newton step: <= 1 ulp
    const float2 dd = make_float2(v0 ? r.x : 0.0f, v1 ? r.y : 0.0f);
    const float2 w_u = fmul2(wu, make_float2(v0 ? k.wsx : 0.0f, v1 ? k.wsx : 0.0f));
    const float2 w_v = fmul2(wv, make_float2(v0 ? k.wsy : 0.0f, v1 ? k.wsy : 0.0f));
    const float2 X = fmul2(x, dd), Y = fmul2(y, dd), aa = fmul2(h, dd);
    const float2 ru = ffma2(X, k.m1, ffma2(tu, k.ifx2, k.ncx2));  // (:308-309) / fx
    const float2 rv = ffma2(Y, k.m1, ffma2(tv, k.ify2, k.ncy2));
    float2 wJu, wJv;
    if (FULL) {
        const float2 nt2 = c[12];
        const float2 Jzu = fmul2(dd, ffma2(nt2, X, t0));  // dl/dd (:322,363) / fx
        const float2 Jzv = fmul2(dd, ffma2(nt2, Y, t1));
        wJu = fmul2(w_u, Jzu), wJv = fmul2(w_v, Jzv);
        *reinterpret_cast<float2 *>(rawp + 6 * kSubTile) = ffma2(wJu, Jzu, fmul2(wJv, Jzv));  // :325,364
        *reinterpret_cast<float2 *>(rawp + 7 * kSubTile) = ffma2(wJu, ru, fmul2(wJv, rv));    // :326,365
    }
    if (stereo) {  // stereo edges contribute nothing beyond C and w (:329,367); uniform over the warp
        if (FULL) {
#pragma unroll
            for (int i = 0; i < 6; i++) *reinterpret_cast<float2 *>(rawp + i * kSubTile) = make_float2(0.f, 0.f);
        }
        return;
    }
    const float2 P2 = fmul2(X, aa), P3 = fmul2(X, Y), P4 = ffma2(X, X, k.one2), P5 = Y;
    const float2 Q2 = fmul2(Y, aa), Q3 = ffma2(Y, Y, k.one2), Q4 = P3, Q5 = fmul2(X, k.m1);
    if (FULL) {  // u = wu Jzu A + wv Jzv B (E_ij column, :351,385)
        *reinterpret_cast<float2 *>(rawp + 0 * kSubTile) = fmul2(wJu, aa);
        *reinterpret_cast<float2 *>(rawp + 1 * kSubTile) = fmul2(wJv, aa);
        *reinterpret_cast<float2 *>(rawp + 2 * kSubTile) = ffma2(wJu, P2, fmul2(wJv, Q2));
        *reinterpret_cast<float2 *>(rawp + 3 * kSubTile) = ffma2(wJu, P3, fmul2(wJv, Q3));
        *reinterpret_cast<float2 *>(rawp + 4 * kSubTile) = ffma2(wJu, P4, fmul2(wJv, Q4));
        *reinterpret_cast<float2 *>(rawp + 5 * kSubTile) = ffma2(wJu, P5, fmul2(wJv, Q5));
    }
    const float2 wA0 = fmul2(w_u, aa), wA2 = fmul2(w_u, P2), wA3 = fmul2(w_u, P3), wA4 = fmul2(w_u, P4), wA5 = fmul2(w_u, P5);
    const float2 wB1 = fmul2(w_v, aa), wB2 = fmul2(w_v, Q2), wB3 = fmul2(w_v, Q3), wB4 = fmul2(w_v, Q4), wB5 = fmul2(w_v, Q5);
    hh[0] = ffma2(wA0, aa, hh[0]);                       // (0,0)
    hh[1] = ffma2(wB1, aa, hh[1]);                       // (1,1)
    hh[2] = ffma2(wA2, aa, hh[2]);                       // (2,0)
    hh[3] = ffma2(wB2, aa, hh[3]);                       // (2,1)
    hh[4] = ffma2(wA2, P2, ffma2(wB2, Q2, hh[4]));       // (2,2)
    hh[5] = ffma2(wA3, aa, hh[5]);                       // (3,0)
    hh[6] = ffma2(wB3, aa, hh[6]);                       // (3,1)
    hh[7] = ffma2(wA3, P2, ffma2(wB3, Q2, hh[7]));       // (3,2)
    hh[8] = ffma2(wA3, P3, ffma2(wB3, Q3, hh[8]));       // (3,3)
    hh[9] = ffma2(wA4, aa, hh[9]);                       // (4,0)
    hh[10] = ffma2(wB4, aa, hh[10]);                     // (4,1)
    hh[11] = ffma2(wA4, P2, ffma2(wB4, Q2, hh[11]));     // (4,2)
    hh[12] = ffma2(wA4, P3, ffma2(wB4, Q3, hh[12]));     // (4,3)
    hh[13] = ffma2(wA4, P4, ffma2(wB4, Q4, hh[13]));     // (4,4)
    hh[14] = ffma2(wA5, aa, hh[14]);                     // (5,0)
    hh[15] = ffma2(wB5, aa, hh[15]);                     // (5,1)
    hh[16] = ffma2(wA5, P2, ffma2(wB5, Q2, hh[16]));     // (5,2)
    hh[17] = ffma2(wA5, P3, ffma2(wB5, Q3, hh[17]));     // (5,3)
    hh[18] = ffma2(wA5, P4, ffma2(wB5, Q4, hh[18]));     // (5,4)
    hh[19] = ffma2(wA5, P5, ffma2(wB5, Q5, hh[19]));     // (5,5)
    hh[20] = ffma2(wA0, ru, hh[20]);                     // v_j (:347,382)
    hh[21] = ffma2(wB1, rv, hh[21]);
    hh[22] = ffma2(wA2, ru, ffma2(wB2, rv, hh[22]));
    hh[23] = ffma2(wA3, ru, ffma2(wB3, rv, hh[23]));
    hh[24] = ffma2(wA4, ru, ffma2(wB4, rv, hh[24]));
    hh[25] = ffma2(wA5, ru, ffma2(wB5, rv, hh[25]));
    hh[26] = ffma2(fmul2(w_u, ru), ru, ffma2(fmul2(w_v, rv), rv, hh[26]));  // energy
}
// the 27 sums of one (edge, chunk): lanes -> one record, true signs
__device__ __forceinline__ void write_edge_record(float2 (&hh)[kEdgeVals], int lane, float *rec) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < kEdgeVals; i++) acc[i] = hh[i].x + hh[i].y;
#pragma unroll
    for (int i = kEdgeVals; i < 32; i++) acc[i] = 0.0f;
    const float tot = warp_transpose_reduce<32>(acc, lane);
    if (lane < kEdgeVals) rec[lane] = tot * hsign(lane);
}

// ---- variants for the two-kernel pipeline of ba_lin4.cu --------------------------------------------------------------
// common front end of a pixel pair: transform, validity, weights, normalised point, residuals (:284-309)
struct JFront {
    float2 dd, w_u, w_v, X, Y, aa, ru, rv, t0, t1;
};
__device__ __forceinline__ void j_front(const JConst &k, const float2 *__restrict__ c, float2 xn, float2 yn, float2 h, float2 tu, float2 tv,
                                        float2 wu, float2 wv, JFront &f) {
    f.t0 = c[9], f.t1 = c[10];
    const float2 t2 = c[11];
    const float2 x = ffma2(c[0], xn, ffma2(c[1], yn, ffma2(h, f.t0, c[2])));
    const float2 y = ffma2(c[3], xn, ffma2(c[4], yn, ffma2(h, f.t1, c[5])));
    const float2 z = ffma2(c[6], xn, ffma2(c[7], yn, ffma2(h, t2, c[8])));
    const bool v0 = k.strict ? (z.x > k.md) : !(z.x < k.md), v1 = k.strict ? (z.y > k.md) : !(z.y < k.md);  // :301
    float2 r = make_float2(rcp_apx(z.x), rcp_apx(z.y));
    r = ffma2(r, ffma2(z, fmul2(r, k.m1), k.one2), r);  // one Newton step: <= 1 ulp
    f.dd = make_float2(v0 ? r.x : 0.0f, v1 ? r.y : 0.0f);
    f.w_u = fmul2(wu, make_float2(v0 ? k.wsx : 0.0f, v1 ? k.wsx : 0.0f));
    f.w_v = fmul2(wv, make_float2(v0 ? k.wsy : 0.0f, v1 ? k.wsy : 0.0f));
    f.X = fmul2(x, f.dd), f.Y = fmul2(y, f.dd), f.aa = fmul2(h, f.dd);
    f.ru = ffma2(f.X, k.m1, ffma2(tu, k.ifx2, k.ncx2));  // (:308-309) / fx
    f.rv = ffma2(f.Y, k.m1, ffma2(tv, k.ify2, k.ncy2));
}

// the edge's contribution to the disparity block of its source pixels: C += J_z^T w J_z, W += J_z^T w r (:325-326,364-365)
__device__ __forceinline__ void cw_step(const JConst &k, const float2 *__restrict__ c, float2 xn, float2 yn, float2 h, float2 tu, float2 tv,
                                        float2 wu, float2 wv, float2 &C, float2 &W) {
    JFront f;
    j_front(k, c, xn, yn, h, tu, tv, wu, wv, f);
    const float2 nt2 = c[12];
    const float2 Jzu = fmul2(f.dd, ffma2(nt2, f.X, f.t0));  // dl/dd (:322,363) / fx
    const float2 Jzv = fmul2(f.dd, ffma2(nt2, f.Y, f.t1));
    const float2 wJu = fmul2(f.w_u, Jzu), wJv = fmul2(f.w_v, Jzv);
    C = ffma2(ffma2(wJu, Jzu, fmul2(wJv, Jzv)), k.one2, C);
    W = ffma2(ffma2(wJu, f.ru, fmul2(wJv, f.rv)), k.one2, W);
}

// the full step of ONE pixel pair with the pose-disparity vector u_hat (sign-stripped, 6 per pixel) returned in registers; hh
// receives this pair's 27 products (assigned, not accumulated: no zero-initialisation needed).  Not for stereo edges.
__device__ __forceinline__ void j_step_u(const JConst &k, const float2 *__restrict__ c, float2 xn, float2 yn, float2 h, float2 tu,
                                         float2 tv, float2 wu, float2 wv, float2 (&hh)[kEdgeVals], float2 (&u)[6]) {
    JFront f;
    j_front(k, c, xn, yn, h, tu, tv, wu, wv, f);
    const float2 nt2 = c[12];
    const float2 Jzu = fmul2(f.dd, ffma2(nt2, f.X, f.t0));
    const float2 Jzv = fmul2(f.dd, ffma2(nt2, f.Y, f.t1));
    const float2 wJu = fmul2(f.w_u, Jzu), wJv = fmul2(f.w_v, Jzv);
    const float2 X = f.X, Y = f.Y, aa = f.aa, ru = f.ru, rv = f.rv, w_u = f.w_u, w_v = f.w_v;
    const float2 P2 = fmul2(X, aa), P3 = fmul2(X, Y), P4 = ffma2(X, X, k.one2), P5 = Y;
    const float2 Q2 = fmul2(Y, aa), Q3 = ffma2(Y, Y, k.one2), Q4 = P3, Q5 = fmul2(X, k.m1);
    u[0] = fmul2(wJu, aa);  // u = wu Jzu A + wv Jzv B (E_ij column, :351,385)
    u[1] = fmul2(wJv, aa);
    u[2] = ffma2(wJu, P2, fmul2(wJv, Q2));
    u[3] = ffma2(wJu, P3, fmul2(wJv, Q3));
    u[4] = ffma2(wJu, P4, fmul2(wJv, Q4));
    u[5] = ffma2(wJu, P5, fmul2(wJv, Q5));
    const float2 wA0 = fmul2(w_u, aa), wA2 = fmul2(w_u, P2), wA3 = fmul2(w_u, P3), wA4 = fmul2(w_u, P4), wA5 = fmul2(w_u, P5);
    const float2 wB1 = fmul2(w_v, aa), wB2 = fmul2(w_v, Q2), wB3 = fmul2(w_v, Q3), wB4 = fmul2(w_v, Q4), wB5 = fmul2(w_v, Q5);
    hh[0] = fmul2(wA0, aa);
    hh[1] = fmul2(wB1, aa);
    hh[2] = fmul2(wA2, aa);
    hh[3] = fmul2(wB2, aa);
    hh[4] = ffma2(wA2, P2, fmul2(wB2, Q2));
    hh[5] = fmul2(wA3, aa);
    hh[6] = fmul2(wB3, aa);
    hh[7] = ffma2(wA3, P2, fmul2(wB3, Q2));
    hh[8] = ffma2(wA3, P3, fmul2(wB3, Q3));
    hh[9] = fmul2(wA4, aa);
    hh[10] = fmul2(wB4, aa);
    hh[11] = ffma2(wA4, P2, fmul2(wB4, Q2));
    hh[12] = ffma2(wA4, P3, fmul2(wB4, Q3));
    hh[13] = ffma2(wA4, P4, fmul2(wB4, Q4));
    hh[14] = fmul2(wA5, aa);
    hh[15] = fmul2(wB5, aa);
    hh[16] = ffma2(wA5, P2, fmul2(wB5, Q2));
    hh[17] = ffma2(wA5, P3, fmul2(wB5, Q3));
    hh[18] = ffma2(wA5, P4, fmul2(wB5, Q4));
    hh[19] = ffma2(wA5, P5, fmul2(wB5, Q5));
    hh[20] = fmul2(wA0, ru);
    hh[21] = fmul2(wB1, rv);
    hh[22] = ffma2(wA2, ru, fmul2(wB2, rv));
    hh[23] = ffma2(wA3, ru, fmul2(wB3, rv));
    hh[24] = ffma2(wA4, ru, fmul2(wB4, rv));
    hh[25] = ffma2(wA5, ru, fmul2(wB5, rv));
    hh[26] = ffma2(fmul2(w_u, ru), ru, fmul2(fmul2(w_v, rv), rv));
}

// Sum of 27 per-lane values over the warp through a 2 KB shared-memory scratch (two passes of 14 / 13 rows): every lane stores
// its values, then (row, quarter) tasks add 8 lanes each and two shuffles join the quarters.  Lane 4 r' of a pass ends up with
// the total of row r'; written to rec with the true signs.  About half the instructions of the shuffle transpose.
__device__ __forceinline__ void write_edge_record_smem(const float2 (&hh)[kEdgeVals], int lane, float *scratch, float *rec) {
    constexpr int RS = 36;  // row stride in floats: 16-byte aligned rows, conflict-free 128-bit reads
    const int qd = lane & 3, r8 = lane >> 2;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        const int base = 14 * pass, nrow = pass == 0 ? 14 : 13;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 14; i++)
            if (i < nrow) scratch[i * RS + lane] = hh[base + i].x + hh[base + i].y;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int row = r8 + 8 * j;
            float s = 0.0f;
            if (row < nrow) {
                const float4 p0 = *reinterpret_cast<const float4 *>(scratch + row * RS + 8 * qd);
                const float4 p1 = *reinterpret_cast<const float4 *>(scratch + row * RS + 8 * qd + 4);
                s = ((p0.x + p0.y) + (p0.z + p0.w)) + ((p1.x + p1.y) + (p1.z + p1.w));
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (qd == 0 && row < nrow) rec[base + row] = s * hsign(base + row);
        }
    }
}

}  // namespace jmath
}  // namespace vba
