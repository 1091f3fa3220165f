// Packed (2 pixels per thread) version of the per-source-frame linearisation + Schur Gram kernel.
//
// Same outputs as linearize_kernel in ba_kernels.cu (epart / gpart / Q / Qw; reference functions
// projective_transform_kernel, accum_kernel, EEt6x6_kernel, Ev6x1_kernel, csrc/slam_ext/geom_kernels.cu:178-432,
// 863-880,994-1080) but written for Blackwell's packed fp32 pipe: the first profile of the scalar kernel showed it
// issue-bound (67 % issue-active, FMA pipe 36 %, 40 % of all instructions in warp reductions), so here
//   * every thread owns a PIXEL PAIR: float2 loads, FFMA2/FMUL2/FADD2 math, one 27-value warp reduction per
//     64 edge-pixels instead of per 32;
//   * the tile Gram is computed by 8-lane groups (4 block pairs per warp at a time, float4 = 4 pixels per lane
//     per step), so its cross-lane reduction spans 8 lanes instead of 32;
//   * the next edge's targets/weights are prefetched while the current edge is processed.
#include "ba_common.cuh"
#include "ba_launch.h"

namespace vba {

__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 splat(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr int kEc2 = 14;  // float2 entries per edge-constant record: R(9), t(3), flags

// reduce 8-lane groups: V values per lane -> lane (l & 7) holds the group sum of value (l & 7) for V = 8
__device__ __forceinline__ float group8_transpose_reduce(float (&v)[8], int lane) {
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; i++) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// NT = pixel-pair slots per tile (TILE = 2 NT pixels); the CTA has 2 NT threads: the two halves share the tile's
// pixels and split its edges (even / odd), which doubles the warps per SM at the same shared-memory footprint
// (the staging buffer, not registers, bounds occupancy).
// UG: the staging buffer U lives in global memory (a.uglobal) instead of shared memory -- for source frames with more
// outgoing edges than shared memory holds (the reference has no out-degree limit); same code, L2 instead of shared memory.
template <bool UG>
__device__ __forceinline__ float4 ldu4(const float *p) {
    if (UG) return __ldcg(reinterpret_cast<const float4 *>(p));
    return *reinterpret_cast<const float4 *>(p);
}

template <int NT, bool MOTION, bool UG = false>
__global__ void __launch_bounds__(2 * NT, (MOTION && NT == 256) ? 2 : 1) linearize2_kernel(const LinArgs a) {
    constexpr int TILE = NT * 2;
    constexpr int NW = NT / 32;   // warps per half
    constexpr int NTH = 2 * NT;   // threads per CTA
    extern __shared__ __align__(16) float smem[];
    const Tables &tb = a.tb;
    const int tile = blockIdx.x;
    const int k = a.flist ? a.flist[blockIdx.y] : tb.k_lo + blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = tid / NT, slot = tid - half * NT, wh = slot >> 5;
    const int src = tb.kx[k];
    const int s0 = tb.fptr[k];
    const int d = tb.fptr[k + 1] - s0;
    if (MOTION && d == 0) return;
    const int HW = tb.HW;

    float2 *ec2 = reinterpret_cast<float2 *>(smem);           // [d][kEc2] duplicated edge constants
    float *red = smem + 2 * kEc2 * d;                         // [d][NW][27]
    float *const after_red = red + ((d * NW * kEdgeVals + 3) & ~3);
    float *U = UG ? a.uglobal + ((size_t)(s0 - tb.slot_lo) * tb.ntile + (size_t)tile * d) * 6 * TILE : after_red;  // [6d][TILE]
    float *Qs = (UG || MOTION) ? after_red : U + 6 * d * TILE;  // [TILE]
    float *Ws = Qs + TILE;                                    // [TILE]

    for (int m = tid; m < d; m += NTH) {
        const int e = tb.fedge[s0 + m];
        RelPose<float> rp;
        relative_pose<float>(a.poses, src, tb.e_jj[e], rp);
        float2 *c = ec2 + m * kEc2;
#pragma unroll
        for (int q = 0; q < 9; q++) c[q] = splat(rp.R[q]);
#pragma unroll
        for (int q = 0; q < 3; q++) c[9 + q] = splat(rp.t[q]);
        c[12] = make_float2(rp.stereo ? 1.0f : 0.0f, __int_as_float(edge_row(tb, s0 + m, e)));
    }
    const float fx = __ldg(a.intr + 0), fy = __ldg(a.intr + 1), cx = __ldg(a.intr + 2), cy = __ldg(a.intr + 3);
    const float2 fx2 = splat(fx), fy2 = splat(fy);

    const int px0 = tile * TILE + slot * 2;
    const bool inb = px0 < HW;  // HW is even
    float2 h = make_float2(0.0f, 0.0f), xn, yn;
    if (inb) h = __ldg(reinterpret_cast<const float2 *>(a.disps + (size_t)src * HW + px0));
    {
        const int row0 = px0 / tb.wd, col0 = px0 - row0 * tb.wd;
        const int px1 = px0 + 1;
        const int row1 = px1 / tb.wd, col1 = px1 - row1 * tb.wd;
        xn = make_float2(__fdiv_rn((float)col0 - cx, fx), __fdiv_rn((float)col1 - cx, fx));  // geom_kernels.cu:289-290
        yn = make_float2(__fdiv_rn((float)row0 - cy, fy), __fdiv_rn((float)row1 - cy, fy));
    }
    float2 Cacc = make_float2(0.0f, 0.0f), Wacc = make_float2(0.0f, 0.0f);

    // software pipeline: loads of edge m+1 are in flight while edge m is processed
    float2 n_tu = make_float2(0, 0), n_tv = n_tu, n_wu = n_tu, n_wv = n_tu;
    auto issue_loads_e = [&](int e) {
        const size_t base = (size_t)e * 2 * HW + px0;
        n_tu = __ldg(reinterpret_cast<const float2 *>(a.targets + base));
        n_tv = __ldg(reinterpret_cast<const float2 *>(a.targets + base + HW));
        n_wu = __ldg(reinterpret_cast<const float2 *>(a.weights + base));
        n_wv = __ldg(reinterpret_cast<const float2 *>(a.weights + base + HW));
    };
    auto issue_loads = [&](int m) {
        if (inb && m < d) issue_loads_e(__float_as_int(ec2[m * kEc2 + 12].y));
    };
    // the first edge's loads leave before the barrier (edge id straight from the table, not from the staged constants):
    // their trip to HBM overlaps the pose prologue of the d threads above
    if (inb && half < d) issue_loads_e(edge_row(tb, s0 + half, tb.fedge[s0 + half]));
    // sensor disparity and damping of this pixel pair (used after the edge loop by the even-edge half): fetched now
    float2 ds_pre = make_float2(0.0f, 0.0f), et_pre = make_float2(0.0f, 0.0f);
    if (!MOTION && inb && half == 0) {
        ds_pre = __ldg(reinterpret_cast<const float2 *>(a.dsens + (size_t)src * HW + px0));
        et_pre = __ldg(reinterpret_cast<const float2 *>(a.eta + (size_t)k * HW + px0));
    }
    __syncthreads();

    for (int m = half; m < d; m += 2) {
        const float2 *c = ec2 + m * kEc2;
        const bool stereo = c[12].x != 0.0f;
        const float2 tu = n_tu, tv = n_tv, wu = n_wu, wv = n_wv;
        issue_loads(m + 2);

        // X_j = R X_i + h t   (actSE3, :295), X_i = (xn, yn, 1, h)
        const float2 t0 = c[9], t1 = c[10], t2 = c[11];
        const float2 x = ffma2(c[0], xn, ffma2(c[1], yn, ffma2(h, t0, c[2])));
        const float2 y = ffma2(c[3], xn, ffma2(c[4], yn, ffma2(h, t1, c[5])));
        const float2 z = ffma2(c[6], xn, ffma2(c[7], yn, ffma2(h, t2, c[8])));
        const bool v0 = depth_valid(z.x, a.opt) && inb, v1 = depth_valid(z.y, a.opt) && inb;  // :301
        // d = 1/z: MUFU seed + one Newton step (<= 1 ulp), forced to 0 where invalid
        float2 r = make_float2(rcp_approx(z.x), rcp_approx(z.y));
        r = ffma2(r, ffma2(neg2(z), r, splat(1.0f)), r);
        const float2 dd = make_float2(v0 ? r.x : 0.0f, v1 ? r.y : 0.0f);
        const float2 vm = make_float2(v0 ? kWeightScale : 0.0f, v1 ? kWeightScale : 0.0f);  // :304-305
        float2 w_u = fmul2(wu, vm), w_v = fmul2(wv, vm);
        const float2 X = fmul2(x, dd), Y = fmul2(y, dd), aa = fmul2(h, dd);
        // residuals (:308-309)
        const float2 ru = ffma2(neg2(fx2), X, fadd2(tu, splat(-cx)));
        const float2 rv = ffma2(neg2(fy2), Y, fadd2(tv, splat(-cy)));
        // dl/dd (:322,363)
        const float2 Jzu = fmul2(fx2, ffma2(neg2(t2), fmul2(X, dd), fmul2(t0, dd)));
        const float2 Jzv = fmul2(fy2, ffma2(neg2(t2), fmul2(Y, dd), fmul2(t1, dd)));
        float2 wJu = fmul2(w_u, Jzu), wJv = fmul2(w_v, Jzv);
        if (!MOTION) {
            Cacc = ffma2(wJu, Jzu, ffma2(wJv, Jzv, Cacc));  // :325,364
            Wacc = ffma2(wJu, ru, ffma2(wJv, rv, Wacc));    // :326,365
        }
        float acc[32];
#pragma unroll
        for (int q = 0; q < 32; q++) acc[q] = 0.0f;
        if (!stereo) {  // stereo edges contribute nothing beyond C and w (:329,367); uniform over the CTA
            const float2 XY = fmul2(X, Y);
            const float2 X2 = ffma2(X, X, splat(1.0f)), Y2 = ffma2(Y, Y, splat(1.0f));
            if (!MOTION) {
                // u = wu Jzu Jj_u + wv Jzv Jj_v   (E_ij column, :351,385)
                const float2 au = fmul2(wJu, fx2), av = fmul2(wJv, fy2);
                float2 u[6];
                u[0] = fmul2(au, aa);
                u[1] = fmul2(av, aa);
                u[2] = fmul2(neg2(aa), ffma2(au, X, fmul2(av, Y)));
                u[3] = neg2(ffma2(au, XY, fmul2(av, Y2)));
                u[4] = ffma2(au, X2, fmul2(av, XY));
                u[5] = ffma2(av, X, fmul2(neg2(au), Y));
#pragma unroll
                for (int q = 0; q < 6; q++) *reinterpret_cast<float2 *>(U + (size_t)(6 * m + q) * TILE + slot * 2) = u[q];
            }
            // J_j rows (:314-319, :356-361): A = fx*(a,0,-Xa,-XY,1+X^2,-Y), B = fy*(0,a,-Ya,-(1+Y^2),XY,X)
            const float2 A0 = fmul2(fx2, aa), A2 = fmul2(neg2(X), A0), A3 = fmul2(neg2(fx2), XY), A4 = fmul2(fx2, X2), A5 = fmul2(neg2(fx2), Y);
            const float2 B1 = fmul2(fy2, aa), B2 = fmul2(neg2(Y), B1), B3 = fmul2(neg2(fy2), Y2), B4 = fmul2(fy2, XY), B5 = fmul2(fy2, X);
            const float2 wA0 = fmul2(w_u, A0), wA2 = fmul2(w_u, A2), wA3 = fmul2(w_u, A3), wA4 = fmul2(w_u, A4), wA5 = fmul2(w_u, A5);
            const float2 wB1 = fmul2(w_v, B1), wB2 = fmul2(w_v, B2), wB3 = fmul2(w_v, B3), wB4 = fmul2(w_v, B4), wB5 = fmul2(w_v, B5);
            float2 hh[27];
            hh[0] = fmul2(wA0, A0);                        // (0,0)
            hh[1] = fmul2(wB1, B1);                        // (1,1)
            hh[2] = fmul2(wA2, A0);                        // (2,0)
            hh[3] = fmul2(wB2, B1);                        // (2,1)
            hh[4] = ffma2(wA2, A2, fmul2(wB2, B2));        // (2,2)
            hh[5] = fmul2(wA3, A0);                        // (3,0)
            hh[6] = fmul2(wB3, B1);                        // (3,1)
            hh[7] = ffma2(wA3, A2, fmul2(wB3, B2));        // (3,2)
            hh[8] = ffma2(wA3, A3, fmul2(wB3, B3));        // (3,3)
            hh[9] = fmul2(wA4, A0);                        // (4,0)
            hh[10] = fmul2(wB4, B1);                       // (4,1)
            hh[11] = ffma2(wA4, A2, fmul2(wB4, B2));       // (4,2)
            hh[12] = ffma2(wA4, A3, fmul2(wB4, B3));       // (4,3)
            hh[13] = ffma2(wA4, A4, fmul2(wB4, B4));       // (4,4)
            hh[14] = fmul2(wA5, A0);                       // (5,0)
            hh[15] = fmul2(wB5, B1);                       // (5,1)
            hh[16] = ffma2(wA5, A2, fmul2(wB5, B2));       // (5,2)
            hh[17] = ffma2(wA5, A3, fmul2(wB5, B3));       // (5,3)
            hh[18] = ffma2(wA5, A4, fmul2(wB5, B4));       // (5,4)
            hh[19] = ffma2(wA5, A5, fmul2(wB5, B5));       // (5,5)
            hh[20] = fmul2(wA0, ru);                       // v_j (:347,382)
            hh[21] = fmul2(wB1, rv);
            hh[22] = ffma2(wA2, ru, fmul2(wB2, rv));
            hh[23] = ffma2(wA3, ru, fmul2(wB3, rv));
            hh[24] = ffma2(wA4, ru, fmul2(wB4, rv));
            hh[25] = ffma2(wA5, ru, fmul2(wB5, rv));
            hh[26] = ffma2(fmul2(w_u, ru), ru, fmul2(fmul2(w_v, rv), rv));  // energy
#pragma unroll
            for (int q = 0; q < kEdgeVals; q++) acc[q] = hh[q].x + hh[q].y;
            const float tot = warp_transpose_reduce<32>(acc, lane);
            if (lane < kEdgeVals) red[(m * NW + wh) * kEdgeVals + lane] = tot;
        } else {
            if (!MOTION) {
#pragma unroll
                for (int q = 0; q < 6; q++) *reinterpret_cast<float2 *>(U + (size_t)(6 * m + q) * TILE + slot * 2) = make_float2(0.0f, 0.0f);
            }
            if (lane < kEdgeVals) red[(m * NW + wh) * kEdgeVals + lane] = 0.0f;
        }
    }

    if (!MOTION) {
        // the odd-edge half hands its partial C, w to the even-edge half
        if (half == 1) {
            *reinterpret_cast<float2 *>(Qs + slot * 2) = Cacc;
            *reinterpret_cast<float2 *>(Ws + slot * 2) = Wacc;
        }
        __syncthreads();
        if (half == 0) {
            const float2 Co = *reinterpret_cast<const float2 *>(Qs + slot * 2), Wo = *reinterpret_cast<const float2 *>(Ws + slot * 2);
            Cacc = fadd2(Cacc, Co);
            Wacc = fadd2(Wacc, Wo);
            // disparity block: damping / sensor prior (:1359-1370), eliminate: Q = 1/C
            float2 qv = make_float2(0.0f, 0.0f), wz = make_float2(0.0f, 0.0f);
            if (inb) {
                const float2 ds = ds_pre, et = et_pre;
                const int fflags = a.opt.frame_flags ? a.opt.frame_flags[k] : 0;
                disparity_block(Cacc.x, Wacc.x, h.x, ds.x, et.x, fflags, a.opt, qv.x, wz.x);
                disparity_block(Cacc.y, Wacc.y, h.y, ds.y, et.y, fflags, a.opt, qv.y, wz.y);
                *reinterpret_cast<float2 *>(a.qbuf + (size_t)k * HW + px0) = qv;
                *reinterpret_cast<float2 *>(a.qwbuf + (size_t)k * HW + px0) = make_float2(qv.x * wz.x, qv.y * wz.y);
            }
            *reinterpret_cast<float2 *>(Qs + slot * 2) = qv;
            *reinterpret_cast<float2 *>(Ws + slot * 2) = wz;
        }
    }
    __syncthreads();

    // per-(edge, tile) record: fixed-order sum over the CTA's warps
    for (int idx = tid; idx < d * kEdgeVals; idx += NTH) {
        const int m = idx / kEdgeVals, r = idx - m * kEdgeVals;
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < NW; w++) s += red[(m * NW + w) * kEdgeVals + r];
        a.epart[((size_t)(s0 + m) * tb.ntile + tile) * kEdgeStride + r] = s;
    }
    if (MOTION) return;

    // Schur Gram of the tile.  Each 8-lane group of a warp takes one (m, m') block pair; a lane covers 4 consecutive
    // pixels per step (float4 loads: the 8 lanes of a group read 128 contiguous bytes), accumulating two packed
    // pixel pairs into the same float2 accumulators.
    const int npairs = d * (d + 1) / 2;
    const int rec = npairs * 36 + 6 * d;
    float *gp = a.gpart + tb.gbase[k] + (size_t)tile * rec;
    const int grp = lane >> 3, l8 = lane & 7;
    for (int p0 = warp * 4; p0 < npairs; p0 += 2 * NW * 4) {
        const int p = p0 + grp;
        const bool active = p < npairs;
        int m = 0, mp = 0;
        if (active) decode_pair(p, m, mp);
        const float *Um = U + (size_t)6 * m * TILE;
        const float *Up = U + (size_t)6 * mp * TILE;
        const bool diag = (m == mp);
        float2 g2[36], sb2[6];
#pragma unroll
        for (int q = 0; q < 36; q++) g2[q] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int q = 0; q < 6; q++) sb2[q] = make_float2(0.0f, 0.0f);
        if (active) {
#pragma unroll 1
            for (int px = 4 * l8; px < TILE; px += 32) {
                const float4 q4 = *reinterpret_cast<const float4 *>(Qs + px);
                float4 ua4[6];
#pragma unroll
                for (int r = 0; r < 6; r++) ua4[r] = ldu4<UG>(Um + r * TILE + px);
                float2 ualo[6], uahi[6];
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    ualo[r] = fmul2(make_float2(ua4[r].x, ua4[r].y), make_float2(q4.x, q4.y));
                    uahi[r] = fmul2(make_float2(ua4[r].z, ua4[r].w), make_float2(q4.z, q4.w));
                }
                // the next column block is loaded while the current one is consumed (one LDS latency per step instead of six)
                float4 ubn = ldu4<UG>(Up + px);
#pragma unroll
                for (int c2 = 0; c2 < 6; c2++) {
                    const float4 ub4 = ubn;
                    if (c2 < 5) ubn = ldu4<UG>(Up + (c2 + 1) * TILE + px);
                    else if (diag) ubn = *reinterpret_cast<const float4 *>(Ws + px);
                    const float2 blo = make_float2(ub4.x, ub4.y), bhi = make_float2(ub4.z, ub4.w);
#pragma unroll
                    for (int r = 0; r < 6; r++) g2[r * 6 + c2] = ffma2(uahi[r], bhi, ffma2(ualo[r], blo, g2[r * 6 + c2]));
                }
                if (diag) {
                    const float4 w4 = ubn;
                    const float2 wlo = make_float2(w4.x, w4.y), whi = make_float2(w4.z, w4.w);
#pragma unroll
                    for (int r = 0; r < 6; r++) sb2[r] = ffma2(uahi[r], whi, ffma2(ualo[r], wlo, sb2[r]));
                }
            }
        }
        // reduce over the 8 lanes of the group, 8 values at a time; lane l8 ends up with value (round*8 + l8)
        float tot[48];
#pragma unroll
        for (int q = 0; q < 36; q++) tot[q] = g2[q].x + g2[q].y;
#pragma unroll
        for (int q = 0; q < 6; q++) tot[36 + q] = sb2[q].x + sb2[q].y;
#pragma unroll
        for (int q = 42; q < 48; q++) tot[q] = 0.0f;
        float outv[6];
#pragma unroll
        for (int rd = 0; rd < 6; rd++) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = tot[rd * 8 + q];
            outv[rd] = group8_transpose_reduce(v, lane);
        }
        if (active) {
#pragma unroll
            for (int rd = 0; rd < 6; rd++) {
                const int idx = rd * 8 + l8;
                if (idx < 36) gp[(size_t)p * 36 + idx] = outv[rd];
                else if (diag && idx < 42) gp[(size_t)npairs * 36 + 6 * m + (idx - 36)] = outv[rd];
            }
        }
    }
}

// =================================================================================================
static size_t lin2_smem_bytes(int d, int NT, bool motion, bool ug = false) {
    const int NW = NT / 32, TILE = NT * 2;
    size_t fl = (size_t)2 * kEc2 * d + (((size_t)d * NW * kEdgeVals + 3) & ~(size_t)3);
    fl += (motion || ug) ? 2 * TILE : (size_t)6 * d * TILE + 2 * TILE;
    return fl * sizeof(float);
}
size_t lin2_staging_bytes(int d, int NT) { return lin2_smem_bytes(d, NT, false); }
size_t lin2_staging_cap() { return 200 * 1024; }

bool tile_config2(int HW, int dmax, bool motion, int &NT) {
    if (HW % 2 != 0) return false;
    // measured at C3: one 512-thread CTA per SM on a 512-pixel tile (0.27 ms) beats two 256-thread CTAs on 256-pixel
    // tiles (0.35 ms), so the big tile is tried first; smaller ones when the staging buffer would not fit
    const size_t cap2 = 110 * 1024, cap1 = 200 * 1024;
    const int cand[] = {256, 128, 64, 32};
    if (lin2_smem_bytes(dmax, 256, motion) <= cap1) {
        NT = 256;
        return true;
    }
    for (int nt : cand) {
        if (nt == 256 || lin2_smem_bytes(dmax, nt, motion) > cap2) continue;
        NT = nt;
        return true;
    }
    for (int nt : cand) {
        if (lin2_smem_bytes(dmax, nt, motion) > cap1) continue;
        NT = nt;
        return true;
    }
    return false;
}

template <int NT, bool MOTION, bool UG = false>
static cudaError_t launch_lin2_t(const LinArgs &a, int nframes, int dmax, cudaStream_t st) {
    const size_t sm = lin2_smem_bytes(dmax, NT, MOTION, UG);
    if (sm > 227 * 1024) return cudaErrorInvalidValue;
    auto kern = linearize2_kernel<NT, MOTION, UG>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (err != cudaSuccess) return err;
    dim3 grid(a.tb.ntile, nframes);
    kern<<<grid, 2 * NT, sm, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_linearize2(const LinArgs &a, int nframes, int dmax, bool motion, int NT, cudaStream_t st, bool ug) {
    if (nframes <= 0) return cudaSuccess;
    if (ug) {  // hub frames: u staged in global memory
        if (motion || !a.uglobal) return cudaErrorInvalidValue;
        if (NT == 256) return launch_lin2_t<256, false, true>(a, nframes, dmax, st);
        if (NT == 128) return launch_lin2_t<128, false, true>(a, nframes, dmax, st);
        if (NT == 64) return launch_lin2_t<64, false, true>(a, nframes, dmax, st);
        if (NT == 32) return launch_lin2_t<32, false, true>(a, nframes, dmax, st);
        return cudaErrorInvalidValue;
    }
#define VBA_LIN2(NT_)                                                                                    \
    if (NT == NT_)                                                                                       \
        return motion ? launch_lin2_t<NT_, true>(a, nframes, dmax, st) : launch_lin2_t<NT_, false>(a, nframes, dmax, st);
    VBA_LIN2(256)
    VBA_LIN2(128)
    VBA_LIN2(64)
    VBA_LIN2(32)
#undef VBA_LIN2
    return cudaErrorInvalidValue;
}

}  // namespace vba
