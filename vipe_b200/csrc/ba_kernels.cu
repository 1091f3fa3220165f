// Dense bundle adjustment: per-source-frame linearisation + Schur Gram, reduced-system assembly,
// back-substitution and retraction kernels (sm_100a).
//
// What the reference does with projective_transform_kernel / accum_kernel / EEt6x6 / Ev6x1 / EvT6x1 /
// pose_retr / disp_retr plus host-side Eigen assembly (csrc/slam_ext/geom_kernels.cu:178-432,863-1098,
// 1100-1281) is re-organised here around the SOURCE FRAME: every quantity that couples through a
// disparity (C, w, E) belongs to the frame the edge leaves from, so one CTA owns a (frame, pixel tile)
// pair, keeps the per-pixel disparity block in registers, stages the per-edge pose-disparity vectors in
// shared memory and eliminates the disparities right there.  Nothing pixel-sized is ever reduced with
// global atomics and the E arrays (2 x 24 B per edge-pixel in the reference) never reach HBM.
#include "ba_common.cuh"
#include "ba_launch.h"

namespace vba {

// ------------------------------------------------------------------------------------------------
// vector loads of PPT consecutive floats (PPT = 1, 2, 4); the host guarantees alignment (HW % PPT == 0).
template <int PPT>
__device__ __forceinline__ void load_px(const float *__restrict__ p, float (&v)[PPT]) {
    if constexpr (PPT == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    } else if constexpr (PPT == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
        v[0] = t.x, v[1] = t.y;
    } else {
        v[0] = __ldg(p);
    }
}
template <int PPT>
__device__ __forceinline__ void store_px(float *p, const float (&v)[PPT]) {
    if constexpr (PPT == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (PPT == 2) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    } else {
        p[0] = v[0];
    }
}

// Per edge-pixel geometry shared by the linearisation and the back-substitution.
// Follows geom_kernels.cu:289-322,356-363 with X = x*d, Y = y*d, a = h*d.
struct PixelGeom {
    float X, Y, a, dinv;  // dinv = d (inverse depth in frame j), 0 when invalid
    float Jzu, Jzv;
    bool valid;
};

__device__ __forceinline__ PixelGeom pixel_geom(const float *__restrict__ ec, float xn, float yn, float h, float fx,
                                                float fy, const Options &opt) {
    PixelGeom g;
    const float x = fmaf(ec[0], xn, fmaf(ec[1], yn, fmaf(h, ec[9], ec[2])));
    const float y = fmaf(ec[3], xn, fmaf(ec[4], yn, fmaf(h, ec[10], ec[5])));
    const float z = fmaf(ec[6], xn, fmaf(ec[7], yn, fmaf(h, ec[11], ec[8])));
    g.valid = depth_valid(z, opt);
    const float d = g.valid ? __frcp_rn(z) : 0.0f;
    g.dinv = d;
    g.X = x * d;
    g.Y = y * d;
    g.a = h * d;
    g.Jzu = fx * (ec[9] * d - ec[11] * (g.X * d));
    g.Jzv = fy * (ec[10] * d - ec[11] * (g.Y * d));
    return g;
}

// u = wu*Jzu*Jj_u + wv*Jzv*Jj_v  (the per-pixel E_ij column, geom_kernels.cu:351,385)
__device__ __forceinline__ void edge_u(const PixelGeom &g, float au /*wu*Jzu*fx*/, float av /*wv*Jzv*fy*/, float (&u)[6]) {
    const float XY = g.X * g.Y;
    u[0] = au * g.a;
    u[1] = av * g.a;
    u[2] = -g.a * fmaf(au, g.X, av * g.Y);
    u[3] = -fmaf(au, XY, av * fmaf(g.Y, g.Y, 1.0f));
    u[4] = fmaf(au, fmaf(g.X, g.X, 1.0f), av * XY);
    u[5] = fmaf(av, g.X, -au * g.Y);
}

__device__ __forceinline__ void write_edge_consts(float *ec, const RelPose<float> &rp, int e) {
#pragma unroll
    for (int q = 0; q < 9; q++) ec[q] = rp.R[q];
    ec[9] = rp.t[0];
    ec[10] = rp.t[1];
    ec[11] = rp.t[2];
    ec[12] = rp.stereo ? 1.0f : 0.0f;
    ec[13] = __int_as_float(e);
}

// =================================================================================================
// Stage 1+2+3a: linearise every edge leaving one source frame over one pixel tile, then eliminate the
// tile's disparities:   per (edge, tile): H_jj (20 unique), v_j (6), energy        -> epart
//                       per (frame, tile): Y_mm' = sum_px Q u_m u_m'^T, sum_px Q w u_m -> gpart
//                       per frame pixel:  Q = 1/C, Q*w                                -> qbuf, qwbuf
template <int NT, int PPT, bool MOTION>
__global__ void __launch_bounds__(NT) linearize_kernel(const LinArgs a) {
    constexpr int TILE = NT * PPT;
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) float smem[];
    const Tables &tb = a.tb;
    const int tile = blockIdx.x;
    const int k = tb.k_lo + blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int src = tb.kx[k];
    const int s0 = tb.fptr[k];
    const int d = tb.fptr[k + 1] - s0;
    if (MOTION && d == 0) return;
    const int HW = tb.HW;

    float *ec = smem;                                          // [d][16]
    float *red = ec + d * kEcStride;                           // [d][NW][27]
    float *U = red + ((d * NW * kEdgeVals + 3) & ~3);          // [6d][TILE]
    float *Qs = U + (MOTION ? 0 : 6 * d * TILE);               // [TILE]
    float *Ws = Qs + TILE;                                     // [TILE]

    for (int m = tid; m < d; m += NT) {
        const int e = tb.fedge[s0 + m];
        RelPose<float> rp;
        relative_pose<float>(a.poses, src, tb.e_jj[e], rp);
        write_edge_consts(ec + m * kEcStride, rp, edge_row(tb, s0 + m, e));
    }
    const float fx = __ldg(a.intr + 0), fy = __ldg(a.intr + 1), cx = __ldg(a.intr + 2), cy = __ldg(a.intr + 3);

    const int px0 = tile * TILE + tid * PPT;
    const bool inb = px0 < HW;  // HW % PPT == 0 => all PPT pixels share it
    float xn[PPT], yn[PPT], h[PPT], Cacc[PPT], Wacc[PPT];
    if (inb) {
        load_px<PPT>(a.disps + (size_t)src * HW + px0, h);
    }
#pragma unroll
    for (int s = 0; s < PPT; s++) {
        const int px = px0 + s;
        const int row = px / tb.wd;
        const int col = px - row * tb.wd;
        xn[s] = __fdiv_rn((float)col - cx, fx);  // geom_kernels.cu:289-290
        yn[s] = __fdiv_rn((float)row - cy, fy);
        if (!inb) h[s] = 0.0f;
        Cacc[s] = 0.0f;
        Wacc[s] = 0.0f;
    }
    __syncthreads();

    for (int m = 0; m < d; m++) {
        const float *c = ec + m * kEcStride;
        const int e = __float_as_int(c[13]);
        const bool stereo = c[12] != 0.0f;
        float tu[PPT], tv[PPT], wu[PPT], wv[PPT];
        if (inb) {
            const size_t base = (size_t)e * 2 * HW + px0;
            load_px<PPT>(a.targets + base, tu);
            load_px<PPT>(a.targets + base + HW, tv);
            load_px<PPT>(a.weights + base, wu);
            load_px<PPT>(a.weights + base + HW, wv);
        } else {
#pragma unroll
            for (int s = 0; s < PPT; s++) tu[s] = tv[s] = wu[s] = wv[s] = 0.0f;
        }
        float acc[32];
#pragma unroll
        for (int q = 0; q < 32; q++) acc[q] = 0.0f;
        float us[6][PPT];
#pragma unroll
        for (int s = 0; s < PPT; s++) {
            const PixelGeom g = pixel_geom(c, xn[s], yn[s], h[s], fx, fy, a.opt);
            float w_u = g.valid ? kWeightScale * wu[s] : 0.0f;  // :304-305
            float w_v = g.valid ? kWeightScale * wv[s] : 0.0f;
            const float ru = tu[s] - fmaf(fx, g.X, cx);  // :308-309
            const float rv = tv[s] - fmaf(fy, g.Y, cy);
            if (!MOTION) {
                Cacc[s] = fmaf(w_u * g.Jzu, g.Jzu, fmaf(w_v * g.Jzv, g.Jzv, Cacc[s]));          // :325,364
                Wacc[s] = fmaf(w_u * ru, g.Jzu, fmaf(w_v * rv, g.Jzv, Wacc[s]));                // :326,365
            }
            if (stereo) w_u = w_v = 0.0f;  // :329,367
            if (!MOTION) {
                float u6[6];
                edge_u(g, w_u * g.Jzu * fx, w_v * g.Jzv * fy, u6);
#pragma unroll
                for (int r = 0; r < 6; r++) us[r][s] = u6[r];
            }
            // J_j rows (:314-319, :356-361): A = fx*(a,0,-Xa,-XY,1+X^2,-Y), B = fy*(0,a,-Ya,-(1+Y^2),XY,X)
            const float XY = g.X * g.Y;
            const float A0 = fx * g.a, A2 = -fx * g.X * g.a, A3 = -fx * XY, A4 = fx * fmaf(g.X, g.X, 1.0f), A5 = -fx * g.Y;
            const float B1 = fy * g.a, B2 = -fy * g.Y * g.a, B3 = -fy * fmaf(g.Y, g.Y, 1.0f), B4 = fy * XY, B5 = fy * g.X;
            const float wA0 = w_u * A0, wA2 = w_u * A2, wA3 = w_u * A3, wA4 = w_u * A4, wA5 = w_u * A5;
            const float wB1 = w_v * B1, wB2 = w_v * B2, wB3 = w_v * B3, wB4 = w_v * B4, wB5 = w_v * B5;
            // H_jj lower triangle in hslot() order
            acc[0] = fmaf(wA0, A0, acc[0]);                       // (0,0)
            acc[1] = fmaf(wB1, B1, acc[1]);                       // (1,1)
            acc[2] = fmaf(wA2, A0, acc[2]);                       // (2,0)
            acc[3] = fmaf(wB2, B1, acc[3]);                       // (2,1)
            acc[4] = fmaf(wA2, A2, fmaf(wB2, B2, acc[4]));        // (2,2)
            acc[5] = fmaf(wA3, A0, acc[5]);                       // (3,0)
            acc[6] = fmaf(wB3, B1, acc[6]);                       // (3,1)
            acc[7] = fmaf(wA3, A2, fmaf(wB3, B2, acc[7]));        // (3,2)
            acc[8] = fmaf(wA3, A3, fmaf(wB3, B3, acc[8]));        // (3,3)
            acc[9] = fmaf(wA4, A0, acc[9]);                       // (4,0)
            acc[10] = fmaf(wB4, B1, acc[10]);                     // (4,1)
            acc[11] = fmaf(wA4, A2, fmaf(wB4, B2, acc[11]));      // (4,2)
            acc[12] = fmaf(wA4, A3, fmaf(wB4, B3, acc[12]));      // (4,3)
            acc[13] = fmaf(wA4, A4, fmaf(wB4, B4, acc[13]));      // (4,4)
            acc[14] = fmaf(wA5, A0, acc[14]);                     // (5,0)
            acc[15] = fmaf(wB5, B1, acc[15]);                     // (5,1)
            acc[16] = fmaf(wA5, A2, fmaf(wB5, B2, acc[16]));      // (5,2)
            acc[17] = fmaf(wA5, A3, fmaf(wB5, B3, acc[17]));      // (5,3)
            acc[18] = fmaf(wA5, A4, fmaf(wB5, B4, acc[18]));      // (5,4)
            acc[19] = fmaf(wA5, A5, fmaf(wB5, B5, acc[19]));      // (5,5)
            // v_j (:347,382) and energy
            acc[20] = fmaf(wA0, ru, acc[20]);
            acc[21] = fmaf(wB1, rv, acc[21]);
            acc[22] = fmaf(wA2, ru, fmaf(wB2, rv, acc[22]));
            acc[23] = fmaf(wA3, ru, fmaf(wB3, rv, acc[23]));
            acc[24] = fmaf(wA4, ru, fmaf(wB4, rv, acc[24]));
            acc[25] = fmaf(wA5, ru, fmaf(wB5, rv, acc[25]));
            acc[26] = fmaf(w_u * ru, ru, fmaf(w_v * rv, rv, acc[26]));
        }
        if (!MOTION) {
#pragma unroll
            for (int r = 0; r < 6; r++) store_px<PPT>(U + (size_t)(6 * m + r) * TILE + tid * PPT, us[r]);
        }
        const float tot = warp_transpose_reduce<32>(acc, lane);
        if (lane < kEdgeVals) red[(m * NW + warp) * kEdgeVals + lane] = tot;
    }

    if (!MOTION) {
        // disparity block: damping / sensor prior (:1359-1370), eliminate: Q = 1/C
        const int fflags = a.opt.frame_flags ? a.opt.frame_flags[k] : 0;
        float qv[PPT], wv2[PPT], qw[PPT];
        if (inb) {
            float ds[PPT], et[PPT];
            load_px<PPT>(a.dsens + (size_t)src * HW + px0, ds);
            load_px<PPT>(a.eta + (size_t)k * HW + px0, et);
#pragma unroll
            for (int s = 0; s < PPT; s++) {
                float W;
                disparity_block(Cacc[s], Wacc[s], h[s], ds[s], et[s], fflags, a.opt, qv[s], W);
                wv2[s] = W;
                qw[s] = qv[s] * W;
            }
            store_px<PPT>(a.qbuf + (size_t)k * HW + px0, qv);
            store_px<PPT>(a.qwbuf + (size_t)k * HW + px0, qw);
        } else {
#pragma unroll
            for (int s = 0; s < PPT; s++) qv[s] = wv2[s] = 0.0f;
        }
        store_px<PPT>(Qs + tid * PPT, qv);
        store_px<PPT>(Ws + tid * PPT, wv2);
    }
    __syncthreads();

    // per-(edge, tile) record: fixed-order sum over the CTA's warps
    for (int idx = tid; idx < d * kEdgeVals; idx += NT) {
        const int m = idx / kEdgeVals, r = idx - m * kEdgeVals;
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < NW; w++) s += red[(m * NW + w) * kEdgeVals + r];
        a.epart[((size_t)(s0 + m) * tb.ntile + tile) * kEdgeStride + r] = s;
    }

    if (MOTION) return;

    // Schur Gram of the tile: one warp per (m, m') block pair, lanes stride over the tile's pixels.
    const int npairs = d * (d + 1) / 2;
    const int rec = npairs * 36 + 6 * d;
    float *gp = a.gpart + tb.gbase[k] + (size_t)tile * rec;
    for (int p = warp; p < npairs; p += NW) {
        int m, mp;
        decode_pair(p, m, mp);
        const float *Um = U + (size_t)6 * m * TILE;
        const float *Up = U + (size_t)6 * mp * TILE;
        // each lane takes pixel PAIRS (float2 loads, packed FFMA2): lanes stride over the tile 64 pixels at a time
        float2 g2[36];
#pragma unroll
        for (int q = 0; q < 36; q++) g2[q] = make_float2(0.0f, 0.0f);
        float2 sb2[6];
#pragma unroll
        for (int q = 0; q < 6; q++) sb2[q] = make_float2(0.0f, 0.0f);
        const bool diag = (m == mp);
#pragma unroll 2
        for (int px = 2 * lane; px < TILE; px += 64) {
            const float2 q = *reinterpret_cast<const float2 *>(Qs + px);
            float2 ua[6], ub[6];
#pragma unroll
            for (int r = 0; r < 6; r++) {
                ua[r] = fmul2(*reinterpret_cast<const float2 *>(Um + r * TILE + px), q);
                ub[r] = *reinterpret_cast<const float2 *>(Up + r * TILE + px);
            }
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c2 = 0; c2 < 6; c2++) g2[r * 6 + c2] = ffma2(ua[r], ub[c2], g2[r * 6 + c2]);
            if (diag) {
                const float2 w = *reinterpret_cast<const float2 *>(Ws + px);
#pragma unroll
                for (int r = 0; r < 6; r++) sb2[r] = ffma2(ua[r], w, sb2[r]);
            }
        }
        float g[36], sb[6];
#pragma unroll
        for (int q = 0; q < 36; q++) g[q] = g2[q].x + g2[q].y;
#pragma unroll
        for (int q = 0; q < 6; q++) sb[q] = sb2[q].x + sb2[q].y;
        float v32[32];
#pragma unroll
        for (int q = 0; q < 32; q++) v32[q] = g[q];
        const float t32 = warp_transpose_reduce<32>(v32, lane);
        float v16[16];
#pragma unroll
        for (int q = 0; q < 4; q++) v16[q] = g[32 + q];
#pragma unroll
        for (int q = 0; q < 6; q++) v16[4 + q] = sb[q];
#pragma unroll
        for (int q = 10; q < 16; q++) v16[q] = 0.0f;
        const float t16 = warp_transpose_reduce<16>(v16, lane);
        gp[(size_t)p * 36 + lane] = t32;
        if (lane < 4) gp[(size_t)p * 36 + 32 + lane] = t16;
        if (diag && lane >= 4 && lane < 10) gp[(size_t)npairs * 36 + 6 * m + (lane - 4)] = t16;
    }
}

// =================================================================================================
// Stage 3b: per source frame, sum the tile partials in fp64, apply the per-edge adjoint G (J_i = G J_j)
// to obtain the source-pose rows, and add everything into the dense reduced camera system.
// With M_mm' = [m==m'] H_jj,m - Y_mm' and g_m = v_j,m - sum_px Q w u_m:
//     block(j_m, j_m') += M_mm'            block(i, j_m') += sum_m G_m M_mm'
//     block(i, i)      += sum_mm' G_m M_mm' G_m'^T
//     rhs(j_m) += g_m                      rhs(i) += sum_m G_m g_m
// which is (A - S) and (b - v) of geom_kernels.cu:1343-1347,1377-1378 assembled in fp64 (Q7).
__device__ __forceinline__ void add_block_entry(double *H, int n, int pa, int r, int pb, int c, double v) {
    // entry (6pa+r, 6pb+c) of the symmetric matrix; only the lower triangle (and full diagonal blocks) is stored
    if (pa > pb) {
        atomicAdd(H + (size_t)(6 * pa + r) * n + 6 * pb + c, v);
    } else if (pa < pb) {
        atomicAdd(H + (size_t)(6 * pb + c) * n + 6 * pa + r, v);
    } else {
        atomicAdd(H + (size_t)(6 * pa + r) * n + 6 * pb + c, v);
    }
}

// sum of `n` fp32 partial records spaced `stride` apart, in tile order, in fp64.  The loads of a chunk of 8 are issued
// before the first add: a rolled `s += p[t * stride]` loop serialises them (the add waits for its load and, issue being
// in order, holds back the next load), which made this latency-bound kernel pay one trip to L2 per tile.
__device__ __forceinline__ double tile_sum(const float *p, size_t stride, int n) {
    double s = 0.0;
    for (int t0 = 0; t0 < n; t0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = (t0 + u < n) ? __ldg(p + (size_t)(t0 + u) * stride) : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; u++) s += (double)v[u];
    }
    return s;
}

__global__ void __launch_bounds__(256) frame_reduce_kernel(const ReduceArgs a) {
    extern __shared__ __align__(16) double dsm[];
    const Tables &tb = a.tb;
    const int k = tb.k_lo + blockIdx.x;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int src = tb.kx[k];
    const int s0 = tb.fptr[k];
    const int d = tb.fptr[k + 1] - s0;
    if (d == 0) return;
    const int ntile = tb.ntile;
    const int prob = tb.frame_prob[k];
    const int n = tb.prob_npad[prob];
    double *const hsys = a.hsys + tb.prob_hoff[prob];
    double *const bsys = hsys + (size_t)n * n;
    double *const adiag = bsys + n;
    const int ai = tb.pose_sys[src];
    const size_t cb = (size_t)a.cbase[k], vb = (size_t)a.vbase[k];

    // hub frames: the per-edge working set lives in global memory (same code, the CTA's barriers order its accesses)
    double *const wbase = a.fs_global ? a.fscratch + (size_t)(s0 - tb.slot_lo) * kReduceDoubles : dsm;
    double *G = wbase;          // [d][36]
    double *T = G + d * 36;     // [d][36]
    double *hs = T + d * 36;    // [d][27]   summed edge records
    double *gv = hs + d * 27;   // [d][6]    g_m
    double *fs = gv + d * 6;    // [d][14]   summed focal records   (focal length free only)
    double *cf = fs + (a.fpart ? d * 14 : 0);  // [d][6]  border column contribution of edge m's target pose
    int *aj = reinterpret_cast<int *>(cf + (a.fpart ? d * 6 : 0));  // [d]

    for (int m = tid; m < d; m += NT) {
        const int e = tb.fedge[s0 + m];
        const int j = tb.e_jj[e];
        RelPose<double> rp;
        relative_pose<double>(a.poses, src, j, rp);
        adjoint_G<double>(rp, G + m * 36);
        aj[m] = rp.stereo ? -1 : tb.pose_sys[j];
    }
    for (int idx = tid; idx < d * kEdgeVals; idx += NT) {
        const int m = idx / kEdgeVals, r = idx - m * kEdgeVals;
        const float *ep = a.epart + ((size_t)(s0 + m) * tb.ntile_e) * kEdgeStride + r;
        hs[idx] = tile_sum(ep, kEdgeStride, tb.ntile_e);
    }
    __syncthreads();

    const int npairs = d * (d + 1) / 2;
    const int rec = npairs * 36 + 6 * d;
    // M(m, m') blocks: written below, read back by every thread for T = sum_m G_m M; shared memory when they fit
    double *msc = a.msc_smem ? reinterpret_cast<double *>(aj + ((d + 1) & ~1)) : a.msc + tb.mbase[k];
    const float *gp = a.gpart + tb.gbase[k];

    // g_m = v_j,m - sum_tiles sb_m
    for (int idx = tid; idx < d * 6; idx += NT) {
        const int m = idx / 6, r = idx - m * 6;
        const double s = a.motion_only ? 0.0 : tile_sum(gp + npairs * 36 + idx, rec, ntile);
        const double g = hs[m * kEdgeVals + 20 + r] - s;
        gv[idx] = g;
        a.cvec[(vb + m) * 6 + r] = g;  // -> rhs(j_m)
        const int sl = hslot(r, r);
        a.cvec2[(vb + m) * 6 + r] = sl >= 0 ? hs[m * kEdgeVals + sl] : 0.0;  // -> diag(A)(j_m): H_jj,m alone
    }
    // M blocks
    for (int idx = tid; idx < npairs * 36; idx += NT) {
        const int p = idx / 36, rc = idx - p * 36;
        const int r = rc / 6, c = rc - r * 6;
        int m, mp;
        decode_pair(p, m, mp);
        const double y = a.motion_only ? 0.0 : tile_sum(gp + idx, rec, ntile);
        double v = -y;
        if (m == mp) {
            const int hi = r >= c ? r : c, lo = r >= c ? c : r;
            const int sl = hslot(hi, lo);
            if (sl >= 0) v += hs[m * kEdgeVals + sl];
        }
        msc[idx] = v;
        a.cblk[(cb + p) * 36 + rc] = v;  // -> block (j_m, j_m') (and its transpose when both edges reach the same pose)
    }
    // Focal length (one more variable, row/column `focal_row` of the reduced system):
    //   H(f, j_m) += J_j^T w J_f - sum_px Q u_f u_m,  H(f,f) += (1 + lm_f) J_f^T w J_f - sum_px Q u_f^2,  b(f) += J_f^T w r - sum_px Q w u_f
    if (a.fpart) {
        const int nf = a.focal_row, ntf = a.ntile_f;
        for (int idx = tid; idx < d * 14; idx += NT) {
            const int m = idx / 14, r = idx - m * 14;
            double s = 0.0;
            if (r < 8 || !a.motion_only) {
                const float *fp = a.fpart + ((size_t)(s0 + m) * ntf) * kFocalStride + r;
                for (int t = 0; t < ntf; t++) s += (double)fp[(size_t)t * kFocalStride];
            }
            fs[idx] = s;
        }
        __syncthreads();
        for (int idx = tid; idx < d * 6; idx += NT) {
            const int m = idx / 6, r = idx - m * 6;
            const double c = fs[m * 14 + r] - fs[m * 14 + 8 + r];
            cf[idx] = c;
            if (aj[m] >= 0) atomicAdd(hsys + (size_t)nf * n + 6 * aj[m] + r, c);
        }
        if (tid == 0) {
            double hff = 0.0, vf = 0.0, yff = 0.0, sf = 0.0;
            for (int m = 0; m < d; m++) {  // every edge counts, whether or not its target pose is free
                hff += fs[m * 14 + 6];
                vf += fs[m * 14 + 7];
            }
            if (!a.motion_only)
                for (int t = 0; t < ntf; t++) {
                    yff += (double)a.ffpart[((size_t)k * ntf + t) * 2];
                    sf += (double)a.ffpart[((size_t)k * ntf + t) * 2 + 1];
                }
            atomicAdd(hsys + (size_t)nf * n + nf, hff * (1.0 + (double)a.focal_lm) - yff);
            atomicAdd(bsys + nf, vf - sf);
        }
    }
    if (ai < 0) return;  // source pose fixed: no source rows (uniform over the CTA)
    __syncthreads();

    // T_m' = sum_m G_m M(m, m')
    for (int idx = tid; idx < d * 36; idx += NT) {
        const int mp = idx / 36, rc = idx - mp * 36;
        const int r = rc / 6, c = rc - r * 6;
        double s = 0.0;
        for (int m = 0; m < d; m++) {
            const double *Gm = G + m * 36 + r * 6;
            if (m <= mp) {
                const double *Mb = msc + (size_t)(mp * (mp + 1) / 2 + m) * 36;
#pragma unroll
                for (int q = 0; q < 6; q++) s += Gm[q] * Mb[q * 6 + c];
            } else {  // stored block is (mp, m): use its transpose
                const double *Mb = msc + (size_t)(m * (m + 1) / 2 + mp) * 36;
#pragma unroll
                for (int q = 0; q < 6; q++) s += Gm[q] * Mb[c * 6 + q];
            }
        }
        T[idx] = s;
        a.cblk[(cb + npairs + mp) * 36 + rc] = s;  // -> block (i, j_m')
    }
    __syncthreads();
    // Z = sum_m' T_m' G_m'^T  -> block (i,i);   rhs(i) += sum_m G_m g_m
    if (tid < 36) {
        const int r = tid / 6, c = tid - r * 6;
        double s = 0.0;
        for (int mp = 0; mp < d; mp++) {
#pragma unroll
            for (int q = 0; q < 6; q++) s += T[mp * 36 + r * 6 + q] * G[mp * 36 + c * 6 + q];
        }
        a.cblk[(cb + npairs + d) * 36 + tid] = s;  // -> block (i, i)
    } else if (tid >= 64 && tid < 70) {
        const int r = tid - 64;
        double s = 0.0;
        for (int m = 0; m < d; m++) {
#pragma unroll
            for (int q = 0; q < 6; q++) s += G[m * 36 + r * 6 + q] * gv[m * 6 + q];
        }
        a.cvec[(vb + d) * 6 + r] = s;  // -> rhs(i)
    } else if (tid >= 128 && tid < 134) {
        if (a.fpart) {  // H(f, i) += sum_m G_m c_m
            const int r = tid - 128;
            double s = 0.0;
            for (int m = 0; m < d; m++) {  // stereo edges have c_m = 0
#pragma unroll
                for (int q = 0; q < 6; q++) s += G[m * 36 + r * 6 + q] * cf[m * 6 + q];
            }
            atomicAdd(hsys + (size_t)a.focal_row * n + 6 * ai + r, s);
        }
    } else if (tid >= 96 && tid < 102) {
        // diagonal of sum_m G_m H_jj,m G_m^T alone (the source pose's Hessian before the Schur complement)
        const int r = tid - 96;
        double s = 0.0;
        for (int m = 0; m < d; m++) {
            const double *Gm = G + m * 36 + r * 6;
            for (int p = 0; p < 6; p++)
                for (int q = 0; q < 6; q++) {
                    const int sl = hslot(p >= q ? p : q, p >= q ? q : p);
                    if (sl >= 0) s += Gm[p] * hs[m * kEdgeVals + sl] * Gm[q];
                }
        }
        a.cvec2[(vb + d) * 6 + r] = s;  // -> diag(A)(i)
    }
}

// Deterministic assembly of the reduced camera system: one 36-thread group per destination block (6 threads per destination
// vector) adds up its contribution slots in the order the plan lists them -- sorted by source frame, then by slot -- so the
// result does not depend on scheduling (the reference's Eigen assembly is deterministic as well).
__global__ void __launch_bounds__(252) assemble_kernel(const AssembleArgs a) {
    const int grp = threadIdx.x / 36, rc = threadIdx.x - grp * 36;
    const int nbg = (a.nb + 6) / 7;
    if ((int)blockIdx.x < nbg) {
        const int b = blockIdx.x * 7 + grp;
        if (b >= a.nb) return;
        const int r = rc / 6, c = rc - r * 6, rct = c * 6 + r;
        double s = 0.0;
        for (int q = a.bsrc_ptr[b]; q < a.bsrc_ptr[b + 1]; q++) {
            const int src = a.bsrc[q];
            s += a.cblk[(size_t)(src >> 1) * 36 + ((src & 1) ? rct : rc)];
        }
        a.sys[a.bdst_off[b] + (long long)r * a.bdst_ld[b] + c] += s;
    } else {
        const int v = ((int)blockIdx.x - nbg) * 42 + threadIdx.x / 6, r = threadIdx.x % 6;
        if (v >= a.nv) return;
        double s = 0.0, s2 = 0.0;
        for (int q = a.vsrc_ptr[v]; q < a.vsrc_ptr[v + 1]; q++) {
            const int src = a.vsrc[q];
            s += a.cvec[(size_t)src * 6 + r];
            s2 += a.cvec2[(size_t)src * 6 + r];
        }
        a.sys[a.vdst_off[v] + r] += s;
        a.sys[a.vdst_off[v] + a.vdst_adiag[v] + r] += s2;
    }
}

cudaError_t launch_assemble(const AssembleArgs &a, cudaStream_t st) {
    const int grid = (a.nb + 6) / 7 + (a.nv + 41) / 42;
    if (grid <= 0) return cudaSuccess;
    assemble_kernel<<<grid, 252, 0, st>>>(a);
    return cudaGetLastError();
}

// =================================================================================================
// Stage 4a: back-substitution + disparity retraction for one (frame, tile):
//   dz = Q (w - sum_m u_m . y_m),  y_m = G_m^T dx_i [0 < i-t0 < P] + dx_j [0 < j-t0 < P]
// (EvT6x1_kernel with its `idx <= 0` skip, geom_kernels.cu:1082-1098,1380-1390; disp_retr_kernel :933-944).
template <int NT, int PPT>
__global__ void __launch_bounds__(NT) backsub_kernel(const BackArgs a) {
    constexpr int TILE = NT * PPT;
    extern __shared__ __align__(16) float smem[];
    const Tables &tb = a.tb;
    const int tile = blockIdx.x;
    const int k = tb.k_lo + blockIdx.y;
    const int tid = threadIdx.x;
    const int src = tb.kx[k];
    const int s0 = tb.fptr[k];
    const int d = tb.fptr[k + 1] - s0;
    const int HW = tb.HW;
    float *ec = smem;                   // [d][16]
    float *ys = ec + d * kEcStride;     // [d][8]

    for (int m = tid; m < d; m += NT) {
        const int e = tb.fedge[s0 + m];
        const int j = tb.e_jj[e];
        RelPose<float> rp;
        relative_pose<float>(a.poses, src, j, rp);
        write_edge_consts(ec + m * kEcStride, rp, edge_row(tb, s0 + m, e));
        float G[36];
        adjoint_G<float>(rp, G);
        const int lo = a.opt.backsub_all_poses ? 0 : 1;  // Q4: the reference skips pose index 0 (of its problem) on purpose
        const bool vi = tb.pose_slot[src] >= lo, vj = tb.pose_slot[j] >= lo;
        const int ri = tb.pose_row[src], rj = tb.pose_row[j];
        float y[6];
#pragma unroll
        for (int c = 0; c < 6; c++) {
            float s = vj ? a.dx[6 * rj + c] : 0.0f;
            if (vi) {
#pragma unroll
                for (int r = 0; r < 6; r++) s = fmaf(G[6 * r + c], a.dx[6 * ri + r], s);
            }
            y[c] = s;
        }
#pragma unroll
        for (int c = 0; c < 6; c++) ys[m * 8 + c] = y[c];
    }
    const float fx = __ldg(a.intr + 0), fy = __ldg(a.intr + 1), cx = __ldg(a.intr + 2), cy = __ldg(a.intr + 3);
    const int px0 = tile * TILE + tid * PPT;
    const bool inb = px0 < HW;
    float xn[PPT], yn[PPT], h[PPT], acc[PPT];
    if (inb) load_px<PPT>(a.disps + (size_t)src * HW + px0, h);
#pragma unroll
    for (int s = 0; s < PPT; s++) {
        const int px = px0 + s;
        const int row = px / tb.wd;
        const int col = px - row * tb.wd;
        xn[s] = __fdiv_rn((float)col - cx, fx);
        yn[s] = __fdiv_rn((float)row - cy, fy);
        if (!inb) h[s] = 0.0f;
        acc[s] = 0.0f;
    }
    __syncthreads();
    if (!inb) return;

    for (int m = 0; m < d; m++) {
        const float *c = ec + m * kEcStride;
        if (c[12] != 0.0f) continue;  // stereo edge: u = 0
        const int e = __float_as_int(c[13]);
        const float *y = ys + m * 8;
        float wu[PPT], wv[PPT];
        const size_t base = (size_t)e * 2 * HW + px0;
        load_px<PPT>(a.weights + base, wu);
        load_px<PPT>(a.weights + base + HW, wv);
#pragma unroll
        for (int s = 0; s < PPT; s++) {
            const PixelGeom g = pixel_geom(c, xn[s], yn[s], h[s], fx, fy, a.opt);
            const float w_u = g.valid ? kWeightScale * wu[s] : 0.0f;
            const float w_v = g.valid ? kWeightScale * wv[s] : 0.0f;
            float u6[6];
            edge_u(g, w_u * g.Jzu * fx, w_v * g.Jzv * fy, u6);
#pragma unroll
            for (int r = 0; r < 6; r++) acc[s] = fmaf(u6[r], y[r], acc[s]);
        }
    }
    if (a.ufbuf) {  // focal step: dz -= Q u_f df
        float uf[PPT];
        load_px<PPT>(a.ufbuf + (size_t)k * HW + px0, uf);
        const float df = a.dx[a.focal_row];
#pragma unroll
        for (int s = 0; s < PPT; s++) acc[s] = fmaf(uf[s], df, acc[s]);
    }
    float q[PPT], qw[PPT], dz[PPT], hn[PPT];
    load_px<PPT>(a.qbuf + (size_t)k * HW + px0, q);
    load_px<PPT>(a.qwbuf + (size_t)k * HW + px0, qw);
#pragma unroll
    for (int s = 0; s < PPT; s++) {
        dz[s] = fmaf(-q[s], acc[s], qw[s]);
        if (dz[s] > a.opt.dz_max) dz[s] = 0.0f;  // retractor.py:41 (off by default)
        hn[s] = h[s] + dz[s];
    }
    store_px<PPT>(a.disps + (size_t)src * HW + px0, hn);
    store_px<PPT>(a.dz_out + (size_t)k * HW + px0, dz);
}

// =================================================================================================
// Focal pass (optimize_focal): everything that involves J_f = d(projection)/d(focal), for one (frame, 256-pixel tile).
// With xn = (col - cx)/fx, yn = (row - cy)/fy and X_j = R (xn, yn, 1) + h t:
//   J_f = jscale * [ X + fx d (g_x - X g_z) ;  Y + fy d (g_y - Y g_z) ],   g = -R (xn/fx, yn/fy, 0)
// i.e. the target camera's d(proj)/df plus the source camera's d(iproj)/df pushed through the transform
// (PinholeCameraModel.iproj_disp / proj_points, cameras.py:153-159,201-205; iproj_i_proj_j_disp, geom.py:282-287;
// both index the same intrinsics row, so the two Jacobians add, terms.py:217-228).
// Loop 1 (reads targets + weights): per edge J_j^T w J_f (6), J_f^T w J_f, J_f^T w r; per pixel u_f = sum_e J_f^T w J_z.
// Loop 2 (weights only, full BA): per edge sum_px Q u_f u_m (6); per tile sum_px Q u_f^2, sum_px Q w u_f.
__global__ void __launch_bounds__(kFocalNT) focal_kernel(const FocalArgs a) {
    constexpr int NW = kFocalNT / 32;
    extern __shared__ __align__(16) float smem[];
    const Tables &tb = a.tb;
    const int tile = blockIdx.x;
    const int k = tb.k_lo + blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int src = tb.kx[k];
    const int s0 = tb.fptr[k];
    const int d = tb.fptr[k + 1] - s0;
    if (d == 0) return;
    const int HW = tb.HW;
    float *ec = smem;                         // [d][16]
    float *red = ec + d * kEcStride;          // [d][NW][16]
    float *red2 = red + d * NW * kFocalStride;  // [NW][2]
    for (int m = tid; m < d; m += kFocalNT) {
        const int e = tb.fedge[s0 + m];
        RelPose<float> rp;
        relative_pose<float>(a.poses, src, tb.e_jj[e], rp);
        write_edge_consts(ec + m * kEcStride, rp, edge_row(tb, s0 + m, e));
    }
    for (int idx = tid; idx < d * NW * kFocalStride; idx += kFocalNT) red[idx] = 0.0f;
    const float fx = __ldg(a.intr + 0), fy = __ldg(a.intr + 1), cx = __ldg(a.intr + 2), cy = __ldg(a.intr + 3);
    const float js = a.opt.focal_jscale;
    const int px = tile * kFocalNT + tid;
    const bool inb = px < HW;
    const float h = inb ? __ldg(a.disps + (size_t)src * HW + px) : 0.0f;
    const int row = px / tb.wd, col = px - row * tb.wd;
    const float xn = __fdiv_rn((float)col - cx, fx), yn = __fdiv_rn((float)row - cy, fy);
    const float xnf = __fdiv_rn(xn, fx), ynf = __fdiv_rn(yn, fy);
    __syncthreads();

    float uf = 0.0f;
    for (int m = 0; m < d; m++) {
        const float *c = ec + m * kEcStride;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = 0.0f;
        if (c[12] == 0.0f && inb) {
            const int e = __float_as_int(c[13]);
            const size_t base = (size_t)e * 2 * HW + px;
            const float tu = __ldg(a.targets + base), tv = __ldg(a.targets + base + HW);
            const float wu = __ldg(a.weights + base), wv = __ldg(a.weights + base + HW);
            const PixelGeom g = pixel_geom(c, xn, yn, h, fx, fy, a.opt);
            const float w_u = g.valid ? kWeightScale * wu : 0.0f;
            const float w_v = g.valid ? kWeightScale * wv : 0.0f;
            const float ru = (tu - cx) - fx * g.X, rv = (tv - cy) - fy * g.Y;
            const float gx = -fmaf(c[0], xnf, c[1] * ynf), gy = -fmaf(c[3], xnf, c[4] * ynf), gz = -fmaf(c[6], xnf, c[7] * ynf);
            const float Jfu = js * fmaf(fx * g.dinv, fmaf(-g.X, gz, gx), g.X);
            const float Jfv = js * fmaf(fy * g.dinv, fmaf(-g.Y, gz, gy), g.Y);
            const float XY = g.X * g.Y;
            const float A0 = fx * g.a, A2 = -g.X * A0, A3 = -fx * XY, A4 = fx * fmaf(g.X, g.X, 1.0f), A5 = -fx * g.Y;
            const float B1 = fy * g.a, B2 = -g.Y * B1, B3 = -fy * fmaf(g.Y, g.Y, 1.0f), B4 = fy * XY, B5 = fy * g.X;
            const float pu = w_u * Jfu, pv = w_v * Jfv;
            v[0] = A0 * pu;
            v[1] = B1 * pv;
            v[2] = fmaf(A2, pu, B2 * pv);
            v[3] = fmaf(A3, pu, B3 * pv);
            v[4] = fmaf(A4, pu, B4 * pv);
            v[5] = fmaf(A5, pu, B5 * pv);
            v[6] = fmaf(pu, Jfu, pv * Jfv);
            v[7] = fmaf(pu, ru, pv * rv);
            uf = fmaf(w_u * g.Jzu, Jfu, fmaf(w_v * g.Jzv, Jfv, uf));
        }
        const float tot = warp_transpose_reduce<8>(v, lane);
        if (lane < 8) red[(m * NW + warp) * kFocalStride + lane] = tot;
    }
    if (!a.motion_only) {
        const float q = inb ? __ldg(a.qbuf + (size_t)k * HW + px) : 0.0f;
        const float qw = inb ? __ldg(a.qwbuf + (size_t)k * HW + px) : 0.0f;
        if (inb) a.ufbuf[(size_t)k * HW + px] = uf;
        const float quf = q * uf;
        float s2[2] = {quf * uf, qw * uf};
        const float t2 = warp_transpose_reduce<2>(s2, lane);
        if (lane < 2) red2[warp * 2 + lane] = t2;
        for (int m = 0; m < d; m++) {
            const float *c = ec + m * kEcStride;
            float y[8];
#pragma unroll
            for (int r = 0; r < 8; r++) y[r] = 0.0f;
            if (c[12] == 0.0f && inb) {
                const int e = __float_as_int(c[13]);
                const size_t base = (size_t)e * 2 * HW + px;
                const float wu = __ldg(a.weights + base), wv = __ldg(a.weights + base + HW);
                const PixelGeom g = pixel_geom(c, xn, yn, h, fx, fy, a.opt);
                const float w_u = g.valid ? kWeightScale * wu : 0.0f;
                const float w_v = g.valid ? kWeightScale * wv : 0.0f;
                float u6[6];
                edge_u(g, w_u * g.Jzu * fx, w_v * g.Jzv * fy, u6);
#pragma unroll
                for (int r = 0; r < 6; r++) y[r] = u6[r] * quf;
            }
            const float tot = warp_transpose_reduce<8>(y, lane);
            if (lane < 6) red[(m * NW + warp) * kFocalStride + 8 + lane] = tot;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < d * 14; idx += kFocalNT) {
        const int m = idx / 14, r = idx - m * 14;
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < NW; w++) s += red[(m * NW + w) * kFocalStride + r];
        a.fpart[((size_t)(s0 + m) * a.ntile_f + tile) * kFocalStride + r] = s;
    }
    if (!a.motion_only && tid < 2) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < NW; w++) s += red2[w * 2 + tid];
        a.ffpart[((size_t)k * a.ntile_f + tile) * 2 + tid] = s;
    }
}

// =================================================================================================
// Stage 4b: pose retraction T <- exp(xi) T without quaternion renormalisation
// (pose_retr_kernel / retrSE3 / expSE3 / expSO3, geom_kernels.cu:116-177,882-931).
__global__ void pose_retr_kernel(float *__restrict__ poses, const float *__restrict__ dx, const int *__restrict__ pose_row,
                                 int n_poses, int renorm, float *intr, int focal_row, float focal_jscale) {
    const int kk = blockIdx.x * blockDim.x + threadIdx.x;
    if (intr && kk == 0) {  // IntrinsicsRetractor.oplus for a pinhole camera: fx, fy += df (retractor.py:55-60)
        const float df = dx[focal_row] * focal_jscale;
        intr[0] += df;
        intr[1] += df;
    }
    if (kk >= n_poses) return;
    const int row = pose_row[kk];
    if (row < 0) return;
    float xi[6];
#pragma unroll
    for (int n = 0; n < 6; n++) xi[n] = dx[6 * row + n];
    const float t[3] = {poses[7 * kk], poses[7 * kk + 1], poses[7 * kk + 2]};
    const float q[4] = {poses[7 * kk + 3], poses[7 * kk + 4], poses[7 * kk + 5], poses[7 * kk + 6]};
    // expSO3
    const float *phi = xi + 3;
    const float th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
    const float th4 = th2 * th2;
    const float th = sqrtf(th2);
    float imag, real;
    if (th2 < 1e-8f) {
        imag = 0.5f - (1.0f / 48.0f) * th2 + (1.0f / 3840.0f) * th4;
        real = 1.0f - (1.0f / 8.0f) * th2 + (1.0f / 384.0f) * th4;
    } else {
        imag = sinf(0.5f * th) / th;
        real = cosf(0.5f * th);
    }
    const float dq[4] = {imag * phi[0], imag * phi[1], imag * phi[2], real};
    // expSE3 translation
    float tau[3] = {xi[0], xi[1], xi[2]};
    float dt[3] = {tau[0], tau[1], tau[2]};
    if (th > 1e-4f) {
        const float ca = (1.0f - cosf(th)) / th2;
        const float c1[3] = {phi[1] * tau[2] - phi[2] * tau[1], phi[2] * tau[0] - phi[0] * tau[2], phi[0] * tau[1] - phi[1] * tau[0]};
        const float cb = (th - sinf(th)) / (th * th2);
        const float c2[3] = {phi[1] * c1[2] - phi[2] * c1[1], phi[2] * c1[0] - phi[0] * c1[2], phi[0] * c1[1] - phi[1] * c1[0]};
#pragma unroll
        for (int n = 0; n < 3; n++) dt[n] += ca * c1[n] + cb * c2[n];
    }
    // retrSE3
    float q1[4];
    q1[0] = dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1];
    q1[1] = dq[3] * q[1] + dq[1] * q[3] + dq[2] * q[0] - dq[0] * q[2];
    q1[2] = dq[3] * q[2] + dq[2] * q[3] + dq[0] * q[1] - dq[1] * q[0];
    q1[3] = dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2];
    if (renorm) {  // lietorch's group product renormalises (so3.h:36-38); the reference CUDA BA does not (Q6)
        const float inv = rsqrtf(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]);
#pragma unroll
        for (int n = 0; n < 4; n++) q1[n] *= inv;
    }
    // actSO3(dq, t) (:69-78)
    const float uv[3] = {2.0f * (dq[1] * t[2] - dq[2] * t[1]), 2.0f * (dq[2] * t[0] - dq[0] * t[2]), 2.0f * (dq[0] * t[1] - dq[1] * t[0])};
    float t1v[3];
    t1v[0] = t[0] + dq[3] * uv[0] + (dq[1] * uv[2] - dq[2] * uv[1]) + dt[0];
    t1v[1] = t[1] + dq[3] * uv[1] + (dq[2] * uv[0] - dq[0] * uv[2]) + dt[1];
    t1v[2] = t[2] + dq[3] * uv[2] + (dq[0] * uv[1] - dq[1] * uv[0]) + dt[2];
#pragma unroll
    for (int n = 0; n < 3; n++) poses[7 * kk + n] = t1v[n];
#pragma unroll
    for (int n = 0; n < 4; n++) poses[7 * kk + 3 + n] = q1[n];
}

// =================================================================================================
// host-side launchers
static size_t lin_smem_bytes(int d, int NT, int PPT, bool motion) {
    const int NW = NT / 32, TILE = NT * PPT;
    size_t fl = (size_t)d * kEcStride + (((size_t)d * NW * kEdgeVals + 3) & ~(size_t)3);
    if (!motion) fl += (size_t)6 * d * TILE + 2 * TILE;
    else fl += 2 * TILE;
    return fl * sizeof(float);
}

template <int NT, int PPT, bool MOTION>
static cudaError_t launch_lin_t(const LinArgs &a, int nframes, int dmax, cudaStream_t st) {
    const size_t sm = lin_smem_bytes(dmax, NT, PPT, MOTION);
    auto kern = linearize_kernel<NT, PPT, MOTION>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (err != cudaSuccess) return err;
    dim3 grid(a.tb.ntile, nframes);
    kern<<<grid, NT, sm, st>>>(a);
    return cudaGetLastError();
}

bool tile_config(int HW, int dmax, bool motion, int &NT, int &PPT) {
    // largest tile whose staging buffer fits in shared memory; wide tiles only when they keep the grid full
    const size_t cap = 200 * 1024;
    const int cand[][2] = {{256, 2}, {256, 1}, {128, 1}, {64, 1}, {32, 1}};
    for (auto &c : cand) {
        if (HW % c[1] != 0) continue;
        if (lin_smem_bytes(dmax, c[0], c[1], motion) > cap) continue;
        if (c[1] == 2) continue;  // PPT=2 is opt-in via tile_config_wide (see launch_linearize)
        NT = c[0];
        PPT = c[1];
        return true;
    }
    return false;
}

cudaError_t launch_linearize(const LinArgs &a, int nframes, int dmax, bool motion, int NT, int PPT, cudaStream_t st) {
    if (nframes <= 0) return cudaSuccess;
#define VBA_LIN(NT_, PPT_)                                                                        \
    if (NT == NT_ && PPT == PPT_)                                                                 \
        return motion ? launch_lin_t<NT_, PPT_, true>(a, nframes, dmax, st)                       \
                      : launch_lin_t<NT_, PPT_, false>(a, nframes, dmax, st);
    VBA_LIN(256, 1)
    VBA_LIN(256, 2)
    VBA_LIN(128, 1)
    VBA_LIN(64, 1)
    VBA_LIN(32, 1)
#undef VBA_LIN
    return cudaErrorInvalidValue;
}

cudaError_t launch_frame_reduce(const ReduceArgs &a, int nframes, int dmax, cudaStream_t st) {
    if (nframes <= 0) return cudaSuccess;
    size_t sm = (size_t)dmax * (36 + 36 + 27 + 6 + (a.fpart ? 20 : 0)) * sizeof(double) + (size_t)(dmax + 2) * sizeof(int) + 16;
    const size_t msc_bytes = (size_t)dmax * (dmax + 1) / 2 * 36 * sizeof(double);
    ReduceArgs b = a;
    b.msc_smem = (sm + msc_bytes <= 96 * 1024) ? 1 : 0;  // backend degrees (~10-20 edges per frame) fit; hubs fall back to global
    if (b.msc_smem) sm += msc_bytes;
    if (sm > 200 * 1024) {  // hub frames beyond shared memory: everything per edge moves to the global scratch
        if (!a.fscratch) return cudaErrorInvalidValue;
        b.fs_global = 1;
        sm = 64;
    }
    cudaError_t err = cudaFuncSetAttribute(frame_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (err != cudaSuccess) return err;
    frame_reduce_kernel<<<nframes, 256, sm, st>>>(b);
    return cudaGetLastError();
}

template <int NT, int PPT>
static cudaError_t launch_back_t(const BackArgs &a, int nframes, int dmax, cudaStream_t st) {
    const size_t sm = (size_t)dmax * (kEcStride + 8) * sizeof(float) + 16;
    auto kern = backsub_kernel<NT, PPT>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (err != cudaSuccess) return err;
    constexpr int TILE = NT * PPT;
    dim3 grid((a.tb.HW + TILE - 1) / TILE, nframes);
    kern<<<grid, NT, sm, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_backsub(const BackArgs &a, int nframes, int dmax, cudaStream_t st) {
    if (nframes <= 0) return cudaSuccess;
    if (a.tb.HW % 4 == 0) return launch_back_t<128, 4>(a, nframes, dmax, st);
    if (a.tb.HW % 2 == 0) return launch_back_t<128, 2>(a, nframes, dmax, st);
    return launch_back_t<128, 1>(a, nframes, dmax, st);
}

cudaError_t launch_pose_retr(float *poses, const float *dx, const int *pose_row, int n_poses, int renorm, float *intr,
                             int focal_row, float focal_jscale, cudaStream_t st) {
    if (n_poses <= 0) return cudaSuccess;
    pose_retr_kernel<<<(n_poses + 127) / 128, 128, 0, st>>>(poses, dx, pose_row, n_poses, renorm, intr, focal_row, focal_jscale);
    return cudaGetLastError();
}

__global__ void add_scalar_kernel(double *p, double v) { *p += v; }
cudaError_t launch_add_scalar(double *p, double v, cudaStream_t st) {
    add_scalar_kernel<<<1, 1, 0, st>>>(p, v);
    return cudaGetLastError();
}

int focal_tiles(int HW) { return (HW + kFocalNT - 1) / kFocalNT; }

cudaError_t launch_focal(const FocalArgs &a, int nframes, int dmax, cudaStream_t st) {
    if (nframes <= 0) return cudaSuccess;
    const size_t sm = ((size_t)dmax * (kEcStride + (kFocalNT / 32) * kFocalStride) + 2 * (kFocalNT / 32)) * sizeof(float) + 16;
    cudaError_t err = cudaFuncSetAttribute(focal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (err != cudaSuccess) return err;
    dim3 grid(a.ntile_f, nframes);
    focal_kernel<<<grid, kFocalNT, sm, st>>>(a);
    return cudaGetLastError();
}

}  // namespace vba
