// Blackwell-native linearisation + Schur Gram ("lin3"): the same outputs as linearize2_kernel (epart / gpart / Q / Qw;
// reference: projective_transform_kernel, accum_kernel, EEt6x6_kernel, Ev6x1_kernel, csrc/slam_ext/geom_kernels.cu:178-432,
// 863-880,994-1080) for source frames with 1..10 outgoing edges, as a persistent, warp-specialised pipeline.  One CTA
// per SM walks a static list of items (frame, 512-pixel chunk); every role free-runs over the item's 64-pixel sub-tiles and
// meets the others only through mbarriers:
//
//   10 J warps      warp w owns edge w of the frame.  It fetches its edge's targets/weights rows itself: cp.async.bulk
//                   (TMA, 1-D, 1 KB per row and 256-pixel stage) into a private double-buffered stage, completion on an
//                   mbarrier.  Per sub-tile: projective transform, residuals, J_j, J_z; H_jj / v_j / energy stay in
//                   registers for the whole item (one cross-lane reduction per item -> epart); u (6 per edge-pixel) and
//                   the edge's C, w contributions go to a raw staging buffer.  One sub-tile later, when Q is known, the
//                   warp rewrites ITS rows as sqrt(Q)*u split into (hi, lo) TF32 planes in the tensor core's K-major
//                   128-byte-swizzled operand layout (software pipeline: compute(t), then split(t-1)).
//   2 Q warps       one pixel per lane: C, w summed over the edges in edge order, damping / prior -> Q, Q*w (global) and
//                   sqrt(Q), sqrt(Q)*w (shared).  Lane 0 of the first one also issues the MMAs of the previous sub-tile:
//                   tcgen05.mma kind::tf32, D = [hi; lo] hi^T (M = 128: the hi.hi and lo.hi products in one instruction),
//                   accumulators in tensor memory.
//   4 flush warps   tcgen05.ld of the accumulators after every sub-tile, round-to-nearest fp32 sums in registers; per item
//                   Y = hh + lh + lh^T and the sign fix-up, written to gpart.
//
// Y = sum_px Q u u^T = hh + lh + lh^T (the lo.lo term, 2^-22 relative, is dropped).  The tensor core truncates its fp32
// accumulator (measured: about -1 ulp of the running sum per K = 8 step, scripts/tc_gram_probe.cu), and A - S cancels
// in the gauge directions, so the in-unit accumulation chains are kept to TILE / 8 / NSETS steps: the K steps of a
// sub-tile rotate over NSETS independent accumulators (which also hides the ~120-cycle dependent-MMA latency) and the
// flush warps add them up in round-to-nearest fp32 after every sub-tile.
#include "ba_common.cuh"
#include "ba_launch.h"
#include "sm100_async.cuh"

namespace vba {
using namespace sm100;

namespace {

constexpr int kNJ = kLin3MaxDeg;  // J warps: warp w owns edge w of the frame (<= kLin3MaxDeg edges)
constexpr int kNQ = 2;            // of the flush warps, the first two also do the per-pixel Q pass (one pixel per lane)
constexpr int kNF = 4;            // flush warps: one per tensor-memory lane quarter
constexpr int kThreads = 32 * (1 + kNF + kNJ);  // warp 0: MMA issuer
constexpr int kRawRows = 8;       // raw staging rows per edge: u (6), c, w
constexpr int TILE = 64;          // pixels per sub-tile = K of one MMA batch
constexpr int STAGE_PX = 256;     // pixels per TMA stage (4 sub-tiles)
constexpr int KSTEPS = TILE / 8;
constexpr int NSETS = 2;          // accumulator sets the K steps rotate over
constexpr int ACC_COLS = 64;      // columns of one accumulator set (N <= 64)
constexpr int TM_BUF = NSETS * ACC_COLS;

// pair index -> (m, mp), m <= mp, p = mp (mp + 1) / 2 + m   (d <= 21 -> 231 pairs)
__constant__ unsigned char c_pair_m[256], c_pair_mp[256];

// optional pipeline trace (build with -DVBA_LIN3_TRACE): clock stamps of CTA 0, [role][event], dumped by the launcher
#ifdef VBA_LIN3_TRACE
__device__ long long *g_trace = nullptr;
constexpr int kTraceLen = 4096;
#define TR(role, cond, tag)                                                                                            \
    do {                                                                                                               \
        if (blockIdx.x == 0 && (cond) && g_trace && tr_n + 2 < kTraceLen) {                                            \
            g_trace[(size_t)(role)*kTraceLen + 1 + tr_n] = ((long long)(tag) << 48) | (clock64() & 0xffffffffffffLL); \
            tr_n++;                                                                                                    \
            g_trace[(size_t)(role)*kTraceLen] = tr_n;                                                                  \
        }                                                                                                              \
    } while (0)
#else
#define TR(role, cond, tag) \
    do {                    \
    } while (0)
#endif

struct Smem3 {
    static constexpr size_t op_bytes = (size_t)(TILE / 32) * 128 * 128;  // one operand buffer: [hi(64) ; lo(64)] rows x TILE
    static constexpr size_t raw_bytes = (size_t)kRawRows * kNJ * TILE * 4;
    static constexpr size_t stage_bytes = (size_t)4 * STAGE_PX * 4;       // tu, tv, wu, wv rows of one edge
    static constexpr size_t off_op = 0;
    static constexpr size_t off_raw = off_op + 2 * op_bytes;
    static constexpr size_t off_ring = off_raw + 2 * raw_bytes;
    static constexpr size_t off_ys = off_ring + (size_t)kNJ * 2 * stage_bytes;
    static constexpr size_t off_ec = off_ys + (size_t)64 * 65 * 4;
    static constexpr size_t off_sq = off_ec + (size_t)kNJ * 16 * 8;
    static constexpr size_t off_bar = off_sq + (size_t)2 * 2 * TILE * 4;
    static constexpr int n_bar = 2 * kNJ + 5 * 2;
    static constexpr size_t off_tmem = off_bar + (size_t)n_bar * 8;
    static constexpr size_t total = off_tmem + 16;
};

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
// pose-disparity vector u_hat -> rawp[row * TILE] (rows 0..5 u_hat, 6 C, 7 w).
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
        *reinterpret_cast<float2 *>(rawp + 6 * TILE) = ffma2(wJu, Jzu, fmul2(wJv, Jzv));  // :325,364
        *reinterpret_cast<float2 *>(rawp + 7 * TILE) = ffma2(wJu, ru, fmul2(wJv, rv));    // :326,365
    }
    if (stereo) {  // stereo edges contribute nothing beyond C and w (:329,367); uniform over the warp
        if (FULL) {
#pragma unroll
            for (int i = 0; i < 6; i++) *reinterpret_cast<float2 *>(rawp + i * TILE) = make_float2(0.f, 0.f);
        }
        return;
    }
    const float2 P2 = fmul2(X, aa), P3 = fmul2(X, Y), P4 = ffma2(X, X, k.one2), P5 = Y;
    const float2 Q2 = fmul2(Y, aa), Q3 = ffma2(Y, Y, k.one2), Q4 = P3, Q5 = fmul2(X, k.m1);
    if (FULL) {  // u = wu Jzu A + wv Jzv B (E_ij column, :351,385)
        *reinterpret_cast<float2 *>(rawp + 0 * TILE) = fmul2(wJu, aa);
        *reinterpret_cast<float2 *>(rawp + 1 * TILE) = fmul2(wJv, aa);
        *reinterpret_cast<float2 *>(rawp + 2 * TILE) = ffma2(wJu, P2, fmul2(wJv, Q2));
        *reinterpret_cast<float2 *>(rawp + 3 * TILE) = ffma2(wJu, P3, fmul2(wJv, Q3));
        *reinterpret_cast<float2 *>(rawp + 4 * TILE) = ffma2(wJu, P4, fmul2(wJv, Q4));
        *reinterpret_cast<float2 *>(rawp + 5 * TILE) = ffma2(wJu, P5, fmul2(wJv, Q5));
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

}  // namespace

// per-edge constants of the frames the pipeline covers: R (9), t (3), -t2, stereo flag -> econst[e][16]
// (relSE3 / the stereo special case, geom_kernels.cu:105-113,219-230)
__global__ void lin3_prep_kernel(const Lin3Item *__restrict__ items, int nframes, int nchunk, const float *__restrict__ poses,
                                 const int *__restrict__ e_jj, float *__restrict__ econst) {
    const int f = blockIdx.x * (blockDim.x / 16) + threadIdx.x / 16, m = threadIdx.x % 16;
    if (f >= nframes) return;
    const Lin3Item it = items[(size_t)f * nchunk];
    if (m >= it.d) return;
    const int e = it.edge[m];
    RelPose<float> rp;
    relative_pose<float>(poses, it.src, e_jj[e], rp);
    float *c = econst + (size_t)e * 16;
#pragma unroll
    for (int q = 0; q < 9; q++) c[q] = rp.R[q];
#pragma unroll
    for (int q = 0; q < 3; q++) c[9 + q] = rp.t[q];
    c[12] = -rp.t[2];
    c[13] = rp.stereo ? 1.0f : 0.0f;
    c[14] = c[15] = 0.0f;
}

// ================================================================================================================
__global__ void __launch_bounds__(kThreads, 1) lin3_kernel(const LinArgs a, const Lin3Item *__restrict__ items, int nitems, int nsub,
                                                           const float *__restrict__ econst) {
    using L = Smem3;
    extern __shared__ __align__(1024) unsigned char smem[];  // operand buffers first: the swizzle needs 1024-byte alignment
    unsigned char *op = smem + L::off_op;
    float *raw = reinterpret_cast<float *>(smem + L::off_raw);
    unsigned char *ring = smem + L::off_ring;
    float *Ys = reinterpret_cast<float *>(smem + L::off_ys);  // [64][65]
    float2 *ecs = reinterpret_cast<float2 *>(smem + L::off_ec);
    float *sqs = reinterpret_cast<float *>(smem + L::off_sq);  // [2][2][TILE]: sqrt(Q), sqrt(Q) w, double-buffered
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::off_bar);
    uint64_t *ring_full = bars;                   // [kNJ][2]
    uint64_t *cw_ready = bars + 2 * kNJ;          // [2] J -> Q: the edges' C, w rows of the sub-tile are in `raw`
    uint64_t *q_ready = cw_ready + 2;             // [2] Q -> J: sqrt(Q), sqrt(Q) w of the sub-tile
    uint64_t *split_ready = q_ready + 2;          // [2] J -> MMA: operand buffer written
    uint64_t *mma_done = split_ready + 2;         // [2] MMA -> flush (accumulators ready) and -> J (operand buffer free)
    uint64_t *tmem_free = mma_done + 2;           // [2] flush -> MMA
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::off_tmem);

    const Tables &tb = a.tb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HW = tb.HW;
    const int spi = nsub * TILE / STAGE_PX;  // TMA stages per item
    int n_my = 0;                            // items of this CTA
    if ((int)blockIdx.x < nitems) n_my = (nitems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
    const uint32_t G = (uint32_t)n_my * nsub;  // sub-tiles of this CTA
#ifdef VBA_LIN3_TRACE
    int tr_n = 0;
#endif

    if (tid == 0) {
        for (int i = 0; i < 2 * kNJ; i++) mbar_init(ring_full + i, 1);
        for (int i = 0; i < 2; i++) {
            mbar_init(cw_ready + i, kNJ);
            mbar_init(q_ready + i, kNQ);
            mbar_init(split_ready + i, kNJ);
            mbar_init(mma_done + i, 1);
            mbar_init(tmem_free + i, kNF);
        }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    // unused operand rows must hold finite numbers (they only feed accumulator entries nobody reads)
    for (int i = tid; i < (int)(2 * L::op_bytes / 16); i += kThreads) reinterpret_cast<float4 *>(op)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ============================================================ MMA issuer (one thread)
        if (lane == 0) {
            uint32_t g = 0;
            for (int it = 0; it < n_my; it++) {
                const int d = items[blockIdx.x + (size_t)it * gridDim.x].d;
                const uint32_t idesc = make_idesc_tf32(128, ((6 * d + 1 + 15) / 16) * 16);
                for (int t = 0; t < nsub; t++, g++) {
                    const int pb = g & 1;
                    const uint32_t sbase = smem_u32(op) + pb * (uint32_t)L::op_bytes;
                    TR(0, true, 1);
                    mbar_wait_relaxed(split_ready + pb, (g >> 1) & 1);
                    mbar_wait_relaxed(tmem_free + pb, ((g >> 1) & 1) ^ 1);
                    TR(0, true, 2);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ks++) {
                        const uint64_t desc = make_desc_k_sw128(sbase + (ks >> 2) * (128 * 128) + (ks & 3) * 32);
                        mma_tf32(tmem + pb * TM_BUF + (ks % NSETS) * ACC_COLS, desc, desc, idesc, ks >= NSETS);
                    }
                    mma_commit(mma_done + pb);
                    TR(0, true, 3);
                }
            }
        }
        __syncwarp();
    } else if (warp < 1 + kNF) {
        // ============================================================ flush warps (the first kNQ of them also do the Q pass)
        const int pt = tid - 32;               // 0..127
        const int q = warp & 3;                // TMEM lane quarter this warp may read
        const uint32_t tm_lane = (uint32_t)(32 * q) << 16;
        const bool qrole = pt < 32 * kNQ;      // this lane owns pixel pt of every sub-tile
        float hpx = 0.f, ds = 0.f, et = 0.f;
        Lin3Item cur = items[blockIdx.x < (unsigned)nitems ? blockIdx.x : 0], prev = cur;
        if (G > 0 && qrole) {
            const int px = cur.px0 + pt;
            hpx = __ldg(a.disps + (size_t)cur.src * HW + px);
            ds = __ldg(a.dsens + (size_t)cur.src * HW + px);
            et = __ldg(a.eta + (size_t)cur.k * HW + px);
        }
        float acc[ACC_COLS];
#pragma unroll
        for (int c = 0; c < ACC_COLS; c++) acc[c] = 0.0f;
        auto flush = [&](uint32_t gg, int N) {
            const int buf = gg & 1;
            TR(1, pt == 0, 4);
            mbar_wait_relaxed(mma_done + buf, (gg >> 1) & 1);
            TR(1, pt == 0, 5);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < ACC_COLS; c0 += 16) {
                if (c0 < N) {
                    float v[NSETS][16];
#pragma unroll
                    for (int s = 0; s < NSETS; s++) tmem_ld16(tmem + tm_lane + buf * TM_BUF + s * ACC_COLS + c0, v[s]);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        float sum = v[0][i];
#pragma unroll
                        for (int s = 1; s < NSETS; s++) sum += v[s][i];
                        acc[c0 + i] += sum;
                    }
                }
            }
            tc_fence_before();
            warp_arrive(tmem_free + buf, lane);
            TR(1, pt == 0, 6);
        };
        // item epilogue: Y = hh + lh + lh^T through shared memory, signs, -> gpart; clears the accumulators
        auto epilogue = [&](const Lin3Item &im) {
            const int d = im.d, R = 6 * d + 1;
            const int r = 32 * (q & 1) + lane;
            if (q < 2) {
#pragma unroll
                for (int c = 0; c < ACC_COLS; c++) Ys[r * 65 + c] = acc[c];
            }
            named_bar(1, 32 * kNF);
            if (q >= 2) {
#pragma unroll
                for (int c = 0; c < ACC_COLS; c++) Ys[r * 65 + c] += acc[c];
            }
            named_bar(1, 32 * kNF);
            if (q >= 2) {
#pragma unroll
                for (int c = 0; c < ACC_COLS; c++) Ys[c * 65 + r] += acc[c];
            }
            named_bar(1, 32 * kNF);
            const int npairs = d * (d + 1) / 2;
            const int rec = npairs * 36 + 6 * d;
            float *gp = a.gpart + tb.gbase[im.k] + (size_t)im.chunk * rec;
            for (int idx = pt; idx < rec; idx += 32 * kNF) {
                float val;
                if (idx < npairs * 36) {
                    const int p = idx / 36, rc = idx - p * 36;
                    const int rr = rc / 6, cc = rc - rr * 6;
                    const int m = c_pair_m[p], mp = c_pair_mp[p];
                    val = Ys[(6 * m + rr) * 65 + 6 * mp + cc] * usign(rr) * usign(cc);
                } else {
                    const int j = idx - npairs * 36;
                    val = Ys[j * 65 + (R - 1)] * usign(j % 6);
                }
                gp[idx] = val;
            }
            named_bar(1, 32 * kNF);
#pragma unroll
            for (int c = 0; c < ACC_COLS; c++) acc[c] = 0.0f;
        };
        uint32_t g = 0;
        for (int it = 0; it < n_my; it++) {
            Lin3Item nxt = cur;
            if (it + 1 < n_my) nxt = items[blockIdx.x + (size_t)(it + 1) * gridDim.x];
            const int flags = a.opt.frame_flags ? a.opt.frame_flags[cur.k] : 0;
            for (int t = 0; t < nsub; t++, g++) {
                const int buf = g & 1;
                if (qrole) {
                    const float h0 = hpx, ds0 = ds, et0 = et;
                    const int px = cur.px0 + t * TILE + pt;
                    if (g + 1 < G) {  // next sub-tile's pixel data leaves now
                        const bool same = t + 1 < nsub;
                        const int srcn = same ? cur.src : nxt.src, kn = same ? cur.k : nxt.k;
                        const int pxn = same ? px + TILE : nxt.px0 + pt;
                        hpx = __ldg(a.disps + (size_t)srcn * HW + pxn);
                        ds = __ldg(a.dsens + (size_t)srcn * HW + pxn);
                        et = __ldg(a.eta + (size_t)kn * HW + pxn);
                    }
                    const float *rb = raw + (size_t)buf * (L::raw_bytes / 4);
                    TR(1, pt == 0, 1);
                    mbar_wait_relaxed(cw_ready + buf, (g >> 1) & 1);
                    TR(1, pt == 0, 2);
                    // disparity block of this pixel: C, w summed over the edges in edge order, damping / prior (:1359-1370), Q = 1/C
                    float C = 0.0f, W = 0.0f;
                    for (int m = 0; m < cur.d; m++) {
                        C += rb[(size_t)(kRawRows * m + 6) * TILE + pt];
                        W += rb[(size_t)(kRawRows * m + 7) * TILE + pt];
                    }
                    float Q, wz;
                    disparity_block(C, W, h0, ds0, et0, flags, a.opt, Q, wz);
                    a.qbuf[(size_t)cur.k * HW + px] = Q;
                    a.qwbuf[(size_t)cur.k * HW + px] = Q * wz;
                    const float sq = sqrtf(Q);
                    sqs[(buf * 2 + 0) * TILE + pt] = sq;
                    sqs[(buf * 2 + 1) * TILE + pt] = sq * wz;
                    warp_arrive(q_ready + buf, lane);
                    TR(1, pt == 0, 3);
                }
                if (g > 0) {
                    const Lin3Item &fi = (t == 0) ? prev : cur;
                    flush(g - 1, ((6 * fi.d + 1 + 15) / 16) * 16);
                    if (t == 0) epilogue(prev);
                }
            }
            prev = cur;
            cur = nxt;
        }
        if (G > 0) {
            flush(G - 1, ((6 * prev.d + 1 + 15) / 16) * 16);
            epilogue(prev);
        }
    } else {
        // ============================================================ J warps
        const int jw = warp - (1 + kNF);
        const JConst kc = make_jconst(a);
        const int wd = tb.wd;
        float *myraw0 = raw + (size_t)kRawRows * jw * TILE;
        float2 *c = ecs + (size_t)jw * 16;
        unsigned char *mystage = ring + (size_t)jw * 2 * L::stage_bytes;
        uint64_t *myfull = ring_full + 2 * jw;
        const uint32_t nstage = (uint32_t)n_my * spi;

        // TMA stage sq of this CTA's sequence: rows tu, tv, wu, wv of this warp's edge over 256 pixels
        auto issue_stage = [&](uint32_t sq) {
            if (sq >= nstage) return;
            const Lin3Item &pi = items[blockIdx.x + (size_t)(sq / spi) * gridDim.x];
            if (pi.edge[jw] < 0) return;
            const int erow = edge_row(tb, pi.s0 + jw, pi.edge[jw]);
            const int b = sq & 1;
            if (lane == 0) mbar_arrive_expect_tx(myfull + b, (uint32_t)L::stage_bytes);
            __syncwarp();
            if (lane < 4) {
                const float *src = (lane < 2 ? a.targets : a.weights) + ((size_t)erow * 2 + (lane & 1)) * HW + pi.px0 + (sq % spi) * STAGE_PX;
                bulk_g2s(mystage + (size_t)b * L::stage_bytes + (size_t)lane * STAGE_PX * 4, src, STAGE_PX * 4, myfull + b);
            }
        };
        issue_stage(0);
        issue_stage(1);

        // split(t): this warp's rows of sub-tile gs, scaled by sqrt(Q), as (hi, lo) TF32 planes of operand buffer gs & 1
        auto split_pass = [&](uint32_t gs, bool act, int dd_) {
            const int buf = gs & 1;
            TR(2, jw == 0 && lane == 0, 5);
            mbar_wait(q_ready + buf, (gs >> 1) & 1);
            if (gs >= 2) mbar_wait(mma_done + buf, ((gs >> 1) - 1) & 1);  // the MMAs that read this buffer two sub-tiles ago are done
            TR(2, jw == 0 && lane == 0, 6);
            const float *sq = sqs + (size_t)(buf * 2) * TILE;
            const float *rbw = myraw0 + (size_t)buf * (L::raw_bytes / 4);
            unsigned char *opb = op + (size_t)buf * L::op_bytes;
            auto put = [&](int row, int quad, float4 v) {
                float4 hi, lo;
                split_tf32(v.x, hi.x, lo.x);
                split_tf32(v.y, hi.y, lo.y);
                split_tf32(v.z, hi.z, lo.z);
                split_tf32(v.w, hi.w, lo.w);
                unsigned char *blk = opb + (size_t)(quad >> 3) * (128 * 128) + sw128_offset(row, 4 * (quad & 7));
                *reinterpret_cast<float4 *>(blk) = hi;
                *reinterpret_cast<float4 *>(blk + 8 * 1024) = lo;  // row + 64
            };
            if (act) {
                // 6 rows x 16 pixel quads = 96 items over 32 lanes: lane handles quad (lane & 15), rows (lane >> 4) + 2 j
                const int quad = lane & 15;
                const float4 s4 = *reinterpret_cast<const float4 *>(sq + 4 * quad);
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const int r6 = (lane >> 4) + 2 * j;
                    const float4 u4 = *reinterpret_cast<const float4 *>(rbw + r6 * TILE + 4 * quad);
                    put(6 * jw + r6, quad, make_float4(u4.x * s4.x, u4.y * s4.y, u4.z * s4.z, u4.w * s4.w));
                }
            }
            if (jw == (dd_ % kNJ) && lane < 16) put(6 * dd_, lane, *reinterpret_cast<const float4 *>(sq + TILE + 4 * lane));  // row 6d: sqrt(Q) w
            fence_async_smem();
            warp_arrive(split_ready + buf, lane);
            TR(2, jw == 0 && lane == 0, 7);
        };

        uint32_t g = 0, sq_cur = 0, cons0 = 0, cons1 = 0;
        bool prev_active = false;
        int prev_d = 0;
        // per-item data (edge constants, first disparities) is fetched one item ahead
        Lin3Item cur = items[blockIdx.x < (unsigned)nitems ? blockIdx.x : 0];
        float cv = 0.0f;
        float2 hfirst = make_float2(0.f, 0.f);
        if (n_my > 0 && cur.edge[jw] >= 0) {
            cv = (lane < 14) ? __ldg(econst + (size_t)cur.edge[jw] * 16 + lane) : 0.0f;
            hfirst = __ldg(reinterpret_cast<const float2 *>(a.disps + (size_t)cur.src * HW + cur.px0 + 2 * lane));
        }
        for (int it = 0; it < n_my; it++) {
            const int d = cur.d, px0 = cur.px0, src = cur.src, chunk = cur.chunk, s0 = cur.s0;
            const bool active = cur.edge[jw] >= 0;
            const bool stereo = active && __shfl_sync(0xffffffffu, cv, 13) != 0.0f;
            if (active && lane < 14) c[lane] = splat2(cv);  // (the previous item's reads of c are behind its warp barriers)
            __syncwarp();
            float2 hnext = hfirst;
            if (it + 1 < n_my) {
                cur = items[blockIdx.x + (size_t)(it + 1) * gridDim.x];
                if (cur.edge[jw] >= 0) {
                    cv = (lane < 14) ? __ldg(econst + (size_t)cur.edge[jw] * 16 + lane) : 0.0f;
                    hfirst = __ldg(reinterpret_cast<const float2 *>(a.disps + (size_t)cur.src * HW + cur.px0 + 2 * lane));
                }
            }
            PxPos pos;
            pos.set(px0 + 2 * lane, wd);
            float2 hh[kEdgeVals];
#pragma unroll
            for (int i = 0; i < kEdgeVals; i++) hh[i] = make_float2(0.0f, 0.0f);
            const float2 *hrow = reinterpret_cast<const float2 *>(a.disps + (size_t)src * HW + px0 + 2 * lane);
            for (int st = 0; st < spi; st++, sq_cur++) {
                const int sb = sq_cur & 1;
                const float *stg = reinterpret_cast<const float *>(mystage + (size_t)sb * L::stage_bytes) + 2 * lane;
                if (active) {  // parity = completions of this buffer consumed so far (stages of edge-less items are never issued)
                    TR(2, jw == 0 && lane == 0, 1);
                    mbar_wait(myfull + sb, (sb ? cons1 : cons0) & 1);
                    TR(2, jw == 0 && lane == 0, 2);
                    if (sb) cons1++; else cons0++;
                }
#pragma unroll
                for (int part = 0; part < STAGE_PX / TILE; part++, g++) {
                    const int buf = g & 1;
                    const int t = st * (STAGE_PX / TILE) + part;
                    if (active) {
                        const float2 h = hnext;
                        if (t + 1 < nsub) hnext = __ldg(hrow + (t + 1) * (TILE / 2));
                        float2 xn, yn;
                        pos.normalised(kc, xn, yn);
                        pos.advance((float)TILE, kc.wdf);
                        const float2 tu_s = *reinterpret_cast<const float2 *>(stg + TILE * part);
                        const float2 tv_s = *reinterpret_cast<const float2 *>(stg + STAGE_PX + TILE * part);
                        const float2 wu_s = *reinterpret_cast<const float2 *>(stg + 2 * STAGE_PX + TILE * part);
                        const float2 wv_s = *reinterpret_cast<const float2 *>(stg + 3 * STAGE_PX + TILE * part);
                        j_step<true>(kc, c, stereo, xn, yn, h, tu_s, tv_s, wu_s, wv_s, hh, myraw0 + (size_t)buf * (L::raw_bytes / 4) + 2 * lane);
                    }
                    if (part == STAGE_PX / TILE - 1) {  // stage consumed (the warp barrier orders its reads before the refill)
                        __syncwarp();
                        issue_stage(sq_cur + 2);
                    }
                    TR(2, jw == 0 && lane == 0, 3);
                    warp_arrive(cw_ready + buf, lane);
                    if (g > 0) split_pass(g - 1, t == 0 ? prev_active : active, t == 0 ? prev_d : d);
                }
            }
            // per-(edge, chunk) record: this warp's edge, summed over its lanes; true signs, image units
            if (active) write_edge_record(hh, lane, a.epart + ((size_t)(s0 + jw) * tb.ntile + chunk) * kEdgeStride);
            prev_active = active;
            prev_d = d;
        }
        if (G > 0) split_pass(G - 1, prev_active, prev_d);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ================================================================================================================
// Motion-only linearisation (stages 1-2 alone: projective transform, J_j, per-edge H_jj / v_j / energy; no disparity block,
// no Schur Gram).  Edges do not couple here, so the work list is flat: unit = (edge slot, 512-pixel chunk), and every WARP
// of the persistent grid walks its own units: it fetches the unit's targets/weights rows with cp.async.bulk into a private
// double-buffered 256-pixel stage (refilled as soon as it has been read), keeps the 27 sums in registers for the whole
// unit and reduces them across lanes once.
constexpr int kMotionWarps = 16;

__global__ void edge_const_kernel(const int *__restrict__ fedge, const int *__restrict__ slot_src, int slot_lo, int nslots,
                                  const float *__restrict__ poses, const int *__restrict__ e_jj, float *__restrict__ econst) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const int e = fedge[slot_lo + s];
    RelPose<float> rp;
    relative_pose<float>(poses, slot_src[slot_lo + s], e_jj[e], rp);
    float *c = econst + (size_t)e * 16;
#pragma unroll
    for (int q = 0; q < 9; q++) c[q] = rp.R[q];
#pragma unroll
    for (int q = 0; q < 3; q++) c[9 + q] = rp.t[q];
    c[12] = -rp.t[2];
    c[13] = rp.stereo ? 1.0f : 0.0f;
    c[14] = c[15] = 0.0f;
}

__global__ void __launch_bounds__(32 * kMotionWarps, 1) lin3_motion_kernel(const LinArgs a, const int *__restrict__ slot_src, int slot_lo,
                                                                          int nslots, int nchunk, int nsub,
                                                                          const float *__restrict__ econst) {
    constexpr size_t stage_bytes = (size_t)4 * STAGE_PX * 4;
    extern __shared__ __align__(1024) unsigned char smem[];
    const Tables &tb = a.tb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HW = tb.HW, wd = tb.wd;
    unsigned char *mystage = smem + (size_t)warp * 2 * stage_bytes;
    float2 *c = reinterpret_cast<float2 *>(smem + (size_t)kMotionWarps * 2 * stage_bytes) + (size_t)warp * 16;
    uint64_t *myfull = reinterpret_cast<uint64_t *>(smem + (size_t)kMotionWarps * (2 * stage_bytes + 16 * 8)) + 2 * warp;
    if (lane == 0) {
        mbar_init(myfull, 1);
        mbar_init(myfull + 1, 1);
        mbar_fence_init();
    }
    __syncwarp();
    const int spi = nsub * TILE / STAGE_PX;
    const long long nunits = (long long)nslots * nchunk;
    const long long wstride = (long long)gridDim.x * kMotionWarps;
    const long long u0 = (long long)blockIdx.x * kMotionWarps + warp;
    const long long n_my = u0 < nunits ? (nunits - 1 - u0) / wstride + 1 : 0;
    const uint32_t nstage = (uint32_t)(n_my * spi);

    const JConst kc = make_jconst(a);

    auto issue_stage = [&](uint32_t sq) {
        if (sq >= nstage) return;
        const long long u = u0 + (long long)(sq / spi) * wstride;
        const int slot = slot_lo + (int)(u / nchunk), ch = (int)(u % nchunk);
        const int erow = edge_row(tb, slot, tb.fedge[slot]);
        const int b = sq & 1;
        if (lane == 0) mbar_arrive_expect_tx(myfull + b, (uint32_t)stage_bytes);
        __syncwarp();
        if (lane < 4) {
            const float *src = (lane < 2 ? a.targets : a.weights) + ((size_t)erow * 2 + (lane & 1)) * HW + (size_t)ch * nsub * TILE + (sq % spi) * STAGE_PX;
            bulk_g2s(mystage + (size_t)b * stage_bytes + (size_t)lane * STAGE_PX * 4, src, STAGE_PX * 4, myfull + b);
        }
    };
    issue_stage(0);
    issue_stage(1);

    // per-unit data (edge constants, first disparities) is fetched one unit ahead
    auto unit_of = [&](long long i, int &slot, int &ch, int &e, int &src) {
        const long long u = u0 + i * wstride;
        slot = slot_lo + (int)(u / nchunk), ch = (int)(u % nchunk);
        e = tb.fedge[slot], src = slot_src[slot];
    };
    int slot = 0, ch = 0, e = 0, src = 0;
    float cv = 0.0f;
    float2 hfirst = make_float2(0.f, 0.f);
    if (n_my > 0) {
        unit_of(0, slot, ch, e, src);
        cv = (lane < 14) ? __ldg(econst + (size_t)e * 16 + lane) : 0.0f;
        hfirst = __ldg(reinterpret_cast<const float2 *>(a.disps + (size_t)src * HW + (size_t)ch * nsub * TILE + 2 * lane));
    }
    uint32_t sq_cur = 0;
    for (long long i = 0; i < n_my; i++) {
        const int slot_c = slot, ch_c = ch, src_c = src;
        const int px0 = ch_c * nsub * TILE;
        const bool stereo = __shfl_sync(0xffffffffu, cv, 13) != 0.0f;
        if (lane < 14) c[lane] = splat2(cv);
        __syncwarp();
        float2 hnext = hfirst;
        if (i + 1 < n_my) {
            unit_of(i + 1, slot, ch, e, src);
            cv = (lane < 14) ? __ldg(econst + (size_t)e * 16 + lane) : 0.0f;
            hfirst = __ldg(reinterpret_cast<const float2 *>(a.disps + (size_t)src * HW + (size_t)ch * nsub * TILE + 2 * lane));
        }
        PxPos pos;
        pos.set(px0 + 2 * lane, wd);
        float2 hh[kEdgeVals];
#pragma unroll
        for (int q = 0; q < kEdgeVals; q++) hh[q] = make_float2(0.0f, 0.0f);
        const float2 *hrow = reinterpret_cast<const float2 *>(a.disps + (size_t)src_c * HW + px0 + 2 * lane);
        for (int st = 0; st < spi; st++, sq_cur++) {
            const int sb = sq_cur & 1;
            const float *stg = reinterpret_cast<const float *>(mystage + (size_t)sb * stage_bytes) + 2 * lane;
            mbar_wait(myfull + sb, (sq_cur >> 1) & 1);
#pragma unroll
            for (int part = 0; part < STAGE_PX / TILE; part++) {
                const float2 h = hnext;
                const int tn = st * (STAGE_PX / TILE) + part + 1;
                if (tn < nsub) hnext = __ldg(hrow + tn * (TILE / 2));
                float2 xn, yn;
                pos.normalised(kc, xn, yn);
                pos.advance((float)TILE, kc.wdf);
                const float2 tu_s = *reinterpret_cast<const float2 *>(stg + TILE * part);
                const float2 tv_s = *reinterpret_cast<const float2 *>(stg + STAGE_PX + TILE * part);
                const float2 wu_s = *reinterpret_cast<const float2 *>(stg + 2 * STAGE_PX + TILE * part);
                const float2 wv_s = *reinterpret_cast<const float2 *>(stg + 3 * STAGE_PX + TILE * part);
                if (part == STAGE_PX / TILE - 1) {  // stage read (the warp barrier orders the loads before the refill)
                    __syncwarp();
                    issue_stage(sq_cur + 2);
                }
                j_step<false>(kc, c, stereo, xn, yn, h, tu_s, tv_s, wu_s, wv_s, hh, nullptr);
            }
        }
        write_edge_record(hh, lane, a.epart + ((size_t)slot_c * tb.ntile + ch_c) * kEdgeStride);
        __syncwarp();  // the constants are rewritten for the next unit
    }
}

cudaError_t launch_lin3_motion(const LinArgs &a, const int *slot_src_dev, int slot_lo, int nslots, int chunk_px, float *econst_dev,
                               int num_sms, cudaStream_t st) {
    if (nslots <= 0) return cudaSuccess;
    const size_t smem = (size_t)kMotionWarps * (2 * (size_t)4 * STAGE_PX * 4 + 16 * 8 + 16);
    cudaError_t e = cudaFuncSetAttribute(lin3_motion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int nchunk = a.tb.HW / chunk_px, nsub = chunk_px / TILE;
    const long long nunits = (long long)nslots * nchunk;
    long long grid = (nunits + kMotionWarps - 1) / kMotionWarps;
    if (grid > num_sms) grid = num_sms;
    edge_const_kernel<<<(nslots + 127) / 128, 128, 0, st>>>(a.tb.fedge, slot_src_dev, slot_lo, nslots, a.poses, a.tb.e_jj, econst_dev);
    lin3_motion_kernel<<<(int)grid, 32 * kMotionWarps, smem, st>>>(a, slot_src_dev, slot_lo, nslots, nchunk, nsub, econst_dev);
    return cudaGetLastError();
}

// ================================================================================================================
static bool g_pair_lut_ready = false;
static cudaError_t upload_pair_lut() {
    if (g_pair_lut_ready) return cudaSuccess;
    unsigned char pm[256], pp[256];
    int p = 0;
    for (int mp = 0; mp < 22 && p < 256; mp++)
        for (int m = 0; m <= mp && p < 256; m++, p++) pm[p] = (unsigned char)m, pp[p] = (unsigned char)mp;
    for (; p < 256; p++) pm[p] = pp[p] = 0;
    cudaError_t e = cudaMemcpyToSymbol(c_pair_m, pm, 256);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_pair_mp, pp, 256);
    if (e != cudaSuccess) return e;
    g_pair_lut_ready = true;
    return cudaSuccess;
}

bool lin3_supported(int HW, int chunk_px) { return chunk_px >= STAGE_PX && chunk_px % STAGE_PX == 0 && HW % chunk_px == 0; }

cudaError_t launch_lin3(const LinArgs &a, const Lin3Item *items_dev, int nframes, int chunk_px, float *econst_dev, int num_sms,
                        cudaStream_t st) {
    if (nframes <= 0) return cudaSuccess;
    using L = Smem3;
    cudaError_t e = upload_pair_lut();
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lin3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total);
    if (e != cudaSuccess) return e;
    const int nchunk = a.tb.HW / chunk_px, nsub = chunk_px / TILE;
    const int nitems = nframes * nchunk;
    const int grid = nitems < num_sms ? nitems : num_sms;
    lin3_prep_kernel<<<(nframes + 7) / 8, 128, 0, st>>>(items_dev, nframes, nchunk, a.poses, a.tb.e_jj, econst_dev);
#ifdef VBA_LIN3_TRACE
    static long long *trace_dev = nullptr;
    if (!trace_dev) {
        cudaMalloc(&trace_dev, 4 * kTraceLen * sizeof(long long));
        cudaMemcpyToSymbol(g_trace, &trace_dev, sizeof(trace_dev));
    }
    cudaMemsetAsync(trace_dev, 0, 4 * kTraceLen * sizeof(long long), st);
#endif
    lin3_kernel<<<grid, kThreads, L::total, st>>>(a, items_dev, nitems, nsub, econst_dev);
#ifdef VBA_LIN3_TRACE
    {
        cudaStreamSynchronize(st);
        static long long host[4 * kTraceLen];
        cudaMemcpy(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost);
        FILE *f = fopen("gpurun_out/lin3_trace.txt", "w");
        if (f) {
            for (int r = 0; r < 4; r++) {
                const long long n = host[(size_t)r * kTraceLen];
                for (long long i = 0; i < n; i++) {
                    const long long v = host[(size_t)r * kTraceLen + 1 + i];
                    fprintf(f, "%d %lld %lld\n", r, v >> 48, v & 0xffffffffffffLL);
                }
            }
            fclose(f);
        }
    }
#endif
    return cudaGetLastError();
}

}  // namespace vba
