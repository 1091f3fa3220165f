// Blackwell-native linearisation + Schur Gram ("lin3"): the same outputs as linearize2_kernel (epart / gpart / Q / Qw;
// reference: projective_transform_kernel, accum_kernel, EEt6x6_kernel, Ev6x1_kernel, csrc/slam_ext/geom_kernels.cu:178-432,
// 863-880,994-1080), organised as a persistent, warp-specialised pipeline:
//
//   producer warp   cp.async.bulk (TMA, 1-D) copies of the targets/weights rows of every (edge, sub-tile) unit into a
//                   shared-memory ring guarded by full/empty mbarriers; also prepares the per-item edge constants.
//   10 J warps      warp w owns edge w of the frame: projective transform, residuals, J_j, J_z of its edge over the
//                   sub-tile; H_jj / v_j / energy stay in registers for the whole item (one cross-lane reduction per item);
//                   u (6 per edge-pixel) and the edge's C, w contributions go to a raw staging buffer.  After a group
//                   barrier the first 128 J threads sum C, w over the edges in edge order and apply the damping
//                   (Q, Q*w); after a second one every warp rewrites ITS rows as sqrt(Q)*u split into (hi, lo) TF32
//                   planes in the tensor core's K-major 128-byte-swizzled operand layout (double-buffered).
//   MMA warp        one thread issues tcgen05.mma kind::tf32: D = [hi; lo] hi^T (M = 128: the hi.hi and lo.hi products in
//                   one instruction), accumulators in tensor memory (double-buffered).
//   4 flush warps   tcgen05.ld of the accumulators after every sub-tile, round-to-nearest fp32 sums in registers; per
//                   item Y = hh + lh + lh^T and the sign fix-up, written to gpart.
//
// Y = sum_px Q u u^T = hh + lh + lh^T (the lo.lo term, 2^-22 relative, is dropped).  The tensor core truncates its fp32
// accumulator (measured: about -1 ulp of the running sum per K = 8 step, scripts/tc_gram_probe.cu), and A - S cancels
// in the gauge directions, so the in-unit accumulation chains are kept to TILE / 8 / NSETS steps: the K steps of a
// sub-tile rotate over NSETS independent accumulators (which also hides the ~120-cycle dependent-MMA latency) and the
// post warps add them up in round-to-nearest fp32 after every sub-tile.
#include "ba_common.cuh"
#include "ba_launch.h"
#include "sm100_async.cuh"

namespace vba {
using namespace sm100;

namespace {

constexpr int kNJ = kLin3MaxDeg;  // J warps: warp w owns edge w of the frame (<= kLin3MaxDeg edges)
constexpr int kNF = 4;            // flush warps: one per tensor-memory lane quarter
constexpr int kThreads = 32 * (2 + kNF + kNJ);
constexpr int kNEC = 4;           // item header / edge-constant ring
constexpr int kEc3 = 16;          // float2 entries per edge-constant record
constexpr int kRawRows = 8;       // raw staging rows per edge: u (6), c, w

// optional pipeline trace (VIPE_BA_LIN3_TRACE=1): clock stamps of CTA 0, [role][event], dumped by the launcher
#ifdef VBA_LIN3_TRACE
__device__ long long *g_trace = nullptr;
constexpr int kTraceLen = 4096;
#define TR(role, cond, tag)                                                                   \
    do {                                                                                      \
        if (blockIdx.x == 0 && (cond) && g_trace && tr_n + 2 < kTraceLen) {                   \
            g_trace[(size_t)(role) * kTraceLen + 1 + tr_n] = ((long long)(tag) << 48) | (clock64() & 0xffffffffffffLL); \
            tr_n++;                                                                           \
            g_trace[(size_t)(role) * kTraceLen] = tr_n;                                       \
        }                                                                                     \
    } while (0)
#else
#define TR(role, cond, tag) do { } while (0)
#endif

// pair index -> (m, mp), m <= mp, p = mp (mp + 1) / 2 + m   (d <= 21 -> 231 pairs)
__constant__ unsigned char c_pair_m[256], c_pair_mp[256];

struct ItemHdr {
    int k, src, d, s0, px0, flags, chunk, pad1;
};

template <int TILE, int NST>
struct Smem3 {
    static constexpr int RS = TILE;                            // raw row stride (floats)
    static constexpr int kAtoms = TILE / 32;
    static constexpr size_t op_bytes = (size_t)kAtoms * 128 * 128;  // one operand buffer: [hi(64) ; lo(64)] rows x TILE
    static constexpr size_t raw_bytes = (size_t)kRawRows * kNJ * RS * 4;
    static constexpr size_t stage_bytes = (size_t)4 * TILE * 4;
    static constexpr size_t off_op = 0;
    static constexpr size_t off_raw = off_op + 2 * op_bytes;
    static constexpr size_t off_ring = off_raw + raw_bytes;
    static constexpr size_t off_ys = off_ring + (size_t)NST * stage_bytes;
    static constexpr size_t off_ec = off_ys + (size_t)64 * 65 * 4;
    static constexpr size_t off_hdr = off_ec + (size_t)kNEC * kNJ * kEc3 * 8;
    static constexpr size_t off_sq = off_hdr + (size_t)kNEC * sizeof(ItemHdr);
    static constexpr size_t off_bar = off_sq + (size_t)2 * TILE * 4;
    static constexpr int n_bar = 2 * NST + 2 + 2 + 2 + 2 * kNEC;
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

// The J warps accumulate sign-stripped quantities (see the J step): with row signs s = (+,+,-,-,+,-), entry (i,j) of H_jj
// carries s_i s_j, v_i carries s_i, and the staged u_hat = s * u.
__device__ __forceinline__ float hsign(int slot) {
    // slot order: (0,0) (1,1) (2,0) (2,1) (2,2) (3,0) (3,1) (3,2) (3,3) (4,0) (4,1) (4,2) (4,3) (4,4) (5,0) (5,1) (5,2) (5,3) (5,4) (5,5), v0..v5, energy
    const unsigned neg = (1u << 2) | (1u << 3) | (1u << 5) | (1u << 6) | (1u << 11) | (1u << 12) | (1u << 14) | (1u << 15) | (1u << 18) |
                         (1u << 22) | (1u << 23) | (1u << 25);
    return ((neg >> slot) & 1u) ? -1.0f : 1.0f;
}
__device__ __forceinline__ float usign(int r) { return (r == 2 || r == 3 || r == 5) ? -1.0f : 1.0f; }

}  // namespace

// ================================================================================================================
template <int TILE, int NST, int NSETS>
__global__ void __launch_bounds__(kThreads, 1) lin3_kernel(const LinArgs a, const int *__restrict__ flist, int nframes, int nchunk,
                                                           int nsub) {
    using L = Smem3<TILE, NST>;
    constexpr int S = TILE / 64;                 // pixel-pair steps per unit
    constexpr int RS = L::RS;
    constexpr int KSTEPS = TILE / 8;
    constexpr int ACC_COLS = 64;                 // columns of one accumulator set (N <= 64)
    constexpr int TM_BUF = NSETS * ACC_COLS;     // columns of one accumulator buffer
    static_assert(2 * TM_BUF <= 512, "tensor memory");
    static_assert(TILE == 128, "one pixel quad per lane in the split pass");
    extern __shared__ __align__(1024) unsigned char smem[];  // operand buffers first: the swizzle needs 1024-byte alignment
    unsigned char *op = smem + L::off_op;
    float *raw = reinterpret_cast<float *>(smem + L::off_raw);
    unsigned char *ring = smem + L::off_ring;
    float *Ys = reinterpret_cast<float *>(smem + L::off_ys);  // [64][65]
    float2 *ecs = reinterpret_cast<float2 *>(smem + L::off_ec);
    ItemHdr *hdrs = reinterpret_cast<ItemHdr *>(smem + L::off_hdr);
    float *sqs = reinterpret_cast<float *>(smem + L::off_sq);
    float *wzs = sqs + TILE;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::off_bar);
    uint64_t *ring_full = bars, *ring_empty = bars + NST;
    uint64_t *split_ready = bars + 2 * NST;
    uint64_t *mma_done = split_ready + 2, *tmem_free = mma_done + 2;
    uint64_t *hdr_ready = tmem_free + 2, *hdr_free = hdr_ready + kNEC;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::off_tmem);

    const Tables &tb = a.tb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HW = tb.HW;
    const int nitems = nframes * nchunk;
#ifdef VBA_LIN3_TRACE
    int tr_n = 0;
#endif

    if (tid == 0) {
        for (int i = 0; i < NST; i++) {
            mbar_init(ring_full + i, 1);
            mbar_init(ring_empty + i, 32);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(split_ready + i, 32 * kNJ);
            mbar_init(mma_done + i, 1);
            mbar_init(tmem_free + i, 32 * kNF);
        }
        for (int i = 0; i < kNEC; i++) {
            mbar_init(hdr_ready + i, 32);
            mbar_init(hdr_free + i, 32 * kNJ + 32 * kNF + 1);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    // unused operand rows must hold finite numbers (they only feed accumulator entries nobody reads)
    for (int i = tid; i < (int)(2 * L::op_bytes / 16); i += kThreads) reinterpret_cast<float4 *>(op)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ============================================================ producer
        uint32_t u = 0;
        int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
            const int slot = it % kNEC;
            mbar_wait_relaxed(hdr_free + slot, ((it / kNEC) & 1) ^ 1);
            const int fi = item / nchunk, ch = item - fi * nchunk;
            const int k = flist[fi];
            const int src = tb.kx[k];
            const int s0 = tb.fptr[k];
            const int d = tb.fptr[k + 1] - s0;
            int my_e = 0;
            if (lane < d) {
                my_e = tb.fedge[s0 + lane];
                RelPose<float> rp;
                relative_pose<float>(a.poses, src, tb.e_jj[my_e], rp);
                float2 *c = ecs + ((size_t)slot * kNJ + lane) * kEc3;
#pragma unroll
                for (int q = 0; q < 9; q++) c[q] = splat2(rp.R[q]);
#pragma unroll
                for (int q = 0; q < 3; q++) c[9 + q] = splat2(rp.t[q]);
                c[12] = splat2(-rp.t[2]);
                c[13] = make_float2(rp.stereo ? 1.0f : 0.0f, __int_as_float(my_e));
            }
            if (lane == 0) {
                ItemHdr h;
                h.k = k, h.src = src, h.d = d, h.s0 = s0, h.px0 = ch * nsub * TILE;
                h.flags = a.opt.frame_flags ? a.opt.frame_flags[k] : 0;
                h.chunk = ch, h.pad1 = 0;
                hdrs[slot] = h;
            }
            mbar_arrive(hdr_ready + slot);
            TR(0, lane == 0, 2);
            const int px_item = ch * nsub * TILE;
            for (int t = 0; t < nsub; t++) {
                for (int m = 0; m < d; m++, u++) {
                    const int e = __shfl_sync(0xffffffffu, my_e, m);
                    if (lane == 0) {
                        const int st = u % NST;
                        TR(0, true, 3);
                        mbar_wait_relaxed(ring_empty + st, ((u / NST) & 1) ^ 1);
                        TR(0, true, 4);
                        mbar_arrive_expect_tx(ring_full + st, (uint32_t)L::stage_bytes);
                        TR(0, true, 5);
                        unsigned char *dst = ring + (size_t)st * L::stage_bytes;
                        const size_t base = (size_t)e * 2 * HW + px_item + t * TILE;
                        bulk_g2s(dst, a.targets + base, TILE * 4, ring_full + st);
                        TR(0, true, 6);
                        bulk_g2s(dst + TILE * 4, a.targets + base + HW, TILE * 4, ring_full + st);
                        bulk_g2s(dst + 2 * TILE * 4, a.weights + base, TILE * 4, ring_full + st);
                        bulk_g2s(dst + 3 * TILE * 4, a.weights + base + HW, TILE * 4, ring_full + st);
                        TR(0, true, 1);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        if (lane == 0) {
            uint32_t g = 0;
            int it = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
                const int slot = it % kNEC;
                mbar_wait(hdr_ready + slot, (it / kNEC) & 1);
                const int d = hdrs[slot].d;
                const int N = ((6 * d + 1 + 15) / 16) * 16;
                const uint32_t idesc = make_idesc_tf32(128, N);
                for (int t = 0; t < nsub; t++, g++) {
                    const int buf = g & 1;
                    const uint32_t sbase = smem_u32(op) + buf * (uint32_t)L::op_bytes;
                    TR(1, true, 1);
                    mbar_wait_relaxed(split_ready + buf, (g >> 1) & 1);
                    TR(1, true, 2);
                    mbar_wait(tmem_free + buf, ((g >> 1) & 1) ^ 1);
                    TR(1, true, 3);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ks++) {
                        const uint64_t desc = make_desc_k_sw128(sbase + (ks >> 2) * (128 * 128) + (ks & 3) * 32);
                        mma_tf32(tmem + buf * TM_BUF + (ks % NSETS) * ACC_COLS, desc, desc, idesc, ks >= NSETS);
                    }
                    mma_commit(mma_done + buf);
                    TR(1, true, 4);
                }
                mbar_arrive(hdr_free + slot);
            }
        }
        __syncwarp();
    } else if (warp < 2 + kNF) {
        // ============================================================ flush warps
        const int pt = tid - 64;               // 0..127
        const int q = warp & 3;                // TMEM lane quarter this warp may read
        const uint32_t tm_lane = (uint32_t)(32 * q) << 16;
        uint32_t g = 0;
        int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
            const int slot = it % kNEC;
            mbar_wait(hdr_ready + slot, (it / kNEC) & 1);
            const ItemHdr hd = hdrs[slot];
            const int d = hd.d;
            const int R = 6 * d + 1;
            const int N = ((R + 15) / 16) * 16;
            float acc[ACC_COLS];
#pragma unroll
            for (int c = 0; c < ACC_COLS; c++) acc[c] = 0.0f;
            for (int t = 0; t < nsub; t++, g++) {
                const int buf = g & 1;
                TR(2, pt == 0, 1);
                mbar_wait_relaxed(mma_done + buf, (g >> 1) & 1);
                TR(2, pt == 0, 2);
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
                mbar_arrive(tmem_free + buf);
                TR(2, pt == 0, 3);
            }
            // ---- item epilogue: Y = hh + lh + lh^T through shared memory, signs, -> gpart
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
            {
                const int npairs = d * (d + 1) / 2;
                const int rec = npairs * 36 + 6 * d;
                float *gp = a.gpart + tb.gbase[hd.k] + (size_t)hd.chunk * rec;
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
            }
            named_bar(1, 32 * kNF);
            mbar_arrive(hdr_free + slot);
            TR(2, pt == 0, 4);
        }
    } else {
        // ============================================================ J warps
        const int jw = warp - (2 + kNF);
        const int jt = tid - 32 * (2 + kNF);   // 0..319
        const float fx = __ldg(a.intr + 0), fy = __ldg(a.intr + 1), cx = __ldg(a.intr + 2), cy = __ldg(a.intr + 3);
        const float ifx = __fdiv_rn(1.0f, fx), ify = __fdiv_rn(1.0f, fy);
        const float2 ifx2 = splat2(ifx), ify2 = splat2(ify);
        const float2 ncx2 = splat2(-cx * ifx), ncy2 = splat2(-cy * ify);
        const float wsx = kWeightScale * fx * fx, wsy = kWeightScale * fy * fy;
        const float2 m1 = splat2(-1.0f), one2 = splat2(1.0f);
        const int wd = tb.wd;
        uint32_t ubase = 0, g = 0;
        int it = 0;
        float *myraw = raw + (size_t)kRawRows * jw * RS;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
            const int slot = it % kNEC;
            mbar_wait(hdr_ready + slot, (it / kNEC) & 1);
            const ItemHdr hd = hdrs[slot];
            const int d = hd.d;
            const bool active = jw < d;
            const float2 *c = ecs + ((size_t)slot * kNJ + jw) * kEc3;
            const bool stereo = active && c[13].x != 0.0f;
            // image coordinates of this lane's first pixel in the current sub-tile (advanced by TILE pixels per sub-tile)
            int rowb = (hd.px0 + 2 * lane) / wd, colb = (hd.px0 + 2 * lane) - rowb * wd;
            float2 hh[kEdgeVals];
#pragma unroll
            for (int i = 0; i < kEdgeVals; i++) hh[i] = make_float2(0.0f, 0.0f);
            float2 hnext[S];
#pragma unroll
            for (int s = 0; s < S; s++)
                hnext[s] = active ? __ldg(reinterpret_cast<const float2 *>(a.disps + (size_t)hd.src * HW + hd.px0 + 64 * s + 2 * lane))
                                  : make_float2(0.f, 0.f);
            for (int t = 0; t < nsub; t++, g++) {
                const int buf = g & 1;
                // the Q pass of this sub-tile is done by the first TILE threads of the J group: their pixel's data leaves now
                float hpx = 0.f, ds = 0.f, et = 0.f;
                const int qpx = hd.px0 + t * TILE + jt;
                if (jt < TILE) {
                    hpx = __ldg(a.disps + (size_t)hd.src * HW + qpx);
                    ds = __ldg(a.dsens + (size_t)hd.src * HW + qpx);
                    et = __ldg(a.eta + (size_t)hd.k * HW + qpx);
                }
                if (active) {
                    float2 h[S], xn[S], yn[S];
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        h[s] = hnext[s];
                        if (t + 1 < nsub)
                            hnext[s] = __ldg(reinterpret_cast<const float2 *>(a.disps + (size_t)hd.src * HW + hd.px0 + (t + 1) * TILE + 64 * s + 2 * lane));
                        // (col - cx) / fx, (row - cy) / fy (geom_kernels.cu:289-290) as a multiplication by the reciprocal
                        int r0 = rowb, c0 = colb + 64 * s;
                        while (c0 >= wd) c0 -= wd, r0++;
                        int r1 = r0, c1 = c0 + 1;
                        if (c1 >= wd) c1 -= wd, r1++;
                        xn[s] = make_float2(((float)c0 - cx) * ifx, ((float)c1 - cx) * ifx);
                        yn[s] = make_float2(((float)r0 - cy) * ify, ((float)r1 - cy) * ify);
                    }
                    colb += TILE;
                    while (colb >= wd) colb -= wd, rowb++;
                    const uint32_t u = ubase + t * d + jw;
                    const int st = u % NST;
                    const float *stg = reinterpret_cast<const float *>(ring + (size_t)st * L::stage_bytes);
                    TR(3, jt == 0, 1);
                    mbar_wait(ring_full + st, (u / NST) & 1);
                    TR(3, jt == 0, 2);
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        const int o = 64 * s + 2 * lane;
                        const float2 tu_s = *reinterpret_cast<const float2 *>(stg + o);
                        const float2 tv_s = *reinterpret_cast<const float2 *>(stg + TILE + o);
                        const float2 wu_s = *reinterpret_cast<const float2 *>(stg + 2 * TILE + o);
                        const float2 wv_s = *reinterpret_cast<const float2 *>(stg + 3 * TILE + o);
                        if (s == S - 1) mbar_arrive(ring_empty + st);  // the stage's last read is in flight: the arrive orders after it
                        // X_j = R X_i + h t (actSE3, :295), X_i = (xn, yn, 1, h)
                        const float2 t0 = c[9], t1 = c[10], t2 = c[11], nt2 = c[12];
                        const float2 x = ffma2(c[0], xn[s], ffma2(c[1], yn[s], ffma2(h[s], t0, c[2])));
                        const float2 y = ffma2(c[3], xn[s], ffma2(c[4], yn[s], ffma2(h[s], t1, c[5])));
                        const float2 z = ffma2(c[6], xn[s], ffma2(c[7], yn[s], ffma2(h[s], t2, c[8])));
                        const bool v0 = depth_valid(z.x, a.opt), v1 = depth_valid(z.y, a.opt);  // :301
                        float2 r = make_float2(rcp_apx(z.x), rcp_apx(z.y));
                        r = ffma2(r, ffma2(z, fmul2(r, m1), one2), r);  // one Newton step: <= 1 ulp
                        const float2 dd = make_float2(v0 ? r.x : 0.0f, v1 ? r.y : 0.0f);
                        // weights in normalised image coordinates: w' = 0.001 w fx^2 (:304-305), residual r' = r / fx
                        const float2 w_u = fmul2(wu_s, make_float2(v0 ? wsx : 0.0f, v1 ? wsx : 0.0f));
                        const float2 w_v = fmul2(wv_s, make_float2(v0 ? wsy : 0.0f, v1 ? wsy : 0.0f));
                        const float2 X = fmul2(x, dd), Y = fmul2(y, dd), aa = fmul2(h[s], dd);
                        const float2 ru = ffma2(X, m1, ffma2(tu_s, ifx2, ncx2));  // (:308-309) / fx
                        const float2 rv = ffma2(Y, m1, ffma2(tv_s, ify2, ncy2));
                        // dl/dd (:322,363) / fx
                        const float2 Jzu = fmul2(dd, ffma2(nt2, X, t0));
                        const float2 Jzv = fmul2(dd, ffma2(nt2, Y, t1));
                        const float2 wJu = fmul2(w_u, Jzu), wJv = fmul2(w_v, Jzv);
                        *reinterpret_cast<float2 *>(myraw + 6 * RS + o) = ffma2(wJu, Jzu, fmul2(wJv, Jzv));  // :325,364
                        *reinterpret_cast<float2 *>(myraw + 7 * RS + o) = ffma2(wJu, ru, fmul2(wJv, rv));    // :326,365
                        if (stereo) {  // stereo edges contribute nothing beyond C and w (:329,367); uniform over the warp
#pragma unroll
                            for (int i = 0; i < 6; i++) *reinterpret_cast<float2 *>(myraw + i * RS + o) = make_float2(0.f, 0.f);
                            continue;
                        }
                        // sign-stripped J_j rows (:314-319,356-361) in normalised coordinates:
                        //   A' = (a, 0, -P2, -P3, P4, -P5),  B' = (0, a, -Q2, -Q3, Q4, -Q5),  Q5 = -X
                        const float2 P2 = fmul2(X, aa), P3 = fmul2(X, Y), P4 = ffma2(X, X, one2), P5 = Y;
                        const float2 Q2 = fmul2(Y, aa), Q3 = ffma2(Y, Y, one2), Q4 = P3, Q5 = fmul2(X, m1);
                        // u_hat = (u0, u1, -u2, -u3, u4, -u5), u = wu Jzu A + wv Jzv B (E_ij column, :351,385)
                        *reinterpret_cast<float2 *>(myraw + 0 * RS + o) = fmul2(wJu, aa);
                        *reinterpret_cast<float2 *>(myraw + 1 * RS + o) = fmul2(wJv, aa);
                        *reinterpret_cast<float2 *>(myraw + 2 * RS + o) = ffma2(wJu, P2, fmul2(wJv, Q2));
                        *reinterpret_cast<float2 *>(myraw + 3 * RS + o) = ffma2(wJu, P3, fmul2(wJv, Q3));
                        *reinterpret_cast<float2 *>(myraw + 4 * RS + o) = ffma2(wJu, P4, fmul2(wJv, Q4));
                        *reinterpret_cast<float2 *>(myraw + 5 * RS + o) = ffma2(wJu, P5, fmul2(wJv, Q5));
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
                }
                TR(3, jt == 0, 3);
                named_bar(2, 32 * kNJ);
                TR(3, jt == 0, 4);
                if (jt < TILE) {
                    // disparity block of pixel jt: C, w summed over the edges in edge order, damping / prior (:1359-1370), Q = 1/C
                    float C = 0.0f, W = 0.0f;
                    for (int m = 0; m < d; m++) {
                        C += raw[(size_t)(kRawRows * m + 6) * RS + jt];
                        W += raw[(size_t)(kRawRows * m + 7) * RS + jt];
                    }
                    float Q, wz;
                    disparity_block(C, W, hpx, ds, et, hd.flags, a.opt, Q, wz);
                    a.qbuf[(size_t)hd.k * HW + qpx] = Q;
                    a.qwbuf[(size_t)hd.k * HW + qpx] = Q * wz;
                    const float sq = sqrtf(Q);
                    sqs[jt] = sq;
                    wzs[jt] = sq * wz;
                }
                TR(3, jt == 0, 5);
                named_bar(2, 32 * kNJ);
                TR(3, jt == 0, 6);
                // sqrt(Q) u -> (hi, lo) TF32 planes of operand buffer `buf`: K-major, 128-byte swizzle.  Lane = pixel quad.
                if (g >= 2) mbar_wait(mma_done + buf, ((g >> 1) - 1) & 1);  // the MMAs that read this buffer two sub-tiles ago are done
                TR(3, jt == 0, 7);
                unsigned char *blk = op + (size_t)buf * L::op_bytes + (size_t)(lane >> 3) * (128 * 128);
                auto put_row = [&](int row, float4 v) {
                    float4 hi, lo;
                    split_tf32(v.x, hi.x, lo.x);
                    split_tf32(v.y, hi.y, lo.y);
                    split_tf32(v.z, hi.z, lo.z);
                    split_tf32(v.w, hi.w, lo.w);
                    const uint32_t off = sw128_offset(row, 4 * (lane & 7));
                    *reinterpret_cast<float4 *>(blk + off) = hi;
                    *reinterpret_cast<float4 *>(blk + off + 8 * 1024) = lo;  // row + 64
                };
                if (active) {
                    const float4 s4 = *reinterpret_cast<const float4 *>(sqs + 4 * lane);
#pragma unroll
                    for (int r6 = 0; r6 < 6; r6++) {
                        const float4 u4 = *reinterpret_cast<const float4 *>(myraw + r6 * RS + 4 * lane);
                        put_row(6 * jw + r6, make_float4(u4.x * s4.x, u4.y * s4.y, u4.z * s4.z, u4.w * s4.w));
                    }
                }
                if (jw == (d % kNJ)) put_row(6 * d, *reinterpret_cast<const float4 *>(wzs + 4 * lane));  // row 6d: sqrt(Q) w
                fence_async_smem();
                mbar_arrive(split_ready + buf);
                TR(3, jt == 0, 8);
            }
            ubase += (uint32_t)nsub * d;
            // per-(edge, chunk) record: this warp's edge, summed over its lanes; true signs, image units
            if (active) {
                float acc[32];
#pragma unroll
                for (int i = 0; i < kEdgeVals; i++) acc[i] = hh[i].x + hh[i].y;
#pragma unroll
                for (int i = kEdgeVals; i < 32; i++) acc[i] = 0.0f;
                const float tot = warp_transpose_reduce<32>(acc, lane);
                if (lane < kEdgeVals) a.epart[((size_t)(hd.s0 + jw) * tb.ntile + hd.chunk) * kEdgeStride + lane] = tot * hsign(lane);
            }
            mbar_arrive(hdr_free + slot);
            TR(3, jt == 0, 9);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
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

bool lin3_supported(int HW, int chunk_px) { return chunk_px >= 128 && chunk_px % 128 == 0 && HW % chunk_px == 0; }

cudaError_t launch_lin3(const LinArgs &a, const int *flist_dev, int nframes, int chunk_px, int num_sms, cudaStream_t st) {
    if (nframes <= 0) return cudaSuccess;
    constexpr int TILE = 128, NST = 12, NSETS = 4;
    using L = Smem3<TILE, NST>;
    cudaError_t e = upload_pair_lut();
    if (e != cudaSuccess) return e;
    auto kern = lin3_kernel<TILE, NST, NSETS>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total);
    if (e != cudaSuccess) return e;
    const int nchunk = a.tb.HW / chunk_px, nsub = chunk_px / TILE;
    const int nitems = nframes * nchunk;
    const int grid = nitems < num_sms ? nitems : num_sms;
#ifdef VBA_LIN3_TRACE
    static long long *trace_dev = nullptr;
    if (!trace_dev) {
        cudaMalloc(&trace_dev, 4 * kTraceLen * sizeof(long long));
        cudaMemcpyToSymbol(g_trace, &trace_dev, sizeof(trace_dev));
    }
    cudaMemsetAsync(trace_dev, 0, 4 * kTraceLen * sizeof(long long), st);
#endif
    kern<<<grid, kThreads, L::total, st>>>(a, flist_dev, nframes, nchunk, nsub);
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
