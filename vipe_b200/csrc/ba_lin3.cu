// Blackwell-native linearisation + Schur Gram ("lin3"): the same outputs as linearize2_kernel (epart / gpart / Q / Qw;
// reference: projective_transform_kernel, accum_kernel, EEt6x6_kernel, Ev6x1_kernel, csrc/slam_ext/geom_kernels.cu:178-432,
// 863-880,994-1080), organised as a persistent, warp-specialised pipeline:
//
//   producer warp   cp.async.bulk (TMA, 1-D) copies of the targets/weights rows of every (edge, sub-tile) unit into a
//                   shared-memory ring guarded by full/empty mbarriers; also prepares the per-item edge constants.
//   8 J warps       one (edge, sub-tile) unit at a time: projective transform, residuals, J_j, J_z; per-edge H_jj / v_j /
//                   energy accumulated in registers over the unit's pixel steps, reduced once per unit; the pose-disparity
//                   vectors u (6 per edge-pixel) and the per-edge C, w contributions go to a raw staging buffer.
//   4 post warps    per sub-tile: C, w summed over the edges in fixed order, damping -> Q, Q*w; sqrt(Q)*u split into
//                   (hi, lo) TF32 planes written in the tensor core's K-major 128-byte-swizzled layout; accumulator flush.
//   MMA warp        one thread issues tcgen05.mma kind::tf32: D = [hi; lo] hi^T (M = 128: the hi.hi and lo.hi products in
//                   one instruction), accumulators in tensor memory.
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

constexpr int kJW = 8;        // J warps
constexpr int kPW = 4;        // post warps
constexpr int kThreads = 32 * (2 + kPW + kJW);
constexpr int kNEC = 4;       // item header / edge-constant ring
constexpr int kEc3 = 16;      // float2 entries per edge-constant record
constexpr int kRawRows = 8;   // raw staging rows per edge: u (6), c, w

// pair index -> (m, mp), m <= mp, p = mp (mp + 1) / 2 + m   (d <= 21 -> 231 pairs)
__constant__ unsigned char c_pair_m[256], c_pair_mp[256];

struct ItemHdr {
    int k, src, d, s0, px0, flags, pad0, pad1;
};

template <int TILE, int DMAX, int NST>
struct Smem3 {
    static constexpr int RS = TILE + 4;                        // raw row stride (floats)
    static constexpr int kAtoms = TILE / 32;
    static constexpr size_t split_bytes = (size_t)kAtoms * 128 * 128;
    static constexpr size_t raw_bytes = (size_t)kRawRows * DMAX * RS * 4;
    static constexpr size_t stage_bytes = (size_t)4 * TILE * 4;
    static constexpr size_t off_split = 0;
    static constexpr size_t off_raw = off_split + split_bytes;
    static constexpr size_t off_ring = off_raw + 2 * raw_bytes;
    static constexpr size_t off_ec = off_ring + (size_t)NST * stage_bytes;
    static constexpr size_t off_hdr = off_ec + (size_t)kNEC * DMAX * kEc3 * 8;
    static constexpr size_t off_eacc = off_hdr + (size_t)kNEC * sizeof(ItemHdr);
    static constexpr size_t off_sq = off_eacc + (size_t)kJW * DMAX * kEdgeStride * 4;
    static constexpr size_t off_bar = off_sq + (size_t)2 * TILE * 4;
    static constexpr int n_bar = 2 * NST + 2 + 2 + 1 + 2 + 2 + 2 * kNEC;
    static constexpr size_t off_tmem = off_bar + (size_t)n_bar * 8;
    static constexpr size_t total = off_tmem + 16 + 1024;  // + alignment slack
};

__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ float2 fadd2p(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 splat2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float rcp_apx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// signs of the sign-stripped quantities the J warps accumulate (see j_step): A' = s_a |.|, B' = s_b |.|
// u_hat = s_u * u_true with s_u = (+,+,-,-,+,-); H entries and v entries carry the products below.
__device__ __forceinline__ float hsign(int slot) {
    // slot order: (0,0) (1,1) (2,0) (2,1) (2,2) (3,0) (3,1) (3,2) (3,3) (4,0) (4,1) (4,2) (4,3) (4,4) (5,0) (5,1) (5,2) (5,3) (5,4) (5,5)
    // v0..v5, energy.  Row signs s = (+,+,-,-,+,-): entry (i,j) carries s_i s_j, v_i carries s_i.
    const unsigned neg = (1u << 2) | (1u << 3) | (1u << 5) | (1u << 6) | (1u << 11) | (1u << 12) | (1u << 14) | (1u << 15) | (1u << 18) |
                         (1u << 22) | (1u << 23) | (1u << 25);
    return ((neg >> slot) & 1u) ? -1.0f : 1.0f;
}
__device__ __forceinline__ float usign(int r) { return (r == 2 || r == 3 || r == 5) ? -1.0f : 1.0f; }

}  // namespace

// ================================================================================================================
template <int TILE, int DMAX, int NST, int NSETS>
__global__ void __launch_bounds__(kThreads, 1) lin3_kernel(const LinArgs a, const int *__restrict__ flist, int nframes, int nchunk,
                                                           int nsub) {
    using L = Smem3<TILE, DMAX, NST>;
    constexpr int S = TILE / 64;                 // pixel-pair steps per unit
    constexpr int RS = L::RS;
    constexpr int KSTEPS = TILE / 8;
    constexpr int ACC_COLS = 64;                 // columns of one accumulator set (N <= 64)
    constexpr int TM_BUF = NSETS * ACC_COLS;     // columns of one accumulator buffer
    static_assert(2 * TM_BUF <= 512, "tensor memory");
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *split = smem + L::off_split;
    float *raw = reinterpret_cast<float *>(smem + L::off_raw);
    unsigned char *ring = smem + L::off_ring;
    float2 *ecs = reinterpret_cast<float2 *>(smem + L::off_ec);
    ItemHdr *hdrs = reinterpret_cast<ItemHdr *>(smem + L::off_hdr);
    float *eacc = reinterpret_cast<float *>(smem + L::off_eacc);
    float *sqs = reinterpret_cast<float *>(smem + L::off_sq);
    float *wzs = sqs + TILE;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::off_bar);
    uint64_t *ring_full = bars, *ring_empty = bars + NST;
    uint64_t *raw_full = bars + 2 * NST, *raw_empty = raw_full + 2;
    uint64_t *split_ready = raw_empty + 2;
    uint64_t *mma_done = split_ready + 1, *tmem_free = mma_done + 2;
    uint64_t *hdr_ready = tmem_free + 2, *hdr_free = hdr_ready + kNEC;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::off_tmem);

    const Tables &tb = a.tb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HW = tb.HW;
    const int nitems = nframes * nchunk;

    if (tid == 0) {
        for (int i = 0; i < NST; i++) {
            mbar_init(ring_full + i, 1);
            mbar_init(ring_empty + i, 32);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(raw_full + i, 32 * kJW);
            mbar_init(raw_empty + i, 32 * kPW);
            mbar_init(mma_done + i, 1);
            mbar_init(tmem_free + i, 32 * kPW);
        }
        mbar_init(split_ready, 32 * kPW);
        for (int i = 0; i < kNEC; i++) {
            mbar_init(hdr_ready + i, 32);
            mbar_init(hdr_free + i, 32 * kJW + 32 * kPW + 1);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    // unused operand rows must hold finite numbers (they only feed accumulator entries nobody reads)
    for (int i = tid; i < (int)(L::split_bytes / 16); i += kThreads) reinterpret_cast<float4 *>(split)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const float fx = __ldg(a.intr + 0), fy = __ldg(a.intr + 1), cx = __ldg(a.intr + 2), cy = __ldg(a.intr + 3);

    if (warp == 0) {
        // ============================================================ producer
        uint32_t u = 0;
        int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
            const int slot = it % kNEC;
            mbar_wait(hdr_free + slot, ((it / kNEC) & 1) ^ 1);
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
                float2 *c = ecs + ((size_t)slot * DMAX + lane) * kEc3;
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
                h.pad0 = ch, h.pad1 = 0;
                hdrs[slot] = h;
            }
            mbar_arrive(hdr_ready + slot);
            const int px_item = ch * nsub * TILE;
            for (int t = 0; t < nsub; t++) {
                for (int m = 0; m < d; m++, u++) {
                    const int e = __shfl_sync(0xffffffffu, my_e, m);
                    if (lane == 0) {
                        const int st = u % NST;
                        mbar_wait(ring_empty + st, ((u / NST) & 1) ^ 1);
                        mbar_arrive_expect_tx(ring_full + st, (uint32_t)L::stage_bytes);
                        unsigned char *dst = ring + (size_t)st * L::stage_bytes;
                        const size_t base = (size_t)e * 2 * HW + px_item + t * TILE;
                        bulk_g2s(dst, a.targets + base, TILE * 4, ring_full + st);
                        bulk_g2s(dst + TILE * 4, a.targets + base + HW, TILE * 4, ring_full + st);
                        bulk_g2s(dst + 2 * TILE * 4, a.weights + base, TILE * 4, ring_full + st);
                        bulk_g2s(dst + 3 * TILE * 4, a.weights + base + HW, TILE * 4, ring_full + st);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        if (lane == 0) {
            uint32_t g = 0;
            int it = 0;
            const uint32_t sbase = smem_u32(split);
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
                const int slot = it % kNEC;
                mbar_wait(hdr_ready + slot, (it / kNEC) & 1);
                const int d = hdrs[slot].d;
                const int N = ((6 * d + 1 + 15) / 16) * 16;
                const uint32_t idesc = make_idesc_tf32(128, N);
                for (int t = 0; t < nsub; t++, g++) {
                    const int buf = g & 1;
                    mbar_wait(split_ready, g & 1);
                    mbar_wait(tmem_free + buf, ((g >> 1) & 1) ^ 1);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ks++) {
                        const uint64_t desc = make_desc_k_sw128(sbase + (ks >> 2) * (128 * 128) + (ks & 3) * 32);
                        mma_tf32(tmem + buf * TM_BUF + (ks % NSETS) * ACC_COLS, desc, desc, idesc, ks >= NSETS);
                    }
                    mma_commit(mma_done + buf);
                }
                mbar_arrive(hdr_free + slot);
            }
        }
        __syncwarp();
    } else if (warp < 2 + kPW) {
        // ============================================================ post warps
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
            auto flush = [&](uint32_t gg) {
                const int buf = gg & 1;
                mbar_wait(mma_done + buf, (gg >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < NSETS; s++) {
#pragma unroll
                    for (int c0 = 0; c0 < ACC_COLS; c0 += 16) {
                        if (c0 < N) {
                            float v[16];
                            tmem_ld16(tmem + tm_lane + buf * TM_BUF + s * ACC_COLS + c0, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; i++) acc[c0 + i] += v[i];
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(tmem_free + buf);
            };
            for (int t = 0; t < nsub; t++, g++) {
                const int buf = g & 1;
                const float *rb = raw + (size_t)buf * (L::raw_bytes / 4);
                // the pixel's own data first (their latency hides behind the wait for the J warps)
                float hpx = 0.f, ds = 0.f, et = 0.f;
                const int px = hd.px0 + t * TILE + pt;
                if (pt < TILE) {
                    hpx = __ldg(a.disps + (size_t)hd.src * HW + px);
                    ds = __ldg(a.dsens + (size_t)hd.src * HW + px);
                    et = __ldg(a.eta + (size_t)hd.k * HW + px);
                }
                mbar_wait(raw_full + buf, (g >> 1) & 1);
                if (pt < TILE) {
                    float C = 0.0f, W = 0.0f;
                    for (int m = 0; m < d; m++) {
                        C += rb[(size_t)(kRawRows * m + 6) * RS + pt];
                        W += rb[(size_t)(kRawRows * m + 7) * RS + pt];
                    }
                    float Q, wz;
                    disparity_block(C, W, hpx, ds, et, hd.flags, a.opt, Q, wz);
                    a.qbuf[(size_t)hd.k * HW + px] = Q;
                    a.qwbuf[(size_t)hd.k * HW + px] = Q * wz;
                    const float sq = sqrtf(Q);
                    sqs[pt] = sq;
                    wzs[pt] = sq * wz;
                }
                named_bar(1, 32 * kPW);
                if (t > 0) mbar_wait(mma_done + ((g - 1) & 1), ((g - 1) >> 1) & 1);  // operand buffer free again
                // sqrt(Q) u -> (hi, lo) planes, K-major, 128-byte swizzle; row R-1 is sqrt(Q) w
                constexpr int QUADS = TILE / 4;
                for (int idx = pt; idx < R * QUADS; idx += 32 * kPW) {
                    const int row = idx / QUADS, quad = idx - row * QUADS;
                    float4 v;
                    if (row < R - 1) {
                        const int m = row / 6, r6 = row - 6 * m;
                        const float4 u4 = *reinterpret_cast<const float4 *>(rb + (size_t)(kRawRows * m + r6) * RS + 4 * quad);
                        const float4 s4 = *reinterpret_cast<const float4 *>(sqs + 4 * quad);
                        v = make_float4(u4.x * s4.x, u4.y * s4.y, u4.z * s4.z, u4.w * s4.w);
                    } else {
                        v = *reinterpret_cast<const float4 *>(wzs + 4 * quad);
                    }
                    float4 hi, lo;
                    split_tf32(v.x, hi.x, lo.x);
                    split_tf32(v.y, hi.y, lo.y);
                    split_tf32(v.z, hi.z, lo.z);
                    split_tf32(v.w, hi.w, lo.w);
                    unsigned char *blk = split + (size_t)(quad >> 3) * (128 * 128);
                    *reinterpret_cast<float4 *>(blk + sw128_offset(row, 4 * (quad & 7))) = hi;
                    *reinterpret_cast<float4 *>(blk + sw128_offset(row + 64, 4 * (quad & 7))) = lo;
                }
                fence_async_smem();
                mbar_arrive(split_ready);
                mbar_arrive(raw_empty + buf);
                if (t > 0) flush(g - 1);
            }
            flush(g - 1);
            // ---- item epilogue: Y = hh + lh + lh^T through shared memory (the operand buffer is idle now), signs, -> gpart
            float *Ys = reinterpret_cast<float *>(split);  // [64][65]
            const int r = 32 * (q & 1) + lane;
            if (q < 2) {
#pragma unroll
                for (int c = 0; c < ACC_COLS; c++) Ys[r * 65 + c] = acc[c];
            }
            named_bar(1, 32 * kPW);
            if (q >= 2) {
#pragma unroll
                for (int c = 0; c < ACC_COLS; c++) Ys[r * 65 + c] += acc[c];
            }
            named_bar(1, 32 * kPW);
            if (q >= 2) {
#pragma unroll
                for (int c = 0; c < ACC_COLS; c++) Ys[c * 65 + r] += acc[c];
            }
            named_bar(1, 32 * kPW);
            {
                const int npairs = d * (d + 1) / 2;
                const int rec = npairs * 36 + 6 * d;
                float *gp = a.gpart + tb.gbase[hd.k] + (size_t)hd.pad0 * rec;
                for (int idx = pt; idx < rec; idx += 32 * kPW) {
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
            named_bar(1, 32 * kPW);
            // the operand rows this item used as scratch are rewritten by the next split pass; rows >= its R keep
            // finite scratch values, which is all the padding needs
            mbar_arrive(hdr_free + slot);
        }
    } else {
        // ============================================================ J warps
        const int jw = warp - (2 + kPW);
        const int jt = tid - 32 * (2 + kPW);   // 0..255
        uint32_t ubase = 0, g = 0;
        int it = 0;
        const float2 ifx2 = splat2(__fdiv_rn(1.0f, fx)), ify2 = splat2(__fdiv_rn(1.0f, fy));
        const float2 ncx2 = splat2(-cx * ifx2.x), ncy2 = splat2(-cy * ify2.x);
        const float wsx = kWeightScale * fx * fx, wsy = kWeightScale * fy * fy;
        const float2 m1 = splat2(-1.0f), one2 = splat2(1.0f);
        float *myacc = eacc + (size_t)jw * DMAX * kEdgeStride;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
            const int slot = it % kNEC;
            mbar_wait(hdr_ready + slot, (it / kNEC) & 1);
            const ItemHdr hd = hdrs[slot];
            const int d = hd.d;
            for (int i = lane; i < d * kEdgeStride; i += 32) myacc[i] = 0.0f;
            __syncwarp();
            for (int t = 0; t < nsub; t++, g++) {
                const int buf = g & 1;
                float *rb = raw + (size_t)buf * (L::raw_bytes / 4);
                mbar_wait(raw_empty + buf, ((g >> 1) & 1) ^ 1);
                bool have_px = false;
                float2 xn[S], yn[S], h[S];
                for (int m = 0; m < d; m++) {
                    const uint32_t u = ubase + t * d + m;
                    if ((int)(u % kJW) != jw) continue;
                    if (!have_px) {
                        have_px = true;
#pragma unroll
                        for (int s = 0; s < S; s++) {
                            const int px = hd.px0 + t * TILE + 64 * s + 2 * lane;
                            h[s] = __ldg(reinterpret_cast<const float2 *>(a.disps + (size_t)hd.src * HW + px));
                            const int row0 = px / tb.wd, col0 = px - row0 * tb.wd;
                            const int row1 = (px + 1) / tb.wd, col1 = (px + 1) - row1 * tb.wd;
                            xn[s] = make_float2(__fdiv_rn((float)col0 - cx, fx), __fdiv_rn((float)col1 - cx, fx));  // geom_kernels.cu:289-290
                            yn[s] = make_float2(__fdiv_rn((float)row0 - cy, fy), __fdiv_rn((float)row1 - cy, fy));
                        }
                    }
                    const int st = u % NST;
                    const float *stg = reinterpret_cast<const float *>(ring + (size_t)st * L::stage_bytes);
                    mbar_wait(ring_full + st, (u / NST) & 1);
                    float2 tu[S], tv[S], wu[S], wv[S];
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        const int o = 64 * s + 2 * lane;
                        tu[s] = *reinterpret_cast<const float2 *>(stg + o);
                        tv[s] = *reinterpret_cast<const float2 *>(stg + TILE + o);
                        wu[s] = *reinterpret_cast<const float2 *>(stg + 2 * TILE + o);
                        wv[s] = *reinterpret_cast<const float2 *>(stg + 3 * TILE + o);
                    }
                    mbar_arrive(ring_empty + st);
                    const float2 *c = ecs + ((size_t)slot * DMAX + m) * kEc3;
                    const bool stereo = c[13].x != 0.0f;
                    float2 hh[kEdgeVals];
#pragma unroll
                    for (int i = 0; i < kEdgeVals; i++) hh[i] = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int s = 0; s < S; s++) {
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
                        const float2 w_u = fmul2(wu[s], make_float2(v0 ? wsx : 0.0f, v1 ? wsx : 0.0f));
                        const float2 w_v = fmul2(wv[s], make_float2(v0 ? wsy : 0.0f, v1 ? wsy : 0.0f));
                        const float2 X = fmul2(x, dd), Y = fmul2(y, dd), aa = fmul2(h[s], dd);
                        const float2 ru = ffma2(X, m1, ffma2(tu[s], ifx2, ncx2));  // (:308-309) / fx
                        const float2 rv = ffma2(Y, m1, ffma2(tv[s], ify2, ncy2));
                        // dl/dd (:322,363) / fx
                        const float2 Jzu = fmul2(dd, ffma2(nt2, X, t0));
                        const float2 Jzv = fmul2(dd, ffma2(nt2, Y, t1));
                        const float2 wJu = fmul2(w_u, Jzu), wJv = fmul2(w_v, Jzv);
                        const int o = 64 * s + 2 * lane;
                        *reinterpret_cast<float2 *>(rb + (size_t)(kRawRows * m + 6) * RS + o) = ffma2(wJu, Jzu, fmul2(wJv, Jzv));  // :325,364
                        *reinterpret_cast<float2 *>(rb + (size_t)(kRawRows * m + 7) * RS + o) = ffma2(wJu, ru, fmul2(wJv, rv));    // :326,365
                        if (stereo) {  // stereo edges contribute nothing beyond C and w (:329,367); uniform over the warp
#pragma unroll
                            for (int i = 0; i < 6; i++) *reinterpret_cast<float2 *>(rb + (size_t)(kRawRows * m + i) * RS + o) = make_float2(0.f, 0.f);
                            continue;
                        }
                        // sign-stripped J_j rows (:314-319,356-361) in normalised coordinates:
                        //   A' = (a, 0, -P2, -P3, P4, -P5),  B' = (0, a, -Q2, -Q3, Q4, -Q5),  Q5 = -X
                        const float2 P2 = fmul2(X, aa), P3 = fmul2(X, Y), P4 = ffma2(X, X, one2), P5 = Y;
                        const float2 Q2 = fmul2(Y, aa), Q3 = ffma2(Y, Y, one2), Q4 = P3, Q5 = fmul2(X, m1);
                        // u_hat = (u0, u1, -u2, -u3, u4, -u5), u = wu Jzu A + wv Jzv B (E_ij column, :351,385)
                        float2 uh[6];
                        uh[0] = fmul2(wJu, aa);
                        uh[1] = fmul2(wJv, aa);
                        uh[2] = ffma2(wJu, P2, fmul2(wJv, Q2));
                        uh[3] = ffma2(wJu, P3, fmul2(wJv, Q3));
                        uh[4] = ffma2(wJu, P4, fmul2(wJv, Q4));
                        uh[5] = ffma2(wJu, P5, fmul2(wJv, Q5));
#pragma unroll
                        for (int i = 0; i < 6; i++) *reinterpret_cast<float2 *>(rb + (size_t)(kRawRows * m + i) * RS + o) = uh[i];
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
                    if (!stereo) {
                        float acc[32];
#pragma unroll
                        for (int i = 0; i < kEdgeVals; i++) acc[i] = hh[i].x + hh[i].y;
#pragma unroll
                        for (int i = kEdgeVals; i < 32; i++) acc[i] = 0.0f;
                        const float tot = warp_transpose_reduce<32>(acc, lane);
                        if (lane < kEdgeVals) myacc[m * kEdgeStride + lane] += tot;
                    }
                }
                mbar_arrive(raw_full + buf);
            }
            ubase += (uint32_t)nsub * d;
            // per-(edge, chunk) record: fixed-order sum over the J warps, true signs, image units
            named_bar(2, 32 * kJW);
            for (int idx = jt; idx < d * kEdgeVals; idx += 32 * kJW) {
                const int m = idx / kEdgeVals, r = idx - m * kEdgeVals;
                float s = 0.0f;
#pragma unroll
                for (int w = 0; w < kJW; w++) s += eacc[((size_t)w * DMAX + m) * kEdgeStride + r];
                // residuals were divided by fx (u) / fy (v): v_j and the energy are sums of w' A' r', already in image units
                a.epart[((size_t)(hd.s0 + m) * tb.ntile + hd.pad0) * kEdgeStride + r] = s * hsign(r);
            }
            named_bar(2, 32 * kJW);
            mbar_arrive(hdr_free + slot);
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
    constexpr int TILE = 128, DMAX = kLin3MaxDeg, NST = 12, NSETS = 4;
    using L = Smem3<TILE, DMAX, NST>;
    cudaError_t e = upload_pair_lut();
    if (e != cudaSuccess) return e;
    auto kern = lin3_kernel<TILE, DMAX, NST, NSETS>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total);
    if (e != cudaSuccess) return e;
    const int nchunk = a.tb.HW / chunk_px, nsub = chunk_px / TILE;
    const int nitems = nframes * nchunk;
    const int grid = nitems < num_sms ? nitems : num_sms;
    kern<<<grid, kThreads, L::total, st>>>(a, flist_dev, nframes, nchunk, nsub);
    return cudaGetLastError();
}

}  // namespace vba
