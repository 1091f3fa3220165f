// Blackwell-native linearisation + Schur Gram on the tensor cores ("lin4"): the same outputs as linearize2_kernel
// (epart / gpart / Q / Qw; reference: projective_transform_kernel, accum_kernel, EEt6x6_kernel, Ev6x1_kernel,
// csrc/slam_ext/geom_kernels.cu:178-432,863-880,994-1080) for source frames with 1..20 outgoing edges, as TWO kernels.
// Opt-in (VIPE_BA_LIN4=1): parity-green, but on B200 the FMA kernel is still faster (DESIGN.md section 10 has the numbers).
//
//   cw_kernel    the disparity block first: per source pixel C = sum_e J_z^T w J_z, w = sum_e J_z^T w r over the frame's
//                edges (edge order), damping / prior, Q = 1/C -> Q, Q w and sqrt(Q), sqrt(Q) w.  A light pass (40 of the
//                ~150 FMAs per edge-pixel) whose only purpose is that the heavy kernel knows Q before it starts, so that no
//                warp ever waits for another warp's pixels.
//   lin4_kernel  persistent, warp-specialised, one CTA per SM.  The work is a flat list of UNITS (edge, 64-pixel sub-tile),
//                ordered (item = (frame, 256-pixel chunk), sub-tile, edge) and dealt round-robin to the J warps:
//                  10 J warps      one pixel pair per lane.  The unit's inputs (targets/weights rows, disparities, sqrt(Q),
//                                  the edge's constants) arrive by per-lane asynchronous copies (cp.async / LDGSTS) in a
//                                  private 3-stage ring, two units ahead (a dedicated producer warp issuing cp.async.bulk for
//                                  everybody was measured first: its 80 serialised bulk copies per round starved the J
//                                  warps).  Projective transform, residuals, J_j, J_z; the 27 sums of H_jj / v_j / energy
//                                  are reduced across the warp through the consumed ring stage -> epart (one record per
//                                  unit); the pose-disparity vector u (6 per edge-pixel) is scaled by sqrt(Q), split into
//                                  (hi, lo) TF32 planes and stored straight into the tensor core's K-major
//                                  128-byte-swizzled operand buffer of the sub-tile (double-buffered).
//                  MMA warp        one thread: tcgen05.mma kind::tf32 per sub-tile, accumulators in tensor memory; a
//                                  second thread turns the buffer-release mbarrier into counters the J warps poll.
//                  4 flush warps   tcgen05.ld of an item's accumulators, Y = hh + lh + lh^T, -> gpart.
//
// Gram: Y = sum_px Q u u^T = hh + lh + lh^T with u = hi + lo (the lo.lo term, 2^-22 relative, is dropped).  Rows live in
// 64-row groups [hi_A ; lo_A ; hi_B ; lo_B]: group A holds edges 0..9 and, in row 60, sqrt(Q) w (its products with the u rows
// are the Schur right-hand side s_m); group B holds edges 10..19.  Up to 10 edges: ONE MMA per K = 8 step,
// D = [hi_A ; lo_A] hi_A^T (M = 128, N = 64), the 32 K steps of an item rotating over 8 accumulator sets; 11..20 edges: four
// MMAs per K step (AA, BB, AB = [hi_A ; lo_A] hi_B^T and hi_A lo_B^T) over 2 sets.  The sets exist because the tensor core
// TRUNCATES its fp32 accumulator (about -1 ulp of the running sum per MMA) and A - S cancels: measured at C3, the error of
// the reduced system grows linearly with the chain length (2.7e-5 at chains of 8 / 32, 1.4e-5 at 4 / 16, 7.9e-6 = the FMA
// kernel's level at 2 / 8); the flush warps add the sets in round-to-nearest fp32.
#include <type_traits>

#include "ba_common.cuh"
#include "ba_jstep.cuh"
#include "ba_launch.h"
#include "sm100_async.cuh"

namespace vba {
using namespace sm100;
using namespace jmath;

namespace {

constexpr int kNJ = kLin4NJ;
constexpr int kRing = 3;                 // ring stages per J warp (two units in flight behind the one being processed)
constexpr int kNF = 4;                   // flush warps: one per tensor-memory lane quarter
constexpr int kFirstFlush = 1, kFirstJ = kFirstFlush + kNF;
constexpr int kThreads = 32 * (kFirstJ + kNJ);  // warp 0: MMA issuer
constexpr int SUB = kSubTile;            // 64 pixels per unit = K of one MMA batch
constexpr int NSUB = kLin4ItemPx / SUB;  // sub-tiles per item
constexpr int KSTEPS = SUB / 8;
constexpr int GROUP = kLin4GroupDeg;     // edges per 64-row group
constexpr int WROW = 6 * GROUP;          // row of sqrt(Q) w inside group A
constexpr int kArrivals = kLin4MaxDeg;   // arrival count of an operand buffer's "full" barrier (edge 0 makes up for d < max)
constexpr int kSetsS = 8, kSetsL = 2;    // accumulator sets the K steps rotate over (64 / 256 columns each)

constexpr uint32_t kStageBytes = 2048;   // tu, tv, wu, wv, h, sqrt(Q), sqrt(Q) w (256 B each), constants (128 B), descriptor (32 B)
constexpr uint32_t kOffH = 1024, kOffSq = 1280, kOffSqw = 1536, kOffEc = 1792, kOffDesc = 1920;
constexpr uint32_t kKBlockBytes = 256 * 128;          // one 32-pixel K block of an operand buffer: 256 rows x 128 B
constexpr uint32_t kOpBytes = 2 * kKBlockBytes;       // 64 KB

struct Smem4 {
    static constexpr size_t off_op = 0;
    static constexpr size_t off_ring = off_op + 2 * (size_t)kOpBytes;
    static constexpr size_t off_ls = off_ring + (size_t)kNJ * kRing * kStageBytes;
    static constexpr size_t off_bar = off_ls + (size_t)64 * 65 * 4;
    static constexpr int n_bar = 6;
    static constexpr size_t off_tmem = off_bar + (size_t)n_bar * 8;
    static constexpr size_t off_done = off_tmem + 8;  // two completion counters, see the watcher thread
    static constexpr size_t total = off_done + 8;
};

__device__ __forceinline__ void mbar_arrive_n(uint64_t *bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
// per-lane asynchronous copy global -> shared (LDGSTS), 8 bytes; completion through commit / wait groups
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// pair index of edges m <= mp in the frame's Gram record
__device__ __forceinline__ int pair_index(int m, int mp) { return mp * (mp + 1) / 2 + m; }

}  // namespace

// per-edge constants, every value stored twice (the J math works on pixel pairs): R (9), t (3), -t2, stereo flag
// (relSE3 / the stereo special case, geom_kernels.cu:105-113,219-230)
__global__ void edge_const2_kernel(const int *__restrict__ fedge, const int *__restrict__ slot_src, int slot_lo, int nslots,
                                   const float *__restrict__ poses, const int *__restrict__ e_jj, float2 *__restrict__ econst2) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const int e = fedge[slot_lo + s];
    RelPose<float> rp;
    relative_pose<float>(poses, slot_src[slot_lo + s], e_jj[e], rp);
    float2 *c = econst2 + (size_t)e * 16;
#pragma unroll
    for (int q = 0; q < 9; q++) c[q] = splat2(rp.R[q]);
#pragma unroll
    for (int q = 0; q < 3; q++) c[9 + q] = splat2(rp.t[q]);
    c[12] = splat2(-rp.t[2]);
    c[13] = splat2(rp.stereo ? 1.0f : 0.0f);
    c[14] = c[15] = splat2(0.0f);
}

// ================================================================================================================
// The disparity block of every owned frame: grid (HW / 256, frames), one pixel pair per thread.
__global__ void __launch_bounds__(128) cw_kernel(const LinArgs a, const float2 *__restrict__ econst2, float *__restrict__ sqbuf,
                                                 float *__restrict__ sqwbuf) {
    extern __shared__ float2 ecs[];  // [d][16]
    const Tables &tb = a.tb;
    const int k = tb.k_lo + blockIdx.y;
    const int s0 = tb.fptr[k], d = tb.fptr[k + 1] - s0, src = tb.kx[k];
    const int HW = tb.HW, tid = threadIdx.x;
    const int px = blockIdx.x * 256 + 2 * tid;
    const float2 h = __ldg(reinterpret_cast<const float2 *>(a.disps + (size_t)src * HW + px));
    const float2 ds = __ldg(reinterpret_cast<const float2 *>(a.dsens + (size_t)src * HW + px));
    const float2 et = __ldg(reinterpret_cast<const float2 *>(a.eta + (size_t)k * HW + px));
    auto rows = [&](int m, float2 &tu, float2 &tv, float2 &wu, float2 &wv) {
        const int erow = edge_row(tb, s0 + m, tb.fedge[s0 + m]);
        const float2 *t = reinterpret_cast<const float2 *>(a.targets + (size_t)erow * 2 * HW + px);
        const float2 *w = reinterpret_cast<const float2 *>(a.weights + (size_t)erow * 2 * HW + px);
        tu = __ldg(t), tv = __ldg(t + HW / 2), wu = __ldg(w), wv = __ldg(w + HW / 2);
    };
    float2 tu = {0, 0}, tv = {0, 0}, wu = {0, 0}, wv = {0, 0};
    if (d > 0) rows(0, tu, tv, wu, wv);
    for (int i = tid; i < d * 16; i += 128) ecs[i] = __ldg(econst2 + (size_t)tb.fedge[s0 + i / 16] * 16 + (i & 15));
    __syncthreads();
    const JConst kc = make_jconst(a);
    PxPos pos;
    pos.set(px, tb.wd);
    float2 xn, yn;
    pos.normalised(kc, xn, yn);
    float2 C = make_float2(0.f, 0.f), W = make_float2(0.f, 0.f);
    for (int m = 0; m < d; m++) {
        const float2 tu_c = tu, tv_c = tv, wu_c = wu, wv_c = wv;
        if (m + 1 < d) rows(m + 1, tu, tv, wu, wv);
        cw_step(kc, ecs + m * 16, xn, yn, h, tu_c, tv_c, wu_c, wv_c, C, W);
    }
    const int flags = a.opt.frame_flags ? a.opt.frame_flags[k] : 0;
    float2 Q, wz;
    disparity_block(C.x, W.x, h.x, ds.x, et.x, flags, a.opt, Q.x, wz.x);
    disparity_block(C.y, W.y, h.y, ds.y, et.y, flags, a.opt, Q.y, wz.y);
    const size_t o = (size_t)k * HW + px;
    *reinterpret_cast<float2 *>(a.qbuf + o) = Q;
    *reinterpret_cast<float2 *>(a.qwbuf + o) = make_float2(Q.x * wz.x, Q.y * wz.y);
    const float2 sq = make_float2(sqrtf(Q.x), sqrtf(Q.y));
    *reinterpret_cast<float2 *>(sqbuf + o) = sq;
    *reinterpret_cast<float2 *>(sqwbuf + o) = make_float2(sq.x * wz.x, sq.y * wz.y);
}

// ================================================================================================================
__global__ void __launch_bounds__(kThreads, 1) lin4_kernel(const LinArgs a, const Lin4Item *__restrict__ items,
                                                           const Lin4Unit *__restrict__ units, const int *__restrict__ cta_item,
                                                           const int *__restrict__ cta_unit, const float2 *__restrict__ econst2,
                                                           const float *__restrict__ sqbuf, const float *__restrict__ sqwbuf) {
    using L = Smem4;
    extern __shared__ __align__(1024) unsigned char smem[];  // operand buffers first: the swizzle needs 1024-byte alignment
    unsigned char *op = smem + L::off_op;
    unsigned char *ring = smem + L::off_ring;
    float *Ls = reinterpret_cast<float *>(smem + L::off_ls);  // [64][65]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::off_bar);
    uint64_t *op_full = bars;           // [2] J -> MMA: every edge's rows of the sub-tile are written
    uint64_t *op_empty = op_full + 2;   // [2] MMA -> J: the MMAs that read the buffer are done
    uint64_t *acc_full = op_empty + 2;  // MMA -> flush: the item's accumulators are complete
    uint64_t *acc_empty = acc_full + 1; // flush -> MMA: tensor memory has been read
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::off_tmem);
    // op_done[b] = how many times operand buffer b has been released by the tensor core.  The J warps poll these counters
    // instead of the mbarrier: a warp with fewer units than sub-tiles (out-degree < kNJ) does not see every phase, and a parity
    // wait that lags two phases behind never returns.  A watcher thread, which does see every phase, keeps the counters.
    uint32_t *op_done = reinterpret_cast<uint32_t *>(smem + L::off_done);

    const Tables &tb = a.tb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HW = tb.HW;
    const int i0 = cta_item[blockIdx.x], i1 = cta_item[blockIdx.x + 1];
    const int u0 = cta_unit[blockIdx.x], nunits = cta_unit[blockIdx.x + 1] - u0;
    const int cta_g0 = nunits > 0 ? units[u0].g : 0;  // the list's running sub-tile number of this CTA's first sub-tile

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(op_full + i, kArrivals);
            mbar_init(op_empty + i, 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, kNF);
        op_done[0] = op_done[1] = 0;
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    // unused operand rows must hold finite numbers (they only feed accumulator entries nobody reads)
    for (int i = tid; i < (int)(2 * kOpBytes / 16); i += kThreads) reinterpret_cast<float4 *>(op)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ============================================================ MMA issuer (one thread)
        if (lane == 0 && i1 > i0) {
            const uint32_t idesc = make_idesc_tf32(128, 64);
            uint32_t g = 0;
            int d_next = items[i0].d;
            for (int it = i0; it < i1; it++) {
                const int d = d_next;
                if (it + 1 < i1) d_next = items[it + 1].d;
                const bool large = d > GROUP;
                const uint32_t n = (uint32_t)(it - i0);
                for (int t = 0; t < NSUB; t++, g++) {
                    const uint32_t buf = g & 1;
                    const uint32_t sbase = smem_u32(op) + buf * kOpBytes;
                    mbar_wait_relaxed(op_full + buf, (g >> 1) & 1);
                    if (t == 0 && n > 0) mbar_wait_relaxed(acc_empty, (n - 1) & 1);  // the previous item's sums have been read
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ks++) {
                        const int kk = t * KSTEPS + ks;
                        const uint32_t ab_addr = sbase + (ks >> 2) * kKBlockBytes + (ks & 3) * 32;
                        const uint64_t dA = make_desc_k_sw128(ab_addr);  // rows 0..127 = [hi_A ; lo_A]; as the B operand: hi_A
                        if (!large) {
                            mma_tf32(tmem + (kk % kSetsS) * 64, dA, dA, idesc, kk >= kSetsS);
                        } else {
                            const uint64_t dB = make_desc_k_sw128(ab_addr + 128 * 128);   // rows 128..255 = [hi_B ; lo_B]; B operand: hi_B
                            const uint64_t dlB = make_desc_k_sw128(ab_addr + 192 * 128);  // rows 192..255 = lo_B
                            const uint32_t acc = tmem + (kk % kSetsL) * 256;
                            const bool accum = kk >= kSetsL;
                            mma_tf32(acc + 0, dA, dA, idesc, accum);     // AA: [hi_A ; lo_A] hi_A^T
                            mma_tf32(acc + 64, dB, dB, idesc, accum);    // BB: [hi_B ; lo_B] hi_B^T
                            mma_tf32(acc + 128, dA, dB, idesc, accum);   // AB: [hi_A ; lo_A] hi_B^T
                            mma_tf32(acc + 192, dA, dlB, idesc, accum);  // hi_A lo_B^T (lanes 0..63)
                        }
                    }
                    mma_commit(op_empty + buf);
                }
                mma_commit(acc_full);
            }
        } else if (lane == 1) {
            // watcher: turns the completions of op_empty into monotonic counters
            const uint32_t nsub_total = (uint32_t)(i1 - i0) * NSUB;
            for (uint32_t g = 0; g < nsub_total; g++) {
                mbar_wait_relaxed(op_empty + (g & 1), (g >> 1) & 1);
                asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(op_done + (g & 1))), "r"((g >> 1) + 1) : "memory");
            }
        }
        __syncwarp();
    } else if (warp < kFirstJ) {
        // ============================================================ flush warps
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        const uint32_t tm_lane = (uint32_t)(32 * q) << 16;
        const int r = 32 * (q & 1) + lane;  // row inside the 64-row group (q < 2: hi rows, q >= 2: lo rows)
        const int m = r / 6, rr = r - 6 * m;
        Lin4Item nxt;
        if (i1 > i0) nxt = items[i0];
        // Y = hh + lh + lh^T of one 64-row group (edges ebase .. ebase + dn - 1) -> the pair blocks of the record
        auto emit_sym = [&](float (&v)[64], float *gp, int ebase, int dn, int npairs, bool with_s) {
            named_bar(1, 32 * kNF);  // the previous readers of Ls are done
            if (q >= 2) {
#pragma unroll
                for (int c = 0; c < 64; c++) Ls[r * 65 + c] = v[c];
            }
            named_bar(1, 32 * kNF);
            if (q < 2) {
#pragma unroll
                for (int c = 0; c < 64; c++) v[c] += Ls[r * 65 + c] + Ls[c * 65 + r];
                if (m < dn) {
#pragma unroll
                    for (int mp = 0; mp < GROUP; mp++) {
                        if (mp >= m && mp < dn) {
                            float2 *dst = reinterpret_cast<float2 *>(gp + (size_t)pair_index(ebase + m, ebase + mp) * 36 + rr * 6);
                            dst[0] = make_float2(v[6 * mp + 0], v[6 * mp + 1]);
                            dst[1] = make_float2(v[6 * mp + 2], v[6 * mp + 3]);
                            dst[2] = make_float2(v[6 * mp + 4], v[6 * mp + 5]);
                        }
                    }
                    if (with_s) gp[(size_t)npairs * 36 + 6 * m + rr] = v[WROW];
                }
            }
        };
        for (int it = i0; it < i1; it++) {
            const Lin4Item im = nxt;
            if (it + 1 < i1) nxt = items[it + 1];
            const int d = im.d;
            const bool large = d > GROUP;
            const int dA = large ? GROUP : d, dB = d - dA;
            const int npairs = d * (d + 1) / 2, rec = npairs * 36 + 6 * d;
            float *gp = a.gpart + tb.gbase[im.k] + (size_t)im.chunk * rec;
            const uint32_t n = (uint32_t)(it - i0);
            const uint32_t acc = tmem + tm_lane;
            mbar_wait_relaxed(acc_full, n & 1);
            tc_fence_after();
            float v[64];
            // v (+)= columns [col, col + 64) of this thread's lane: 4 loads in flight, one wait
            auto load_cols = [&](uint32_t col, bool add) {
                float t16[4][16];
#pragma unroll
                for (int j = 0; j < 4; j++) tmem_ld16(acc + col + 16 * j, t16[j]);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int i = 0; i < 16; i++) v[16 * j + i] = add ? v[16 * j + i] + t16[j][i] : t16[j][i];
            };
            if (!large) {
#pragma unroll
                for (int s = 0; s < kSetsS; s++) load_cols(64 * s, s > 0);
                tc_fence_before();
                warp_arrive(acc_empty, lane);
                emit_sym(v, gp, 0, dA, npairs, true);
            } else {
                load_cols(0, false);
                load_cols(256, true);
                emit_sym(v, gp, 0, dA, npairs, true);
                load_cols(64, false);
                load_cols(256 + 64, true);
                emit_sym(v, gp, GROUP, dB, npairs, false);
                // AB: rows of group A x columns of group B: hi.hi + lo.hi (lanes 64..127) + hi.lo
                load_cols(128, false);
                load_cols(256 + 128, true);
                if (q < 2) {
                    load_cols(192, true);
                    load_cols(256 + 192, true);
                }
                tc_fence_before();
                warp_arrive(acc_empty, lane);
                named_bar(1, 32 * kNF);
                if (q >= 2) {
#pragma unroll
                    for (int c = 0; c < 64; c++) Ls[r * 65 + c] = v[c];
                }
                named_bar(1, 32 * kNF);
                if (q < 2) {
#pragma unroll
                    for (int c = 0; c < 64; c++) v[c] += Ls[r * 65 + c];
                    if (m < dA) {
#pragma unroll
                        for (int mp = 0; mp < GROUP; mp++) {
                            if (mp < dB) {
                                float2 *dst = reinterpret_cast<float2 *>(gp + (size_t)pair_index(m, GROUP + mp) * 36 + rr * 6);
                                dst[0] = make_float2(v[6 * mp + 0], v[6 * mp + 1]);
                                dst[1] = make_float2(v[6 * mp + 2], v[6 * mp + 3]);
                                dst[2] = make_float2(v[6 * mp + 4], v[6 * mp + 5]);
                            }
                        }
                    } else if (r == WROW) {  // the sqrt(Q) w row against group B's columns: s_m of edges 10..
#pragma unroll
                        for (int c = 0; c < 6 * GROUP; c++)
                            if (c < 6 * dB) gp[(size_t)npairs * 36 + 6 * GROUP + c] = v[c];
                    }
                }
            }
        }
    } else {
        // ============================================================ J warps
        const int jw = warp - kFirstJ;
        const JConst kc = make_jconst(a);
        const float wdf = (float)tb.wd;
        const int ntile_e = tb.ntile_e;
        unsigned char *myring = ring + (size_t)jw * kRing * kStageBytes;
        const int my_n = nunits > jw ? (nunits - 1 - jw) / kNJ + 1 : 0;  // units of this warp: u0 + jw + kNJ i
        // per-lane bases of the asynchronous copies (every lane copies its own pixel pair)
        const float *t_lane = a.targets + 2 * lane, *w_lane = a.weights + 2 * lane;
        const float *h_lane = a.disps + 2 * lane, *sq_lane = sqbuf + 2 * lane, *sqw_lane = sqwbuf + 2 * lane;
        const int *unit_words = reinterpret_cast<const int *>(units + u0 + jw) + (lane & 7);
        // descriptor of unit i: lane l < 8 holds its l-th word
        auto load_desc = [&](int i) -> int { return (i < my_n && lane < 8) ? __ldg(unit_words + (size_t)i * (kNJ * 8)) : 0; };
        // asynchronous copies of unit i's inputs into the ring stage at `stg`
        auto issue = [&](int i, int dreg, unsigned char *stg) {
            if (i < my_n) {
                const int slot = __shfl_sync(0xffffffffu, dreg, 0), edge = __shfl_sync(0xffffffffu, dreg, 1);
                const int srcpx = __shfl_sync(0xffffffffu, dreg, 2), kpx = __shfl_sync(0xffffffffu, dreg, 3);
                const int ed = __shfl_sync(0xffffffffu, dreg, 4), px0 = __shfl_sync(0xffffffffu, dreg, 6);
                const size_t toff = (size_t)(unsigned)edge_row(tb, slot, edge) * (unsigned)(2 * HW) + (unsigned)px0;
                unsigned char *dst = stg + 8 * lane;
                cp_async8(dst, t_lane + toff);
                cp_async8(dst + 256, t_lane + toff + HW);
                cp_async8(dst + 512, w_lane + toff);
                cp_async8(dst + 768, w_lane + toff + HW);
                cp_async8(dst + kOffH, h_lane + srcpx);
                cp_async8(dst + kOffSq, sq_lane + kpx);
                if ((ed & 255) == 0) cp_async8(dst + kOffSqw, sqw_lane + kpx);
                if (lane < 16) cp_async8(stg + kOffEc + 8 * lane, econst2 + (size_t)edge * 16 + lane);
                if (lane < 8) *reinterpret_cast<int *>(stg + kOffDesc + 4 * lane) = dreg;
            }
            cp_async_commit();
        };
        // per-lane constants of the operand stores: K block of the pixel pair, position inside the 16-byte chunk, chunk index
        const uint32_t lane_off = (uint32_t)(lane >> 4) * kKBlockBytes + (lane & 1) * 8;
        const uint32_t chunk_x = (uint32_t)((lane & 15) >> 1) << 4;
        // one value pair -> (hi, lo) TF32 planes; `off` = byte offset of the row's 128-byte line, r7 = row % 8 (the swizzle key)
        auto put = [&](unsigned char *base, uint32_t off, uint32_t r7, float2 val) {
            float2 hi;
            hi.x = __uint_as_float((__float_as_uint(val.x) + 0x1000u) & 0xFFFFE000u);
            hi.y = __uint_as_float((__float_as_uint(val.y) + 0x1000u) & 0xFFFFE000u);
            const float2 lo = ffma2(hi, kc.m1, val);
            unsigned char *p = base + off + (chunk_x ^ (r7 << 4));
            *reinterpret_cast<float2 *>(p) = hi;
            *reinterpret_cast<float2 *>(p + 64 * 128) = lo;  // row + 64
        };
        int dr2 = load_desc(2);
        {
            const int dr0 = load_desc(0), dr1 = load_desc(1);
            issue(0, dr0, myring);
            issue(1, dr1, myring + kStageBytes);
        }
        unsigned char *stg = myring, *stg2 = myring + 2 * kStageBytes;  // stage of unit i, stage of unit i + 2
        for (int i = 0; i < my_n; i++) {
            issue(i + 2, dr2, stg2);  // that stage was last read by unit i - 1 (behind the warp barriers of that iteration)
            dr2 = load_desc(i + 3);
            cp_async_wait<2>();
            __syncwarp();  // the constants and the descriptor were written by other lanes
            const int4 d0 = *reinterpret_cast<const int4 *>(stg + kOffDesc);
            const int4 d1 = *reinterpret_cast<const int4 *>(stg + kOffDesc + 16);
            const int slot = d0.x, e = d1.x & 255, d = d1.x >> 8, px0 = d1.z;
            const uint32_t g = (uint32_t)(d1.w - cta_g0);
            const float *sf = reinterpret_cast<const float *>(stg) + 2 * lane;
            const float2 tu = *reinterpret_cast<const float2 *>(sf), tv = *reinterpret_cast<const float2 *>(sf + 64);
            const float2 wu = *reinterpret_cast<const float2 *>(sf + 128), wv = *reinterpret_cast<const float2 *>(sf + 192);
            const float2 h = *reinterpret_cast<const float2 *>(sf + kOffH / 4);
            const float2 sq = *reinterpret_cast<const float2 *>(sf + kOffSq / 4);
            const float2 *c = reinterpret_cast<const float2 *>(stg + kOffEc);
            const bool stereo = c[13].x != 0.0f;
            PxPos pos;
            pos.row = (float)(d1.y >> 16), pos.col = (float)(d1.y & 0xffff);
            pos.advance((float)(2 * lane), wdf);
            float2 xn, yn;
            pos.normalised(kc, xn, yn);
            float2 hh[kEdgeVals], u[6];
            if (!stereo) {
                j_step_u(kc, c, xn, yn, h, tu, tv, wu, wv, hh, u);
            } else {  // stereo edges contribute nothing beyond C and w (:329,367)
#pragma unroll
                for (int q = 0; q < 6; q++) u[q] = make_float2(0.f, 0.f);
            }

            // the buffer of sub-tile g is free once the MMAs of sub-tile g - 2 are done
            const uint32_t buf = g & 1;
            if (g >= 2) {
                uint32_t cnt, spins = 0;
                do {
                    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(cnt) : "r"(smem_u32(op_done + buf)) : "memory");
                    if (++spins > (1u << 26)) mbar_timeout(op_empty + buf, g);
                } while (cnt < (g >> 1));
            }

            // sqrt(Q) u, true signs, as (hi, lo) TF32 planes -> rows of this edge in the sub-tile's operand buffer.  Rows
            // 6 eg .. 6 eg + 5 of a 64-row group: (6 eg) % 8 takes four values, so the swizzle keys are compile-time constants
            // behind a four-way switch.
            const int eg = e < GROUP ? e : e - GROUP;
            unsigned char *opb = op + (size_t)buf * kOpBytes + lane_off + (e < GROUP ? 0u : 128u * 128u) + (uint32_t)((6 * eg) >> 3) * 1024u;
            const float2 nsq = make_float2(-sq.x, -sq.y);
            auto rows = [&](auto c0_tag) {
                constexpr uint32_t c0 = decltype(c0_tag)::value;
#pragma unroll
                for (uint32_t q = 0; q < 6; q++) {
                    const uint32_t rq = c0 + q;
                    put(opb, (rq >> 3) * 1024u + (rq & 7u) * 128u, rq & 7u, fmul2(u[q], (q == 2 || q == 3 || q == 5) ? nsq : sq));
                }
            };
            switch (eg & 3) {
                case 0: rows(std::integral_constant<uint32_t, 0>{}); break;
                case 1: rows(std::integral_constant<uint32_t, 6>{}); break;
                case 2: rows(std::integral_constant<uint32_t, 4>{}); break;
                default: rows(std::integral_constant<uint32_t, 2>{}); break;
            }
            if (e == 0) {
                const float2 sqw = *reinterpret_cast<const float2 *>(sf + kOffSqw / 4);
                put(op + (size_t)buf * kOpBytes + lane_off, (uint32_t)(WROW >> 3) * 1024u + (uint32_t)(WROW & 7) * 128u, (uint32_t)(WROW & 7), sqw);
            }
            fence_async_smem();
            __syncwarp();  // also: every lane is done with the stage, which now serves as the scratch of the reduction
            if (lane == 0) mbar_arrive_n(op_full + buf, e == 0 ? (uint32_t)(kArrivals - d + 1) : 1u);

            // the 27 sums of this unit: lanes -> one record, true signs
            float *rec = a.epart + ((size_t)slot * ntile_e + (px0 >> 6)) * kEdgeStride;
            if (!stereo) {
                write_edge_record_smem(hh, lane, reinterpret_cast<float *>(stg), rec);
            } else if (lane < kEdgeVals) {
                rec[lane] = 0.0f;
            }
            __syncwarp();
            unsigned char *nx = stg + kStageBytes;  // rotate the ring
            stg2 = stg;
            stg = nx == myring + kRing * kStageBytes ? myring : nx;
        }
        cp_async_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ================================================================================================================
bool lin4_supported(int HW, int dmax) { return HW % 256 == 0 && HW % kLin4ItemPx == 0 && dmax >= 1 && dmax <= kLin4MaxDeg; }
size_t lin4_smem_bytes() { return Smem4::total; }

cudaError_t launch_lin4(const LinArgs &a, const Lin4Launch &l, cudaStream_t st) {
    if (l.nframes <= 0) return cudaSuccess;
    using L = Smem4;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(lin4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    if (l.nslots > 0)
        edge_const2_kernel<<<(l.nslots + 127) / 128, 128, 0, st>>>(a.tb.fedge, l.slot_src, l.slot_lo, l.nslots, a.poses, a.tb.e_jj,
                                                                   l.econst2);
    dim3 grid(a.tb.HW / 256, l.nframes);  // its own 256-pixel tiles
    cw_kernel<<<grid, 128, (size_t)l.dmax * 16 * sizeof(float2), st>>>(a, l.econst2, l.sqbuf, l.sqwbuf);
    if (l.grid > 0)
        lin4_kernel<<<l.grid, kThreads, L::total, st>>>(a, l.items, l.units, l.cta_item, l.cta_unit, l.econst2, l.sqbuf, l.sqwbuf);
    return cudaGetLastError();
}

}  // namespace vba
