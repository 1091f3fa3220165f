// TMA-fed motion-only linearisation ("lin3_motion"): stages 1-2 of the path alone (projective transform, J_j, per-edge
// H_jj / v_j / energy; reference: projective_transform_kernel, csrc/slam_ext/geom_kernels.cu:178-432) as a persistent kernel
// whose warps fetch their own targets/weights rows with cp.async.bulk (1-D TMA copies completing on mbarriers).  The full
// linearisation with the tensor-core Schur Gram lives in ba_lin4.cu.
#include "ba_common.cuh"
#include "ba_launch.h"
#include "sm100_async.cuh"
#include "ba_jstep.cuh"

namespace vba {
using namespace sm100;
using namespace jmath;

namespace {

constexpr int TILE = kSubTile;  // pixels per warp step (one pixel pair per lane)
constexpr int STAGE_PX = 256;   // pixels per TMA stage (4 steps)

}  // namespace

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

bool lin3_supported(int HW, int chunk_px) { return chunk_px >= STAGE_PX && chunk_px % STAGE_PX == 0 && HW % chunk_px == 0; }

}  // namespace vba
