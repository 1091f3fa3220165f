// Shared declarations of the dense-BA kernels (sm_100a).  See DESIGN.md for the data layout.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vba {

constexpr float kMinDepth = 0.25f;      // reference: geom_kernels.cu:33
constexpr float kWeightScale = 0.001f;  // reference: geom_kernels.cu:304-305
constexpr float kAlpha = 0.05f;         // reference: geom_kernels.cu:1359
constexpr float kStereoBaseline = -0.1f;  // reference: geom_kernels.cu:222

constexpr int kEdgeVals = 27;    // 20 unique non-zero entries of H_jj, 6 of v_j, 1 energy
constexpr int kEdgeStride = 28;  // padded stride of the per-(edge,tile) partial record
constexpr int kEcStride = 16;    // floats per edge constant record in shared memory

// Index tables + sizes shared by every kernel of one plan (device pointers live in the workspace head).
struct Tables {
    const int *kx;      // [K]   source-frame id of kx position k
    const int *fptr;    // [K+1] CSR of edges by source frame (slots)
    const int *fedge;   // [E]   slot -> edge id
    const int *e_jj;    // [E]   edge id -> target frame
    const long long *gbase;  // [K+1] float offset of frame k's Gram partials (already multiplied by ntile)
    const long long *mbase;  // [K+1] double offset of frame k's M scratch
    // A plan holds C >= 1 independent problems (C > 1: batched small problems, e.g. many motion-only clips).
    const int *pose_slot;        // [N] index of the pose in ITS problem's reduced system, -1 if fixed
    const int *pose_row;         // [N] row of the pose in the concatenated dx output, -1 if fixed
    const int *pose_sys;         // [N] position of the pose in ITS problem's reduced system (a fill-reducing order), -1 if fixed
    const int *frame_prob;       // [K] problem of kx frame k
    const long long *prob_hoff;  // [C] double offset of the problem's [H ; b ; diag(A)] block
    const int *prob_npad;        // [C] padded system size of the problem
    const int *prob_n;           // [C] system size 6 * (free poses)
    const int *prob_row0;        // [C] first dx row of the problem
    int K, E, N, HW, wd, ntile, k_lo, k_hi, C;
    int ntile_e = 0;  // partial records per edge slot in epart (= ntile, except behind ba_lin4.cu: one per 64-pixel unit)
    // owner-only pixel inputs (multi-GPU): targets/weights hold just the rows of this rank's edges, in CSR slot order
    int rows_by_slot = 0, slot_lo = 0;
};

// row of edge `e` (CSR slot `slot`) in the targets/weights arrays
__device__ __forceinline__ int edge_row(const Tables &tb, int slot, int e) { return tb.rows_by_slot ? slot - tb.slot_lo : e; }

// Semantic switches (defaults = the reference's CUDA BA; the other values serve the conventions of vipe/slam's Python BA,
// SURVEY.md section 8(a')).  Mirrors vipe_ba_options in include/vipe_ba.h.
struct Options {
    float min_depth = kMinDepth;
    int depth_strict = 0;   // 0: invalid iff z < min_depth (geom_kernels.cu:301); 1: valid iff z > min_depth (geom.py:263)
    float alpha = kAlpha;
    int sensor_mode = 0;    // 0: per-pixel mask (geom_kernels.cu:1361-1369); 1: per-frame gate, every pixel (terms.py:244-300)
    float eta_scale = 1.0f, eta_bias = 0.0f;  // disparity damping = eta_scale * eta + eta_bias
    float dz_max = 3.0e38f;  // dz > dz_max -> 0 (retractor.py:41)
    int renorm_quat = 0;     // renormalise the quaternion after the pose retraction (lietorch so3.h:36-38)
    int damp_on_pose_hessian = 0;  // 0: lm scales diag(A - S) (geom_kernels.cu:1176); 1: lm scales diag(A), the pose Hessian before
                                   //    the Schur complement (solver.py:161-164)
    int backsub_all_poses = 0;     // 0: EvT6x1 drops pose index 0 (geom_kernels.cu:1089, Q4); 1: every free pose reaches dz (solver.py:182)
    const unsigned char *frame_flags = nullptr;  // DEV [K]: bit 0 sensor gate (mode 1), bit 1 disparity fixed
    // Focal length as one more variable, shared by fx and fy (pinhole IntrinsicsRetractor, retractor.py:51-62; the
    // backend's optimize_intrinsics, buffer.py:496-498).  Its row/column borders the reduced camera system.
    int optimize_focal = 0;
    float focal_jscale = 1.0f;               // d(intrinsics passed to the kernels)/d(variable) = 1/intrinsics_factor (terms.py:186,224)
    float focal_lm = 0.0f, focal_ep = 0.0f;  // LHS_ff += focal_lm * H_ff + focal_ep (buffer.py:496, solver.py:161-164)
};

__device__ __forceinline__ bool depth_valid(float z, const Options &o) { return o.depth_strict ? (z > o.min_depth) : !(z < o.min_depth); }

// disparity block of one pixel: C, w from the edge sums + prior/damping -> (Q, w)
__device__ __forceinline__ void disparity_block(float Cacc, float Wacc, float h, float ds, float et, int flags, const Options &o,
                                                float &Q, float &W) {
    float C;
    if (o.sensor_mode == 0) {
        const bool mk = ds > 0.0f;
        C = Cacc + (mk ? o.alpha : fmaf(o.eta_scale, et, o.eta_bias));
        W = Wacc - (mk ? o.alpha * (h - ds) : 0.0f);
    } else {
        const bool gate = (flags & 1) != 0;
        C = Cacc + fmaf(o.eta_scale, et, o.eta_bias) + (gate ? o.alpha : 0.0f);
        W = Wacc - (gate ? o.alpha * (h - ds) : 0.0f);
    }
    Q = (flags & 2) ? 0.0f : __fdiv_rn(1.0f, C);  // a fixed disparity is eliminated with an infinite block: Q = 0
}

struct LinArgs {
    Tables tb;
    Options opt;
    const float *poses, *disps, *intr, *dsens, *targets, *weights, *eta;
    float *epart;  // [E_slots][ntile][kEdgeStride]
    float *gpart;  // per frame: [ntile][npairs*36 + 6*d]
    float *qbuf;   // [K][HW]  Q = 1/C
    float *qwbuf;  // [K][HW]  Q*w
    const int *flist = nullptr;  // DEV: kx positions of the frames this launch covers (null: k_lo + blockIdx.y)
    // staging of the pose-disparity vectors u in GLOBAL memory, for source frames with more outgoing edges than the
    // shared-memory staging buffer holds: [owned edge slots][ntile][6][TILE] floats
    float *uglobal = nullptr;
};

constexpr int kFocalNT = 256;     // the focal pass tiles a frame into 256-pixel tiles of its own
constexpr int kFocalStride = 16;  // per-(edge, focal tile) record: J_j^T w J_f (6), J_f^T w J_f, J_f^T w r, sum_px Q u_f u_m (6)

// Focal-length pass (runs after the linearisation when Options::optimize_focal): everything that involves J_f.
struct FocalArgs {
    Tables tb;
    Options opt;
    const float *poses, *disps, *intr, *targets, *weights;
    const float *qbuf, *qwbuf;  // from the linearisation
    float *fpart;               // [E_slots][ntile_f][kFocalStride]
    float *ffpart;              // [K][ntile_f][2]: sum_px Q u_f^2, sum_px Q w u_f
    float *ufbuf;               // [K][HW] u_f = sum_edges J_f^T w J_z (the focal/disparity coupling, kept for the back-substitution)
    int ntile_f, motion_only;
};

constexpr int kReduceDoubles = 36 + 36 + 27 + 6 + 14 + 6 + 1;  // G, T, summed records, g, focal records, border column, target index

struct ReduceArgs {
    Tables tb;
    const float *poses;
    const float *epart, *gpart;
    const float *fpart = nullptr, *ffpart = nullptr;  // focal pass partials (null: focal length fixed)
    int ntile_f = 0, focal_row = 0;
    float focal_lm = 0.0f;
    double *msc;   // M scratch (used when a frame's blocks do not fit shared memory)
    int msc_smem = 0;  // 1: M lives in the kernel's shared memory (set by launch_frame_reduce)
    double *fscratch = nullptr;  // per-edge working set [E_owned][kReduceDoubles] in global memory, used when a frame's does not
                                 // fit shared memory (hub frames; set by launch_frame_reduce when needed)
    int fs_global = 0;
    // Deterministic assembly: every frame writes its 6x6 block / 6-vector contributions to fixed slots of `cblk` / `cvec`
    // (`cvec2`: the same slots for diag(A)); assemble_kernel then sums the slots of each destination in a fixed order.
    double *cblk = nullptr;           // [sum over owned frames of (npairs + d + 1)][36]
    double *cvec = nullptr, *cvec2 = nullptr;  // [sum over owned frames of (d + 1)][6]
    const long long *cbase = nullptr;  // [K] first block slot of frame k
    const long long *vbase = nullptr;  // [K] first vector slot of frame k
    double *hsys;  // per problem: [n*n] row-major lower triangle (+ full diagonal blocks), then b [n], then diag(A) [n]
                   //   (diag(A): the pose Hessian alone, before the Schur complement, for damp_on_pose_hessian)
    int motion_only;
};

// Destinations of the deterministic assembly (built by the plan): dst_off = element (0,0) of the 6x6 block / first of the 6 vector
// entries inside the system buffer, src = contribution slots (blocks: slot * 2 + transposed).
struct AssembleArgs {
    const double *cblk, *cvec, *cvec2;
    double *sys;
    const long long *bdst_off;  // [nb]
    const int *bdst_ld;         // [nb] row stride (npad of the problem)
    const int *bsrc_ptr;        // [nb + 1]
    const int *bsrc;            // [...]
    int nb;
    const long long *vdst_off;  // [nv] offset of the rhs entries; diag(A) lives vdst_adiag[] further
    const int *vdst_adiag;      // [nv] distance from the rhs entries to the diag(A) entries
    const int *vsrc_ptr;        // [nv + 1]
    const int *vsrc;            // [...]
    int nv;
};

struct BackArgs {
    Tables tb;
    Options opt;
    const float *poses, *intr, *weights;
    float *disps;
    const float *qbuf, *qwbuf;
    const float *dx;  // [P][6] fp32
    float *dz_out;    // [K][HW]
    const float *ufbuf = nullptr;  // focal/disparity coupling (null: focal length fixed)
    int focal_row = 0;             // index of the focal step in dx
};

// ------------------------------------------------------------------------------------------------
// per-edge relative pose.  Same algebra as relSE3/actSO3 (geom_kernels.cu:69-114) written as a matrix:
// actSO3(q, X) = X + 2w (v x X) + 2 v x (v x X) is linear in X for ANY q (unit or not), and its matrix is
// the usual quaternion rotation matrix.  T is float or double.
template <typename T>
struct RelPose {
    T R[9];
    T t[3];
    bool stereo;
};

template <typename T>
__device__ __forceinline__ void quat_to_mat(const T *q, T *R) {
    const T x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = T(1) - T(2) * (y * y + z * z);
    R[1] = T(2) * (x * y - w * z);
    R[2] = T(2) * (x * z + w * y);
    R[3] = T(2) * (x * y + w * z);
    R[4] = T(1) - T(2) * (x * x + z * z);
    R[5] = T(2) * (y * z - w * x);
    R[6] = T(2) * (x * z - w * y);
    R[7] = T(2) * (y * z + w * x);
    R[8] = T(1) - T(2) * (x * x + y * y);
}

template <typename T>
__device__ __forceinline__ void relative_pose(const float *__restrict__ poses, int i, int j, RelPose<T> &rp) {
    rp.stereo = (i == j);
    if (rp.stereo) {  // geom_kernels.cu:219-230
        rp.R[0] = rp.R[4] = rp.R[8] = T(1);
        rp.R[1] = rp.R[2] = rp.R[3] = rp.R[5] = rp.R[6] = rp.R[7] = T(0);
        rp.t[0] = T(kStereoBaseline);
        rp.t[1] = rp.t[2] = T(0);
        return;
    }
    T ti[3], qi[4], tj[3], qj[4], qij[4];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        ti[k] = T(poses[7 * i + k]);
        tj[k] = T(poses[7 * j + k]);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        qi[k] = T(poses[7 * i + 3 + k]);
        qj[k] = T(poses[7 * j + 3 + k]);
    }
    // qij = qj * conj(qi)   (geom_kernels.cu:105-108)
    qij[0] = -qj[3] * qi[0] + qj[0] * qi[3] - qj[1] * qi[2] + qj[2] * qi[1];
    qij[1] = -qj[3] * qi[1] + qj[1] * qi[3] - qj[2] * qi[0] + qj[0] * qi[2];
    qij[2] = -qj[3] * qi[2] + qj[2] * qi[3] - qj[0] * qi[1] + qj[1] * qi[0];
    qij[3] = qj[3] * qi[3] + qj[0] * qi[0] + qj[1] * qi[1] + qj[2] * qi[2];
    quat_to_mat(qij, rp.R);
    // tij = tj - Rij ti   (geom_kernels.cu:110-113)
#pragma unroll
    for (int k = 0; k < 3; k++) rp.t[k] = tj[k] - (rp.R[3 * k] * ti[0] + rp.R[3 * k + 1] * ti[1] + rp.R[3 * k + 2] * ti[2]);
}

// G (6x6, row-major) with J_i = G J_j, i.e. G = -M where M X = adjSE3(tij, qij, X) (geom_kernels.cu:88-102,332-333):
//   M = [[R^T, 0], [-R^T [t]x, R^T]]   =>   G = [[-R^T, 0], [R^T [t]x, -R^T]]
template <typename T>
__device__ __forceinline__ void adjoint_G(const RelPose<T> &rp, T *G) {
    const T *R = rp.R;
    const T *t = rp.t;
    // [t]x
    const T tx[9] = {T(0), -t[2], t[1], t[2], T(0), -t[0], -t[1], t[0], T(0)};
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const T rt = R[3 * c + r];  // R^T[r][c]
            G[6 * r + c] = -rt;
            G[6 * r + 3 + c] = T(0);
            G[6 * (r + 3) + 3 + c] = -rt;
            // (R^T [t]x)[r][c] = sum_k R[k][r] * tx[k][c]
            G[6 * (r + 3) + c] = R[0 + r] * tx[0 + c] + R[3 + r] * tx[3 + c] + R[6 + r] * tx[6 + c];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Warp "transpose" reductions: V values live in every lane; afterwards lane l holds the sum over all
// 32 lanes of value (l mod V).  Costs V-1 (+ log2(32/V)) shuffles instead of 5V.
template <int V>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[V], int lane) {
    static_assert(V == 32 || V == 16 || V == 8 || V == 4 || V == 2, "power of two");
#pragma unroll
    for (int s = V / 2; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; i++) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    float r = v[0];
#pragma unroll
    for (int s = V; s < 32; s <<= 1) r += __shfl_xor_sync(0xffffffffu, r, s);
    return r;
}

// Blackwell packed fp32 math (FFMA2/FMUL2): two lanes of work per issue slot.  The FMA pipe rate is unchanged
// (measured 115 scalar FMA/clk/SM either way) but the kernels here are issue-bound, not pipe-bound.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}

// decode pair index p -> (m, mp) with m <= mp, p = mp*(mp+1)/2 + m
__device__ __forceinline__ void decode_pair(int p, int &m, int &mp) {
    int c = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
    while ((c + 1) * (c + 2) / 2 <= p) c++;
    while (c * (c + 1) / 2 > p) c--;
    mp = c;
    m = p - c * (c + 1) / 2;
}

// position of the non-zero unique entries of H_jj (lower triangle, (1,0) is structurally zero) in the
// 27-value record:  slot(r,c) for r >= c.
__host__ __device__ __forceinline__ int hslot(int r, int c) {
    // order: (0,0) (1,1) (2,0) (2,1) (2,2) (3,0) (3,1) (3,2) (3,3) (4,0)..(4,4) (5,0)..(5,5)
    if (r == 0) return 0;
    if (r == 1) return c == 1 ? 1 : -1;
    return 2 + (r * (r + 1) / 2 - 3) + c;  // r=2 -> base 2, r=3 -> 5, r=4 -> 9, r=5 -> 14
}

}  // namespace vba
