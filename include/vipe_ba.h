/*
 * vipe_ba.h -- C ABI of the B200-native dense bundle adjustment (libvipe_ba.so).
 *
 * This is the drop-in boundary for ONE reference operator:
 *
 *     slam_ext.ba(poses, disps, intrinsics, disps_sens, targets, weights, eta, ii, jj,
 *                 t0, t1, iterations, lm, ep, motion_only) -> [dx, dz]
 *
 * bound in the reference at csrc/slam_ext/slam.cpp:24-27,32 (pybind) and implemented by
 * `ba_cuda`, csrc/slam_ext/geom_kernels.cu:1283-1404.  The entry points below are what a binding for
 * that operator calls; vipe_b200/ext/slam_ext.py is such a binding (ctypes + torch tensors) and
 * INTEGRATION.md shows the one a ViPE maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch types.  Device pointers are marked DEV, host pointers HOST.
 *  - all floating point device arrays are fp32, contiguous, in the reference's layouts
 *    (geom_kernels.cu:1287-1299): poses[N,7]=(tx,ty,tz,qx,qy,qz,qw), disps[N,ht,wd], intrinsics[4],
 *    disps_sens[N,ht,wd], targets/weights[E,2,ht,wd], eta[K,ht,wd].
 *  - poses and disps are updated IN PLACE (geom_kernels.cu:1353,1393,1397); everything else is read-only.
 *  - every function returns 0 on success, non-zero on error; vipe_ba_last_error() describes the last
 *    failure on the calling thread.  A failed Cholesky factorisation is NOT an error: like the reference
 *    (geom_kernels.cu:1181-1188) it yields dx = 0 for that iteration.
 *  - nothing in this library synchronises the stream or allocates device memory: the caller supplies one
 *    workspace (size from vipe_ba_workspace_bytes) and a cudaStream_t (passed as void*).
 */
#ifndef VIPE_BA_H_
#define VIPE_BA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vipe_ba_plan vipe_ba_plan;

/* ABI version of this header (bumped on any signature change). */
int vipe_ba_abi_version(void);
const char *vipe_ba_last_error(void);

/*
 * Index bookkeeping of one BA problem; replaces geom_kernels.cu:1301-1308 (ts / ii_exp / jj_exp / _unique),
 * the four per-iteration host CSR builds of accum_cuda (:946-981) and the triple list of schur_block
 * (:1209-1240).  Depends only on (ii, jj, N, ht, wd, t0, t1), so it is built once per call (or cached by
 * the caller across calls on the same graph).
 *
 * ii, jj: HOST int64[n_edges].  rank/world: keyframe sharding; rank 0 of 1 = whole problem.
 */
int vipe_ba_plan_create(const int64_t *ii, const int64_t *jj, int64_t n_edges, int64_t n_frames, int ht, int wd,
                        int t0, int t1, int rank, int world, vipe_ba_plan **out);
/*
 * Batched plan: n_problems INDEPENDENT small problems in one set of launches (e.g. motion-only BA of many video clips,
 * BASELINE config 5).  Frames, edges and tensors are concatenated; problem c owns frames
 * [frame_ptr[c], frame_ptr[c+1]) and optimises the poses of its window [t0s[c], t1s[c]) (global frame ids); edges must
 * stay inside one problem.  Each reduced system is solved by its own CTA, so 6*(t1-t0) <= 128 per problem.
 * dx_out is [sum_c (t1-t0), 6] in problem order (vipe_ba_plan_num_free_poses rows).  The semantics of every problem are
 * those of a separate slam_ext.ba call.
 */
int vipe_ba_plan_create_batch(const int64_t *ii, const int64_t *jj, int64_t n_edges, int64_t n_frames, int ht, int wd,
                              int n_problems, const int64_t *frame_ptr, const int64_t *t0s, const int64_t *t1s,
                              vipe_ba_plan **out);
int64_t vipe_ba_plan_num_free_poses(const vipe_ba_plan *plan);
void vipe_ba_plan_destroy(vipe_ba_plan *plan);

/* K = |kx|, kx = sorted unique of cat(arange(t0,t1), ii)  (geom_kernels.cu:1305-1308). */
int64_t vipe_ba_plan_num_kx(const vipe_ba_plan *plan);
/* HOST int64[K] <- kx;  HOST int64[(t1-t0)+E] <- kk_exp (inverse indices of the unique). */
int vipe_ba_plan_copy_kx(const vipe_ba_plan *plan, int64_t *kx_out);
int vipe_ba_plan_copy_kk_exp(const vipe_ba_plan *plan, int64_t *kk_out);
/* Per-source-frame CSR (the ptrs/idxs of accum_cuda(.., ii, kx)): HOST int64[K+1], HOST int64[E]. */
int vipe_ba_plan_copy_csr(const vipe_ba_plan *plan, int64_t *ptrs_out, int64_t *idxs_out);
/* Owned kx positions [lo, hi) of this rank, and of any rank. */
int vipe_ba_plan_owned_range(const vipe_ba_plan *plan, int rank, int64_t *lo, int64_t *hi);
/* Number of (pose_a, pose_b, frame) triples schur_block would enumerate (:1225-1240); bookkeeping check. */
/* Position of every free pose in the reduced camera system, out[t1 - t0] in pose order: unknowns 6*out[p] .. 6*out[p]+5 of
 * vipe_ba_system_buffer() belong to pose t0 + p.  Today this is the natural order (a fill-reducing pose order does not
 * survive the solver's 64 x 64 tiling); callers that read the system should go through it anyway.
 * dx is always returned in pose order. */
int vipe_ba_plan_copy_sys_order(const vipe_ba_plan *plan, int64_t *out);
int64_t vipe_ba_plan_num_schur_triples(const vipe_ba_plan *plan);
/* Largest number of edges leaving one owned source frame. */
int vipe_ba_plan_max_degree(const vipe_ba_plan *plan);

/* Bytes of DEVICE workspace one run needs (index tables + per-tile partials + reduced camera system). */
size_t vipe_ba_workspace_bytes(const vipe_ba_plan *plan);
/* Copy the plan's index tables into the head of the workspace (async on `stream`). Call once per workspace. */
int vipe_ba_plan_upload(const vipe_ba_plan *plan, void *workspace /*DEV*/, void *stream);

typedef struct vipe_ba_tensors {
    float *poses;             /* DEV [N,7]      in/out */
    float *disps;             /* DEV [N,ht,wd]  in/out */
    const float *intrinsics;  /* DEV [4] */
    const float *disps_sens;  /* DEV [N,ht,wd] */
    const float *targets;     /* DEV [E,2,ht,wd] */
    const float *weights;     /* DEV [E,2,ht,wd] */
    const float *eta;         /* DEV [K,ht,wd]; may be NULL when motion_only */
    float *dx_out;            /* DEV [t1-t0,6]  update of the last iteration */
    float *dz_out;            /* DEV [K,ht*wd]  update of the last iteration; may be NULL when motion_only */
} vipe_ba_tensors;

/*
 * The whole operator: `iterations` Gauss-Newton steps, no host synchronisation, all work on `stream`.
 * Single-rank plans only (world == 1).  Replaces ba_cuda, geom_kernels.cu:1283-1404.
 */
int vipe_ba_run(const vipe_ba_plan *plan, const vipe_ba_tensors *t, void *workspace, int iterations, float lm,
                float ep, int motion_only, void *stream);

/*
 * Semantic switches.  The defaults reproduce the reference's CUDA BA exactly as shipped
 * (csrc/slam_ext/geom_kernels.cu); the other settings express the conventions of the Python BA that vipe/slam
 * actually runs (GraphBuffer.bundle_adjustment, vipe/slam/components/buffer.py:373-525; SURVEY.md section 8(a')) so that
 * an adapter can serve those callers with this operator.  vipe_ba_options_default fills the defaults.
 */
typedef struct vipe_ba_options {
    float min_depth;     /* 0.25 (geom_kernels.cu:33);  Python path: 0.1 (vipe/utils/cameras.py:48) */
    int depth_strict;    /* 0: invalid iff z < min_depth (:301);  1: valid iff z > min_depth (vipe/slam/maths/geom.py:263) */
    float alpha;         /* 0.05 (:1359);  Python path: ba.dense_disp_alpha */
    int sensor_mode;     /* 0: per-pixel mask disps_sens > 0, C += m*alpha + (1-m)*eta (:1361-1369)
                            1: per-frame gate (frame_flags bit 0), prior on every pixel, damping everywhere
                               (vipe/slam/ba/terms.py:244-300, buffer.py:470-489) */
    float eta_scale;     /* disparity damping = eta_scale*eta + eta_bias;  1, 0  (Python path: 0.2, 2e-7) */
    float eta_bias;
    float dz_max;        /* dz > dz_max -> 0;  +inf  (Python path: 10, vipe/slam/maths/retractor.py:41) */
    int renorm_quat;     /* 0 (Q6);  1: renormalise after the pose retraction like lietorch (so3.h:36-38) */
    int damp_on_pose_hessian; /* 0: lm scales diag(A - S) (:1176);  1: lm scales diag(A) before the Schur complement (solver.py:161-164) */
    int backsub_all_poses;    /* 0: pose index 0 never reaches dz (:1089, Q4);  1: every free pose does (solver.py:182) */
    const unsigned char *frame_flags; /* DEV [K] or NULL.  bit 0: sensor gate (mode 1); bit 1: disparity of this kx frame fixed */
    /* The focal length as one more variable, shared by fx and fy: the pinhole case of the backend's optimize_intrinsics
       (vipe/slam/components/buffer.py:496-498, vipe/slam/maths/retractor.py:51-62, vipe/slam/ba/terms.py:217-228).
       Needs damp_on_pose_hessian and a single-problem, single-rank plan.  While it is on,
         - vipe_ba_tensors.intrinsics is UPDATED IN PLACE (fx, fy += focal_jscale * step) every iteration, and
         - dx_out must hold 6*(t1-t0) + 1 floats: the focal step of the last iteration follows the pose steps. */
    int optimize_focal;  /* 0 */
    float focal_jscale;  /* d(intrinsics handed to the kernels)/d(variable): 1/intrinsics_factor = 1/8 in the Python path (terms.py:186) */
    float focal_lm;      /* LHS_ff += focal_lm * H_ff + focal_ep;  Python path: 1e-6, 1e-6 (buffer.py:496) */
    float focal_ep;
} vipe_ba_options;
void vipe_ba_options_default(vipe_ba_options *opt);
int vipe_ba_set_options(vipe_ba_plan *plan, const vipe_ba_options *opt);

/* CUDA-graph replay of repeated vipe_ba_run calls with identical arguments (default on). */
int vipe_ba_set_graphs(vipe_ba_plan *plan, int on);

/*
 * The same iteration split at its one exchange point, for keyframe-sharded multi-GPU runs:
 *   vipe_ba_linearize   : stages 1-3 on this rank's source frames -> partial reduced camera system
 *   (caller all-reduces the buffer returned by vipe_ba_system_buffer over NCCL)
 *   vipe_ba_solve_update: damped Cholesky solve (replicated), back-substitution + retraction of owned frames.
 */
int vipe_ba_linearize(const vipe_ba_plan *plan, const vipe_ba_tensors *t, void *workspace, int motion_only,
                      void *stream);
int vipe_ba_solve_update(const vipe_ba_plan *plan, const vipe_ba_tensors *t, void *workspace, float lm, float ep,
                         int motion_only, void *stream);
/*
 * Multi-GPU without a separate all-reduce (needs NVSwitch multicast memory, e.g. torch symmetric memory):
 * `accum_local` is this rank's instance of a symmetric buffer of `count_out` doubles (vipe_ba_system_buffer) and
 * `accum_multicast` the multicast address of the same buffer.  While set, vipe_ba_linearize accumulates this rank's
 * partial system into `accum_local` instead of the workspace, and vipe_ba_solve_update reads its input THROUGH THE
 * SWITCH (multimem.ld_reduce: the sum over all ranks' instances) tile by tile as the factorisation reaches it -- the
 * all-reduce is fused into the Cholesky's loads.  The caller provides the cross-rank barrier between the two calls and
 * alternates two buffers from iteration to iteration (a rank may start clearing buffer k+2 only after every rank has
 * finished reading buffer k, which the barrier of iteration k+1 guarantees).  Needs 6*(t1-t0) > 128.  NULL, NULL: off.
 */
int vipe_ba_set_peer_system(vipe_ba_plan *plan, double *accum_local, const double *accum_multicast);
/*
 * The same reduction in two steps (better from 4 GPUs on: the factorisation then reads local memory at full speed):
 * with vipe_ba_set_peer_system(plan, accum_local, accum_multicast) still routing the accumulation, call
 * vipe_ba_peer_reduce between two cross-rank barriers -- every rank sums 1/world of the tiles over all instances of `accum`
 * (multimem.ld_reduce) and multicasts them into all instances of `reduced` (multimem.st) -- and point the solve at this
 * rank's instance of `reduced` with vipe_ba_set_solve_buffer (the solve factorises it in place; NULL: back to the
 * workspace / the fused loads).  One accumulation buffer suffices in this mode.
 */
int vipe_ba_peer_reduce(const vipe_ba_plan *plan, const double *accum_multicast, double *reduced_multicast, int rank, int world,
                        void *stream);
int vipe_ba_set_solve_buffer(vipe_ba_plan *plan, double *reduced_local);
/*
 * Distributed factorisation (one process per GPU, NVSwitch multicast memory): instead of every rank factorising the whole
 * reduced system, tile column j of the 64 x 64 tile grid belongs to rank j % world.  An owner reads its input tiles summed
 * over the ranks' partial systems (the multicast address given to vipe_ba_set_peer_system -- required), factorises, and
 * writes L, y = L^-1 b and the inverted diagonal tiles through the MULTICAST aliases, so they land in every rank's instance
 * together with the tile's ready flag; consumers wait on their local flags and read local memory.  The backward
 * substitution is replicated, so every rank ends with the same dx.  `factor_*`: symmetric buffer of n*n + 2n doubles
 * (vipe_ba_system_buffer's count), `aux_*`: symmetric buffer of vipe_ba_dist_aux_bytes(plan) bytes, zero-filled once when it
 * is created and never touched by the caller afterwards (its ready flags carry a per-solve epoch instead of being reset).
 * One cross-rank barrier per iteration, between vipe_ba_linearize and vipe_ba_solve_update, as with the fused loads; the
 * accumulation buffers alternate in the same way.  Replaces the reference's single-GPU SimplicialLLT solve
 * (csrc/slam_ext/geom_kernels.cu:1172-1191).  NULL factor_local: off.
 */
int64_t vipe_ba_dist_aux_bytes(const vipe_ba_plan *plan);
int vipe_ba_set_dist_solve(vipe_ba_plan *plan, double *factor_local, double *factor_multicast, void *aux_local,
                           void *aux_multicast, int rank, int world);
/*
 * Owner-only pixel inputs for sharded plans (a rank HOLDS the targets/weights of its edges): when
 * on, `targets` / `weights` in vipe_ba_tensors are [n_owned, 2, ht, wd], the rows of this rank's edges in the order of
 * vipe_ba_plan_copy_owned_edges (the plan's CSR order over the owned source frames).  Everything else stays replicated.
 */
int vipe_ba_set_owned_rows(vipe_ba_plan *plan, int on);
int64_t vipe_ba_plan_num_owned_edges(const vipe_ba_plan *plan);
int vipe_ba_plan_copy_owned_edges(const vipe_ba_plan *plan, int64_t *edge_ids_out);

/* DEV fp64 buffer holding [H (n x n, row-major, lower triangle valid) ; b (n) ; diag of the pose Hessian (n)],
   n = 6*(t1-t0) rounded up to 64; count_out = n*n + 2n.  This is what a sharded run all-reduces. */
void *vipe_ba_system_buffer(const vipe_ba_plan *plan, void *workspace, int64_t *n_out, int64_t *count_out);

/* Test hooks: DEV fp32 [K,HW] buffers written by the last linearize (Q = 1/C and Q*w of geom_kernels.cu:1365-1370). */
float *vipe_ba_debug_q(const vipe_ba_plan *plan, void *workspace);
float *vipe_ba_debug_qw(const vipe_ba_plan *plan, void *workspace);
/*
 * Optional per-stage timing (tracing hook).  When enabled, vipe_ba_run records CUDA events on the run's stream
 * around each stage of each iteration (linearise+Schur | assemble | solve | back-substitute+retract).  After the
 * caller has synchronised the stream, vipe_ba_profile_read returns the summed milliseconds per stage of the last
 * run and the number of iterations they cover.
 */
int vipe_ba_profile_enable(vipe_ba_plan *plan, int on);
int vipe_ba_profile_read(const vipe_ba_plan *plan, float ms_out[4], int *iterations_out);

/* Number of kernels the last vipe_ba_run / linearize+solve_update pair enqueued (for bench.py's gpu_launches). */
int64_t vipe_ba_launch_count(const vipe_ba_plan *plan);

/*
 * The other operators of the reference's slam_ext module (csrc/slam_ext/slam.cpp:33-36), same layouts and semantics
 * as their *_cuda hosts (geom_kernels.cu:1406-1507).  Index arrays are DEVICE int64; all work is enqueued on `stream`.
 */
/* projmap_cuda (:1436-1460): coords DEV [E,ht,wd,3] (third channel 0), valid DEV [E,ht,wd,1]. */
int vipe_projmap(const float *poses, const float *disps, const float *intrinsics /*[4]*/, const int64_t *ii,
                 const int64_t *jj, int64_t n_edges, int ht, int wd, float *coords, float *valid, void *stream);
/* frame_distance_cuda (:1406-1434): intrinsics DEV [Q,4]; pi,pj pose ids, qi,qj intrinsics ids, di disparity ids; dist DEV [M]. */
int vipe_frame_distance(const float *poses, const float *disps, const float *intrinsics, const int64_t *pi,
                        const int64_t *pj, const int64_t *qi, const int64_t *qj, const int64_t *di, int64_t n_pairs, int ht,
                        int wd, float beta, float *dist, void *stream);
/* depth_filter_cuda (:1462-1486): counter DEV [n_ix,ht,wd]; neighbours ix-3..ix+3 inside [0, n_frames). */
int vipe_depth_filter(const float *poses, const float *disps, const float *intrinsics /*[4]*/, const int64_t *ix,
                      const float *thresh, int64_t n_ix, int64_t n_frames, int ht, int wd, float *counter, void *stream);
/* iproj_cuda (:1488-1507): points DEV [N,ht,wd,3]. */
int vipe_iproj(const float *poses, const float *disps, const float *intrinsics /*[4]*/, int64_t n_frames, int ht, int wd,
               float *points, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VIPE_BA_H_ */
