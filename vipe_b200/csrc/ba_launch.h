// Host-side launcher declarations shared between the kernel translation units and the C ABI.
#pragma once

#include <cuda_runtime.h>

#include "ba_common.cuh"

namespace vba {

// choose the (threads, pixels-per-thread) tile shape of the frame kernels for a given max degree
bool tile_config(int HW, int dmax, bool motion, int &NT, int &PPT);

cudaError_t launch_linearize(const LinArgs &a, int nframes, int dmax, bool motion, int NT, int PPT, cudaStream_t st);
// packed 2-pixels-per-thread variant (ba_linearize2.cu); needs an even HW.  TILE = 2 * NT.
bool tile_config2(int HW, int dmax, bool motion, int &NT);
cudaError_t launch_linearize2(const LinArgs &a, int nframes, int dmax, bool motion, int NT, cudaStream_t st, bool ug = false);
// shared-memory bytes of the packed kernel's staging buffer for out-degree d, and what a CTA may use
size_t lin2_staging_bytes(int d, int NT);
size_t lin2_staging_cap();
// TMA-fed motion-only pipeline (ba_lin3.cu); `econst_dev` is E x 16 floats of scratch for the per-edge constants.
bool lin3_supported(int HW, int chunk_px);
// motion-only variant: flat (edge slot, chunk) units over the slots [slot_lo, slot_lo + nslots); `slot_src_dev[slot]` = source frame id
cudaError_t launch_lin3_motion(const LinArgs &a, const int *slot_src_dev, int slot_lo, int nslots, int chunk_px, float *econst_dev,
                               int num_sms, cudaStream_t st);
// Second-generation pipeline (ba_lin4.cu): cw_kernel (disparity blocks -> Q) + lin4_kernel (TMA-fed J warps, tcgen05 Gram) for
// frames with 1 <= out-degree <= kLin4MaxDeg.  Items = (frame, 256-pixel chunk), units = (item, 64-pixel sub-tile, edge) in that
// order; CTA b of the persistent grid owns the contiguous item range [cta_item[b], cta_item[b + 1]) and its units
// [cta_unit[b], cta_unit[b + 1]).  epart holds one record per unit (Tables::ntile_e = HW / 64).
constexpr int kLin4MaxDeg = 20, kLin4GroupDeg = 10, kLin4NJ = 10, kLin4ItemPx = 256;
struct Lin4Item {
    int k, d, chunk, pad;  // kx position, out-degree, chunk index
};
struct Lin4Unit {
    int slot, edge;  // CSR slot and edge id
    int srcpx, kpx;  // src * HW + px0, k * HW + px0
    int ed;          // position of the edge in its frame | out-degree of the frame << 8
    int rc;          // image row of px0 << 16 | column of px0
    int px0, g;      // first pixel inside the frame; running sub-tile number (item * 4 + sub-tile) in list order
};
struct Lin4Launch {
    const Lin4Item *items;
    const Lin4Unit *units;
    const int *cta_item, *cta_unit;  // [grid + 1]
    int grid, nframes, dmax;
    const int *slot_src;
    int slot_lo, nslots;
    float2 *econst2;         // [E][16] per-edge constants, each value twice
    float *sqbuf, *sqwbuf;   // [K][HW] sqrt(Q), sqrt(Q) w
};
bool lin4_supported(int HW, int dmax);
cudaError_t launch_lin4(const LinArgs &a, const Lin4Launch &l, cudaStream_t st);
cudaError_t launch_frame_reduce(const ReduceArgs &a, int nframes, int dmax, cudaStream_t st);
cudaError_t launch_assemble(const AssembleArgs &a, cudaStream_t st);
cudaError_t launch_backsub(const BackArgs &a, int nframes, int dmax, cudaStream_t st);
// `intr` non-null: also apply the focal step dx[focal_row] * focal_jscale to fx and fy
cudaError_t launch_pose_retr(float *poses, const float *dx, const int *pose_row, int n_poses, int renorm, float *intr,
                             int focal_row, float focal_jscale, cudaStream_t st);
// focal-length pass (optimize_focal): its own 256-pixel tiling of a frame
int focal_tiles(int HW);
cudaError_t launch_add_scalar(double *p, double v, cudaStream_t st);
cudaError_t launch_focal(const FocalArgs &a, int nframes, int dmax, cudaStream_t st);

// Distributed factorisation over `world` GPUs (tile columns cyclic over the ranks, results multicast through the NVSwitch):
// H / scratch / linvT passed to launch_damped_solve are this rank's instances of symmetric buffers, the fields below their
// multicast aliases; Ain (the multicast address of the ranks' partial systems) is required; `epoch` must be the same on every
// rank and different from the previous solve's.
struct CholDist {
    int rank, world, epoch;
    int colblk;  // tile columns are dealt to the ranks in blocks of this many (neighbouring columns on one rank keep the
                 // single-GPU shortcuts of the column-to-column hand-over; 1 = plain cyclic)
    double *Hmc;
    int *scratch_mc;
    double *linvT_mc;
};
// Dense damped Cholesky solve of the reduced camera system (chol.cu).
//   H: [npad x npad] fp64 row-major, lower triangle valid on entry (destroyed);  b: [npad] fp64 (destroyed)
//   diag += ep + lm*diag (geom_kernels.cu:1176); factor; solve; dx[n] fp32.  Failure => dx = 0 (:1186-1188).
// Returns the number of kernels launched through *launches.
// `scratch` is chol_scratch_ints(npad) ints of device memory: tile ready flags, the words of the backward substitution's x
// exchange and the fast copies of the factorised diagonal tiles (chol.cu).  The solve zeroes all of it itself with one
// stream-ordered memset per call (zero = "not written yet" for the self-validating words); `epoch` is unused.
cudaError_t launch_damped_solve(double *H, double *b, int n, int npad, float lm, float ep, float *dx, int *scratch,
                                double *dinv /*[npad]*/, double *linvT /*[npad/64][4096]*/,
                                const double *dampdiag /*[npad] or null*/,
                                const double *Ain /*null, or multicast address of the ranks' [H;b;diag(A)] partials*/,
                                int epoch, const unsigned char *tstruct /*DEV [T][T] tile structure of L, or null = dense*/,
                                const int *rowmap /*DEV [npad] system unknown -> dx index, or null = identity*/, cudaStream_t st,
                                int *launches, const CholDist *dist = nullptr);
size_t chol_scratch_ints(int npad);
// Two-step in-switch all-reduce of the ranks' partial systems [H (lower tiles) ; b ; diag(A)]: this rank sums every world-th
// tile over all instances (multimem.ld_reduce through `accum_mc`) and multicasts the sums into every rank's instance of
// `reduced_mc` (multimem.st).  Between two cross-rank barriers.
cudaError_t launch_peer_reduce(const double *accum_mc, double *reduced_mc, int npad, int rank, int world, cudaStream_t st);
// zero every problem's [H ; b ; diag(A)] block and put the identity on the padded diagonals
cudaError_t launch_system_clear(double *sys, size_t total_doubles, const long long *prob_hoff, const int *prob_n,
                                const int *prob_npad, int n_prob, bool any_padding, cudaStream_t st);
// batched solve of C > 1 small problems (each at most 2 tiles), one CTA per problem
cudaError_t launch_small_solve_batch(double *sys, const long long *prob_hoff, const int *prob_n, const int *prob_npad,
                                     const int *prob_row0, int n_prob, float lm, float ep, float *dx, bool damp_on_A,
                                     cudaStream_t st, int *launches);

constexpr int kCholBlock = 64;

// records msg as the calling thread's last error (vipe_ba_last_error) and returns 1
int set_last_error(const char *msg);

}  // namespace vba
