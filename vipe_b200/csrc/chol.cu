// Damped dense Cholesky solve of the reduced camera system, entirely on the device, fp64.
//
// Replaces SparseBlock::solve (csrc/slam_ext/geom_kernels.cu:1172-1191): the reference copies the system to
// the host and runs Eigen::SimplicialLLT<double>; here the factorisation never leaves the GPU.
//
// Tile dataflow (64x64 tiles, left-looking): persistent CTAs claim tiles in column-major order from an atomic
// counter; tile (i,j) subtracts L_ik L_jk^T for k < j as soon as those tiles are published (per-tile ready flags,
// release/acquire at gpu scope), then is factorised (i == j) or solved against L_jj (i > j) and published.
// The right-hand side rides along as one extra tile row (forward substitution for free); a second dataflow
// kernel does the backward substitution.  Claim order respects the dependencies, so every spin-wait is on a tile
// that a resident CTA is already working on: no cooperative launch is needed.
// At backend sizes the solve is bound by the latency of the chain of diagonal tiles (potrf -> tile solve -> update ->
// potrf ...), so the hops on that chain avoid flags and fences: the factorised diagonal tile and the backward
// substitution's x travel as self-validating 8-byte words (zero = not written yet), the diagonal tile in two halves so
// that its readers start their tile solve while the second half is still being factorised (see "self-validating
// exchange" below).  Everything not on that chain uses the ordinary ready flags.
// A non-positive (or NaN) pivot raises *fail and the solve writes dx = 0, like the reference (:1186-1188).
#include <type_traits>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ba_launch.h"

namespace vba {

constexpr int TB = kCholBlock;  // 64
constexpr int LD = TB + 2;      // 66: keeps 16-byte alignment of 4-double groups in the [k][row] layout
constexpr int CT = 256;         // threads per CTA
constexpr int kMaxSparseT = 1024;  // largest tile count the per-tile step list (shared memory) holds; beyond it the solve is dense

// ------------------------------------------------------------------------------------------------
__global__ void pad_identity_kernel(double *sys, const long long *prob_hoff, const int *prob_n, const int *prob_npad) {
    const int c = blockIdx.x;
    double *H = sys + prob_hoff[c];
    const int n = prob_n[c], npad = prob_npad[c];
    for (int i = n + threadIdx.x; i < npad; i += blockDim.x) H[(size_t)i * npad + i] = 1.0;
}

cudaError_t launch_system_clear(double *sys, size_t total_doubles, const long long *prob_hoff, const int *prob_n,
                                const int *prob_npad, int n_prob, bool any_padding, cudaStream_t st) {
    cudaError_t err = cudaMemsetAsync(sys, 0, total_doubles * sizeof(double), st);
    if (err != cudaSuccess) return err;
    if (any_padding) {
        pad_identity_kernel<<<n_prob, 64, 0, st>>>(sys, prob_hoff, prob_n, prob_npad);
        return cudaGetLastError();
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int *p) {  // (a hint in the distributed solve: relaxed is enough)
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// `dist`: the flag may have been set by another GPU (distributed factorisation): system-scope acquire
__device__ __forceinline__ bool flag_is(const int *f, int epoch, bool dist) { return (dist ? ld_acquire_sys(f) : ld_acquire(f)) == epoch; }
// all threads call; returns after flag == epoch is visible to the whole CTA
__device__ __forceinline__ void wait_flag(const int *f, int epoch, bool dist = false) {
    if (threadIdx.x == 0) {
        while (!flag_is(f, epoch, dist)) __nanosleep(32);
    }
    __syncthreads();
}
// ---- distributed factorisation: results go to every rank's instance through the NVSwitch multicast address --------------
__device__ __forceinline__ bool is_hole(double x) { return __double_as_longlong(x) == 0ll; }
__device__ __forceinline__ void mc_store(double *p, double v) {
    if (is_hole(v)) v = -0.0;
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// Everything an owner multicasts is SELF-VALIDATING: the receiving buffers are zeroed before the solve, every 8-byte word that
// travels has a non-zero bit pattern (+0.0 goes as -0.0), and a reader re-fetches until it sees no zero word.  The ready flags
// are then mere hints, stored relaxed right behind the data: no fence, no release.  (A system-scope release behind 32 KB of
// multicast stores was measured at ~10 us per tile -- on the critical path of every tile column.)
// two adjacent doubles in one 16-byte multicast store (multimem.st has no f64 vector form; the bits travel as 4 x f32)
__device__ __forceinline__ void mc_store2(double *p, double x, double y) {
    unsigned long long bx = (unsigned long long)__double_as_longlong(x), by = (unsigned long long)__double_as_longlong(y);
    if (bx == 0ull) bx = 0x8000000000000000ull;
    if (by == 0ull) by = 0x8000000000000000ull;
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float((unsigned)bx)),
                 "f"(__uint_as_float((unsigned)(bx >> 32))), "f"(__uint_as_float((unsigned)by)), "f"(__uint_as_float((unsigned)(by >> 32)))
                 : "memory");
}
// all threads call after their multicast writes: the flag lands in every rank's flag array behind the data
__device__ __forceinline__ void publish_flag_mc(int *f_mc, int epoch) {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("multimem.st.relaxed.sys.global.b32 [%0], %1;" ::"l"(f_mc), "r"(epoch) : "memory");
}
// all threads call after their global writes
__device__ __forceinline__ void publish_flag(int *f, int epoch) {
    __syncthreads();
    // st.release.gpu is cumulative over everything ordered before it by the barrier; an extra __threadfence() here
    // compiled to a second (sequentially consistent) MEMBAR on the critical path of every tile
    if (threadIdx.x == 0) st_release(f, epoch);
}

#ifdef VBA_CHOL_TRACE
__device__ long long *g_trace = nullptr;  // [tile][8] globaltimer stamps (developer builds only)
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(tile, slot)                                                          \
    do {                                                                           \
        if (threadIdx.x == 0 && g_trace) g_trace[(size_t)(tile) * 8 + (slot)] = gtime(); \
    } while (0)
#else
#define TRACE(tile, slot) \
    do {                  \
    } while (0)
#endif

struct CholArgs {
    double *H;   // [npad x npad] row-major, lower triangle valid; overwritten by L
    double *b;   // [npad]; overwritten by y = L^-1 b, then by x
    int ld, T, n;
    float lm, ep;
    int *flags;    // [(T+1) x T] tile ready flags, value == epoch when ready; row T is the rhs row
    unsigned long long *xs;  // [npad] x of the backward substitution as self-validating words (0 = not there yet)
    unsigned long long *ldiag;  // [T][64*64 + 64] factorised diagonal tiles + 1/diag as self-validating words (fast path)
    int *preflags; // [T] tile (j+1, j) is available BEFORE its solve against L_jj, parked in the unused upper tile (j, j+1)
    int *counter;  // [2] tile counters (factorisation, backward), zero on entry
    int *fail;     // zero on entry
    int epoch;
    float *dx;
    double *dinv;  // [npad] 1 / diag(L)
    const double *dampdiag;  // [npad] or null: lm scales this instead of the matrix's own diagonal
    double *linvT;           // [T][64*64] inverse-transposes of the diagonal tiles (for the backward substitution)
    // Multi-GPU, optional: multicast address of the ranks' partial systems [H ; b ; diag(A)] (same layout, ld*ld + 2*ld
    // doubles).  When set, the INPUT of the factorisation is read with multimem.ld_reduce, i.e. summed over the ranks
    // inside the NVSwitch, tile by tile, as the dataflow reaches it -- the all-reduce of the reduced camera system is
    // fused into the Cholesky's own loads.  Outputs (L, y, x) still go to the local H / b.
    const double *Ain;
    // Sparsity, optional: ts[i * T + j] != 0 when tile (i, j) of L can be non-zero (tile-level symbolic factorisation done by
    // the plan on a fill-reducing pose order); zero tiles are never claimed, read or waited for.  rowmap: system unknown ->
    // index in dx (undoes that order).
    const unsigned char *ts;
    const int *rowmap;
    // Distributed factorisation (world > 1): tile column j (and its rhs tile) belongs to rank j % world.  Every rank holds a
    // full copy of L, y, L_jj^-T and the ready flags in symmetric memory; an owner writes its results through the multicast
    // aliases below, so they land in every rank's copy, and consumers wait on their LOCAL flags and read LOCAL memory.  The
    // input is summed over the ranks' partial systems by the owner's loads (Ain).  Flags hold the solve's epoch (no reset
    // between solves: a reset could race with a faster rank's publications).  The backward substitution is replicated.
    int rank = 0, world = 1, colblk = 1;  // columns are dealt in blocks of `colblk`: owner(j) = (j / colblk) % world
    double *Hmc = nullptr, *bmc = nullptr, *linvT_mc = nullptr;
    int *flags_mc = nullptr, *fail_mc = nullptr;
    int *lflags = nullptr, *lflags_mc = nullptr;  // [T] L_jj^-T of column j has arrived (the backward substitution reads it)
};
__device__ __forceinline__ int col_owner(const CholArgs &a, int j) { return (j / a.colblk) % a.world; }

// one element of the rank-summed input system through the switch (NVLS in-switch reduction)
__device__ __forceinline__ double mc_load_sum(const double *p) {
    double v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Tile GEMM on the fp64 tensor cores (mma.sync.m8n8k4.f64 -> DMMA).  Measured on B200 (scripts/dmma_lat.cu): one DMMA
// (256 FMA) issues every 16 clk per SM sub-partition, i.e. 64 FMA/clk/SM, a dependent DMMA follows after 32-42 clk; a
// DFMA outer-product loop fed from shared memory sustains only 26-28 FMA/clk/SM (operand delivery), so the trailing
// updates of the factorisation use DMMA.  A 64x64x64 tile product is 4096 clk of tensor pipe per SM.
//   acc -= A * B^T, A = L_ik and B = L_jk as ROW-MAJOR tiles in shared memory with row stride RS = 68 doubles
//   (68 = 4 mod 16 makes both fragment loads conflict-free: lane l reads [l/4][k0 + l%4]).
// A warp owns output column block cw (columns 8cw..8cw+7) and the row blocks mt >= mt0: 8 accumulator fragments; the
// C fragment of row block mt holds rows 8mt + l/4, columns 8cw + 2(l%4) + {0,1}.
//   off-diagonal tiles: cw = warp, mt0 = 0.
//   diagonal tiles (symmetric, only the lower blocks are ever read): cw = diag_col(warp), mt0 = cw.  Column c has 8 - c
//   lower blocks; the tensor pipe belongs to the sub-partition, which hosts warps s and s+4, so giving those two
//   warps columns s and 7-s loads every pipe with 9 blocks instead of 16: 0.56 of the time of the full update.
constexpr int RS = 68;
__device__ __forceinline__ int diag_col(int warp) { return warp < 4 ? warp : 11 - warp; }
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
template <int MT0>
__device__ __forceinline__ void tile_gemm_sub_from(double (&acc)[8][2], const double *ap, const double *bp) {
#pragma unroll 4
    for (int k0 = 0; k0 < TB; k0 += 4) {
        const double bneg = -bp[k0];
#pragma unroll
        for (int mt = MT0; mt < 8; mt++) dmma884(acc[mt], ap[mt * 8 * RS + k0], bneg);
    }
}
__device__ __forceinline__ void tile_gemm_sub(double (&acc)[8][2], const double *As, const double *Bs, int cw, int mt0) {
    const int lane = threadIdx.x & 31;
    const double *ap = As + (lane >> 2) * RS + (lane & 3);
    const double *bp = Bs + (8 * cw + (lane >> 2)) * RS + (lane & 3);
    // a branch per first row block, not a predicate per DMMA: a predicated-off DMMA still holds the tensor pipe (measured:
    // the predicated form of the lower-blocks-only update took exactly as long as the full one)
    switch (mt0) {
        case 0: tile_gemm_sub_from<0>(acc, ap, bp); break;
        case 1: tile_gemm_sub_from<1>(acc, ap, bp); break;
        case 2: tile_gemm_sub_from<2>(acc, ap, bp); break;
        case 3: tile_gemm_sub_from<3>(acc, ap, bp); break;
        case 4: tile_gemm_sub_from<4>(acc, ap, bp); break;
        case 5: tile_gemm_sub_from<5>(acc, ap, bp); break;
        case 6: tile_gemm_sub_from<6>(acc, ap, bp); break;
        default: tile_gemm_sub_from<7>(acc, ap, bp); break;
    }
}
__device__ __forceinline__ void tile_gemm_sub(double (&acc)[8][2], const double *As, const double *Bs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    tile_gemm_sub_from<0>(acc, As + (lane >> 2) * RS + (lane & 3), Bs + (8 * warp + (lane >> 2)) * RS + (lane & 3));
}

// ---- self-validating exchange of the factorised diagonal tiles ------------------------------------------------
// The hop "diagonal tile factorised -> its consumers start" is on the critical path twice per column.  Through a
// ready flag it costs a barrier, a release fence and the flag store on the writer's side, then a flag poll and only
// then the tile load on the reader's side.  Here the tile itself is the message: the buffer is zeroed before the
// solve, the writer stores every 8-byte word with a non-zero bit pattern (+0.0 travels as -0.0), and every reader
// thread polls exactly the words it needs until all of them are non-zero -- no fence, no flag, one trip to L2 after
// the data lands.  (H and the regular flag are still written afterwards for the readers that are not in a hurry.)
constexpr int kFastTile = TB * TB + TB;  // words per diagonal tile: L (row-major 64x64, zero above the diagonal) + 1/diag
__device__ __forceinline__ unsigned long long nz_bits(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return b == 0ull ? 0x8000000000000000ull : b;
}
// The tile travels in two halves: columns 0..31 (L11, L21, 1/diag[0:32]) are final before the second 32x32 block is
// factorised, so the otherwise idle warps 1..7 send them while warp 0 walks the second half of the pivot chain, and
// the readers run the first four column blocks of their tile solve under it.  Half B = L22 and 1/diag[32:64]; the
// upper right block is never read.
//   `t` / `nt`: index and number of the calling threads (half A is stored by 224 threads, half B by all 256)
__device__ __forceinline__ void fast_half_store(unsigned long long *ft, const double *S, const double *dinv, int half, int t, int nt) {
    const int row0 = half ? 32 : 0, col0 = half ? 32 : 0, nrows = half ? 32 : 64;
    for (int idx = t; idx < nrows * 16; idx += nt) {
        const int r = row0 + (idx >> 4), c = col0 + (idx & 15) * 2;
        const double2 v = *reinterpret_cast<const double2 *>(S + r * RS + c);
        const unsigned long long b0 = nz_bits(c <= r ? v.x : 0.0), b1 = nz_bits(c + 1 <= r ? v.y : 0.0);
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(ft + r * TB + c), "l"(b0), "l"(b1) : "memory");
    }
    if (t < 32)
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(ft + TB * TB + col0 + t), "l"(nz_bits(dinv[col0 + t])) : "memory");
}
// all threads call; returns with the half in S (row stride RS) and its 1/diag in dinv (caller adds the barrier)
template <int HALF>
__device__ __forceinline__ void fast_half_load(double *S, double *dinv, const unsigned long long *ft) {
    constexpr int row0 = HALF ? 32 : 0, col0 = HALF ? 32 : 0, nrows = HALF ? 32 : 64;
    constexpr int U = nrows * 16 / CT;
    unsigned long long lo[U], hi[U], dv = 1ull;
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int idx = threadIdx.x + u * CT;
            const int r = row0 + (idx >> 4), c = col0 + (idx & 15) * 2;
            asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo[u]), "=l"(hi[u]) : "l"(ft + r * TB + c) : "memory");
        }
        if (threadIdx.x < 32)
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(dv) : "l"(ft + TB * TB + col0 + threadIdx.x) : "memory");
#pragma unroll
        for (int u = 0; u < U; u++) ok = ok && lo[u] != 0ull && hi[u] != 0ull;
        ok = ok && dv != 0ull;
        if (ok) break;
        __nanosleep(20);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int idx = threadIdx.x + u * CT;
        const int r = row0 + (idx >> 4), c = col0 + (idx & 15) * 2;
        *reinterpret_cast<double2 *>(S + r * RS + c) = make_double2(__longlong_as_double((long long)lo[u]), __longlong_as_double((long long)hi[u]));
    }
    if (threadIdx.x < 32) dinv[col0 + threadIdx.x] = __longlong_as_double((long long)dv);
}

// load a 64x64 tile (global row-major, leading dim ld) into shared memory row-major with stride RS
__device__ __forceinline__ void load_tile_R(double *S, const double *__restrict__ G, int ld) {
    double2 v[TB * TB / 2 / CT];
#pragma unroll
    for (int u = 0; u < TB * TB / 2 / CT; u++) {
        const int idx = threadIdx.x + u * CT;
        const int r = idx >> 5, k2 = (idx & 31) * 2;
        v[u] = __ldcg(reinterpret_cast<const double2 *>(G + (size_t)r * ld + k2));
    }
#pragma unroll
    for (int u = 0; u < TB * TB / 2 / CT; u++) {
        const int idx = threadIdx.x + u * CT;
        const int r = idx >> 5, k2 = (idx & 31) * 2;
        *reinterpret_cast<double2 *>(S + r * RS + k2) = v[u];
    }
}

// asynchronous version of load_tile_R (LDGSTS, L2-only caching: the tiles are written by other SMs): the copy of the
// next k-step's tiles runs under the current step's tensor-core update
__device__ __forceinline__ void tile_cp_async(double *S, const double *__restrict__ G, int ld) {
#pragma unroll
    for (int u = 0; u < TB * TB / 2 / CT; u++) {
        const int idx = threadIdx.x + u * CT;
        const int r = idx >> 5, k2 = (idx & 31) * 2;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(S + r * RS + k2);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(G + (size_t)r * ld + k2) : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// load a 64x64 tile (global row-major, leading dim ld) into shared memory transposed: S[k][row].
// All 8 loads of a thread are issued before the first store so that one L2 round trip covers the tile.
__device__ __forceinline__ void load_tile_T(double *S, const double *__restrict__ G, int ld) {
    double2 v[TB * TB / 2 / CT];
#pragma unroll
    for (int u = 0; u < TB * TB / 2 / CT; u++) {
        const int idx = threadIdx.x + u * CT;
        const int r = idx >> 5, k2 = (idx & 31) * 2;
        v[u] = __ldcg(reinterpret_cast<const double2 *>(G + (size_t)r * ld + k2));
    }
#pragma unroll
    for (int u = 0; u < TB * TB / 2 / CT; u++) {
        const int idx = threadIdx.x + u * CT;
        const int r = idx >> 5, k2 = (idx & 31) * 2;
        S[k2 * LD + r] = v[u].x;
        S[(k2 + 1) * LD + r] = v[u].y;
    }
}

// Self-validating variants for the distributed solve (see mc_store2): every thread re-fetches its own 16-byte pieces until
// they hold no zero word.  fix_holes repairs what tile_cp_async left in shared memory (same thread -> element mapping).
__device__ __forceinline__ void load_tile_R_valid(double *S, const double *__restrict__ G, int ld) {
    double2 v[TB * TB / 2 / CT];
    for (;;) {
        bool bad = false;
#pragma unroll
        for (int u = 0; u < TB * TB / 2 / CT; u++) {
            const int idx = threadIdx.x + u * CT;
            v[u] = __ldcg(reinterpret_cast<const double2 *>(G + (size_t)(idx >> 5) * ld + (idx & 31) * 2));
        }
#pragma unroll
        for (int u = 0; u < TB * TB / 2 / CT; u++) bad |= is_hole(v[u].x) | is_hole(v[u].y);
        if (!bad) break;
        __nanosleep(100);
    }
#pragma unroll
    for (int u = 0; u < TB * TB / 2 / CT; u++) {
        const int idx = threadIdx.x + u * CT;
        *reinterpret_cast<double2 *>(S + (idx >> 5) * RS + (idx & 31) * 2) = v[u];
    }
}
__device__ __forceinline__ void fix_holes(double *S, const double *__restrict__ G, int ld) {
#pragma unroll
    for (int u = 0; u < TB * TB / 2 / CT; u++) {
        const int idx = threadIdx.x + u * CT;
        const int r = idx >> 5, k2 = (idx & 31) * 2;
        double2 v = *reinterpret_cast<const double2 *>(S + r * RS + k2);
        if (is_hole(v.x) | is_hole(v.y)) {
            do {
                __nanosleep(100);
                v = __ldcg(reinterpret_cast<const double2 *>(G + (size_t)r * ld + k2));
            } while (is_hole(v.x) | is_hole(v.y));
            *reinterpret_cast<double2 *>(S + r * RS + k2) = v;
        }
    }
}
__device__ __forceinline__ void load_tile_T_valid(double *S, const double *__restrict__ G, int ld) {
    double2 v[TB * TB / 2 / CT];
    for (;;) {
        bool bad = false;
#pragma unroll
        for (int u = 0; u < TB * TB / 2 / CT; u++) {
            const int idx = threadIdx.x + u * CT;
            v[u] = __ldcg(reinterpret_cast<const double2 *>(G + (size_t)(idx >> 5) * ld + (idx & 31) * 2));
        }
#pragma unroll
        for (int u = 0; u < TB * TB / 2 / CT; u++) bad |= is_hole(v[u].x) | is_hole(v[u].y);
        if (!bad) break;
        __nanosleep(100);
    }
#pragma unroll
    for (int u = 0; u < TB * TB / 2 / CT; u++) {
        const int idx = threadIdx.x + u * CT;
        const int r = idx >> 5, k2 = (idx & 31) * 2;
        S[k2 * LD + r] = v[u].x;
        S[(k2 + 1) * LD + r] = v[u].y;
    }
}

// 1/sqrt(x) in fp64: MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-22) and one third-order correction
//   e = 1 - x y^2,  y <- y + y e (1/2 + 3/8 e)          (error O(e^3) ~ 2^-66)
// Branch-free: it sits on the critical path of every Cholesky pivot.  Valid for normal positive x.
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);
}


// Warp-level Cholesky of a 32x32 block: lane r holds row r in registers.  On exit a[c] = L[r][c] (0 above the
// diagonal).  colbuf: 32 doubles of shared memory private to the warp.  No block-wide barriers.
__device__ __forceinline__ bool warp_potrf32(double (&a)[32], double *colbuf, int lane, double &my_inv) {
    bool ok = true;
    my_inv = 0.0;
    // every lane tracks its own diagonal entry so that the next pivot does not wait for the column exchange
    double diag = 0.0;
#pragma unroll
    for (int c = 0; c < 32; c++)
        if (lane == c) diag = a[c];
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const double piv = __shfl_sync(0xffffffffu, diag, j);
        // the validity test stays OFF the dependency chain: a bad pivot poisons the tile with NaN/Inf, the flag makes
        // the solve return dx = 0 anyway (geom_kernels.cu:1186-1188)
        ok = ok && (piv > 1e-290) && (piv < 1e290);
        const double inv = fast_rsqrt(piv);
        if (lane == j) my_inv = inv;
        double l = ((lane == j) ? piv : a[j]) * inv;  // lane j: piv * inv = sqrt(piv)
        if (lane < j) l = 0.0;
        a[j] = l;
        diag = fma(-l, l, diag);  // lanes > j: a_rr -= l_rj^2  (lane j's diag is not used again)
        double *cb = colbuf + (j & 1) * 32;
        cb[lane] = l;
        __syncwarp();
#pragma unroll
        for (int k = j + 1; k < 32; k++) a[k] = fma(-l, cb[k], a[k]);
    }
    __syncwarp();
    return ok;
}

// One copy of the unrolled register kernel in the binary: the second call on a tile then runs out of a warm
// instruction cache (straight-line code executed once costs ~3 clk per instruction in fetch).  Called by one warp;
// factorises the 32x32 block at Dblk (row stride RS) in place and writes 1/diag.
__device__ __noinline__ bool warp_potrf32_smem(double *Dblk, double *dinv_out, double *colbuf) {
    const int lane = threadIdx.x & 31;
    double a[32];
#pragma unroll
    for (int c = 0; c < 32; c++) a[c] = Dblk[lane * RS + c];
    double my_inv;
    const bool ok = warp_potrf32(a, colbuf, lane, my_inv);
#pragma unroll
    for (int c = 0; c < 32; c++) Dblk[lane * RS + c] = a[c];
    dinv_out[lane] = my_inv;
    return ok;
}

template <int NBLK>
__device__ __noinline__ void tile_trsm_mma_t(double *X, const double *L, const double *dinv, double *linv8, double *tmp);

// Tensor-core version of tile_potrf for tiles stored with row stride RS: potrf32 (warp 0, registers) ->
// L21 = A21 L11^-T (DMMA) -> A22 -= L21 L21^T (DMMA) -> potrf32.  All threads must call.
struct NoMid {
    __device__ __forceinline__ void operator()(int) const {}
};
// `mid(t)`, t = 0..223, runs on warps 1..7 while warp 0 factorises the second 32x32 block (columns 0..31 of L are final
// by then): the dataflow solver sends the first half of the tile to its readers there.
template <class Mid = NoMid>
__device__ bool tile_potrf_mma(double *D, double *dinv, double *colbuf, double *linv8, double *tmp, int *sh_ok, Mid mid = Mid()) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        const bool ok = warp_potrf32_smem(D, dinv, colbuf);
        if (!ok && lane == 0) *sh_ok = 0;
    }
    __syncthreads();
    tile_trsm_mma_t<4>(D + 32 * RS, D, dinv, linv8, tmp);  // rows 32..63, columns 0..31
    if (warp < 4) {  // A22 -= L21 L21^T: warp w -> columns 32+8w.., 4 row tiles
        const int fr = lane >> 2, fk = lane & 3;
        double acc[4][2];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
            const double2 v = *reinterpret_cast<const double2 *>(D + (32 + 8 * mt + fr) * RS + 32 + 8 * warp + 2 * fk);
            acc[mt][0] = v.x, acc[mt][1] = v.y;
        }
        const double *bp = D + (32 + 8 * warp + fr) * RS + fk;
        const double *ap = D + (32 + fr) * RS + fk;
        double acc2[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};  // two chains: a dependent DMMA follows after 32-42 clk
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 8) {
            const double bneg = -bp[k0], bneg2 = -bp[k0 + 4];
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                dmma884(acc[mt], ap[mt * 8 * RS + k0], bneg);
                dmma884(acc2[mt], ap[mt * 8 * RS + k0 + 4], bneg2);
            }
        }
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
            *reinterpret_cast<double2 *>(D + (32 + 8 * mt + fr) * RS + 32 + 8 * warp + 2 * fk) =
                make_double2(acc[mt][0] + acc2[mt][0], acc[mt][1] + acc2[mt][1]);
    }
    __syncthreads();
    if (warp == 0) {
        const bool ok = warp_potrf32_smem(D + 32 * RS + 32, dinv + 32, colbuf);
        if (!ok && lane == 0) *sh_ok = 0;
    } else {
        mid(tid - 32);
    }
    __syncthreads();
    return *sh_ok != 0;
}

// X <- X L^-T on the fp64 tensor cores.  X[8 NBLK][RS] and L[8 NBLK][RS] (lower, row-major) live in shared memory.
// Right-looking over column blocks of 8:  X_b = S_b inv(L_bb)^T  (the 8x8 inverses of the diagonal blocks are computed
// here by 8 NBLK threads), then every later block is updated at once, S_b' -= X_b L_{b',b}^T for b' > b.  What matters
// is the number of DEPENDENT steps (a DMMA result is back after 32-42 clk, a fragment's round trip through shared
// memory after ~85): per column block one DMMA for X_b (its two k-halves go to separate accumulators) and one for
// the update of block b+1 (all other updates are independent of it) -- the left-looking form had b + 2 of them.  Warp w owns rows 8w..8w+7 end to end, so the only synchronisation
// inside is __syncwarp.
//   linv8: [8][96] scratch (8x8 inverse blocks, row stride 12);  tmp: [8 warps][160] scratch (8x8, row stride 20)
template <int NBLK>
__device__ __noinline__ void tile_trsm_mma_t(double *X, const double *L, const double *dinv, double *linv8, double *tmp) {
    // NBLK: column blocks of 8 (size of L / 8) = row groups of 8 (rows of X / 8), one warp each
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 8 * NBLK) {  // inverse of the 8x8 lower-triangular diagonal block b, column c
        const int b = tid >> 3, c = tid & 7;
        double x[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
            for (int p = 0; p < 8; p++)
                if (p < r && p >= c) s = fma(-L[(8 * b + r) * RS + 8 * b + p], x[p], s);
            x[r] = (r >= c) ? s * dinv[8 * b + r] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 8; r++) linv8[b * 96 + r * 12 + c] = x[r];
    }
    __syncthreads();
    const int fr = lane >> 2, fk = lane & 3;  // fragment row / k index
    if (warp < NBLK) {
        double *xrow = X + (8 * warp + fr) * RS;
        double *tw = tmp + warp * 160;
        double c[NBLK][2], c2[NBLK][2];  // the two k-halves of every update accumulate separately
#pragma unroll
        for (int b = 0; b < NBLK; b++) {
            const double2 v = *reinterpret_cast<const double2 *>(xrow + 8 * b + 2 * fk);
            c[b][0] = v.x, c[b][1] = v.y;
            c2[b][0] = c2[b][1] = 0.0;
        }
#pragma unroll
        for (int b = 0; b < NBLK; b++) {
            *reinterpret_cast<double2 *>(tw + fr * 20 + 2 * fk) = make_double2(c[b][0] + c2[b][0], c[b][1] + c2[b][1]);
            __syncwarp();
            double x0[2] = {0.0, 0.0}, x1[2] = {0.0, 0.0};
            dmma884(x0, tw[fr * 20 + fk], linv8[b * 96 + fr * 12 + fk]);
            dmma884(x1, tw[fr * 20 + 4 + fk], linv8[b * 96 + fr * 12 + 4 + fk]);
            *reinterpret_cast<double2 *>(xrow + 8 * b + 2 * fk) = make_double2(x0[0] + x1[0], x0[1] + x1[1]);
            __syncwarp();
            if (b + 1 < NBLK) {
                const double a0 = xrow[8 * b + fk], a1 = xrow[8 * b + 4 + fk];
#pragma unroll
                for (int bp = b + 1; bp < NBLK; bp++) {
                    const double *lrow = L + (8 * bp + fr) * RS + 8 * b + fk;
                    dmma884(c[bp], a0, -lrow[0]);
                    dmma884(c2[bp], a1, -lrow[4]);
                }
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void tile_trsm_mma(double *X, const double *L, const double *dinv, double *linv8, double *tmp) {
    tile_trsm_mma_t<8>(X, L, dinv, linv8, tmp);
}

// X <- X L_jj^-T with L_jj taken from its fast copy `ft` as the two halves arrive (see fast_half_store): same right-looking
// solve as tile_trsm_mma_t<8>, column blocks 0..3 after half A, 4..7 after half B.  L (shared, row stride RS), dinv:
// staging for the tile.  All threads must call; ends with a block barrier.
__device__ __noinline__ void tile_trsm_fast(double *X, double *L, double *dinv, double *linv8, double *tmp, const unsigned long long *ft) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    double *xrow = X + (8 * warp + fr) * RS;
    double *tw = tmp + warp * 160;
    double c[8][2], c2[8][2];
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const double2 v = *reinterpret_cast<const double2 *>(xrow + 8 * b + 2 * fk);
        c[b][0] = v.x, c[b][1] = v.y;
        c2[b][0] = c2[b][1] = 0.0;
    }
#pragma unroll
    for (int half = 0; half < 2; half++) {
        if (half == 0) fast_half_load<0>(L, dinv, ft);
        else fast_half_load<1>(L, dinv, ft);
        __syncthreads();
        if (tid < 32) {  // inverses of this half's four 8x8 diagonal blocks, one column per thread
            const int b = 4 * half + (tid >> 3), cc = tid & 7;
            double x[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                double sacc = (r == cc) ? 1.0 : 0.0;
#pragma unroll
                for (int p = 0; p < 8; p++)
                    if (p < r && p >= cc) sacc = fma(-L[(8 * b + r) * RS + 8 * b + p], x[p], sacc);
                x[r] = (r >= cc) ? sacc * dinv[8 * b + r] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < 8; r++) linv8[b * 96 + r * 12 + cc] = x[r];
        }
        __syncthreads();
#pragma unroll
        for (int bb = 0; bb < 4; bb++) {
            const int b = 4 * half + bb;
            *reinterpret_cast<double2 *>(tw + fr * 20 + 2 * fk) = make_double2(c[b][0] + c2[b][0], c[b][1] + c2[b][1]);
            __syncwarp();
            double x0[2] = {0.0, 0.0}, x1[2] = {0.0, 0.0};
            dmma884(x0, tw[fr * 20 + fk], linv8[b * 96 + fr * 12 + fk]);
            dmma884(x1, tw[fr * 20 + 4 + fk], linv8[b * 96 + fr * 12 + 4 + fk]);
            *reinterpret_cast<double2 *>(xrow + 8 * b + 2 * fk) = make_double2(x0[0] + x1[0], x0[1] + x1[1]);
            __syncwarp();
            if (b + 1 < 8) {
                const double a0 = xrow[8 * b + fk], a1 = xrow[8 * b + 4 + fk];
#pragma unroll
                for (int bp = b + 1; bp < 8; bp++) {
                    const double *lrow = L + (8 * bp + fr) * RS + 8 * b + fk;
                    dmma884(c[bp], a0, -lrow[0]);
                    dmma884(c2[bp], a1, -lrow[4]);
                }
            }
        }
    }
    __syncthreads();
}

// fragment-distributed tile (acc) -> row-major shared tile S[64][RS]
__device__ __forceinline__ void acc_to_smem_rs(const double (&acc)[8][2], double *S, int cw) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int mt = 0; mt < 8; mt++)
        *reinterpret_cast<double2 *>(S + (8 * mt + (lane >> 2)) * RS + 8 * cw + 2 * (lane & 3)) = make_double2(acc[mt][0], acc[mt][1]);
}
__device__ __forceinline__ void acc_to_smem_rs(const double (&acc)[8][2], double *S) { acc_to_smem_rs(acc, S, threadIdx.x >> 5); }

// MINB = 1: latency-bound sizes, the whole register file for the unrolled register kernels; MINB = 2: throughput-bound
// DIST: the distributed factorisation (a compile-time switch: the single-GPU kernel is latency-bound and keeps its exact code)
template <int MINB, bool DIST>
__global__ void __launch_bounds__(CT, MINB) chol_factor_kernel(const CholArgs a) {
    extern __shared__ __align__(16) double sm[];
    double *As = sm;             // [64][RS]
    double *Bs = sm + TB * RS;   // [64][RS]
    double *col = Bs + TB * RS;  // [64]
    double *dinv = col + TB;     // [64]
    double *linv8 = dinv + TB;   // [8][96]
    double *tmpw = linv8 + 8 * 96;  // [8][160]
    double *As2 = tmpw + 8 * 160;   // second tile pair of the double-buffered k-loop (MINB == 1 only)
    double *Bs2 = As2 + TB * RS;
    __shared__ int sh_tile, sh_ok, sh_ready, sh_nk;
    __shared__ unsigned short klist[kMaxSparseT];
    const int T = a.T, ld = a.ld, tid = threadIdx.x;
    const int total = T * (T + 1) / 2 + T;  // lower tiles + one rhs tile per column
    constexpr bool dist = DIST;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            sh_tile = atomicAdd(a.counter, 1);
            sh_ok = 1;
        }
        __syncthreads();
        const int t = sh_tile;
        if (t >= total) return;
        // column j holds T - j + 1 tiles (rows j..T, row T = rhs); start(j) = j*(T+1) - j*(j-1)/2
        int j = 0;
        {
            int lo = 0, hi = T - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (mid * (T + 1) - mid * (mid - 1) / 2 <= t) lo = mid;
                else hi = mid - 1;
            }
            j = lo;
        }
        const int i = j + (t - (j * (T + 1) - j * (j - 1) / 2));
        const int j0 = j * TB;
        if (dist && col_owner(a, j) != a.rank) continue;  // another rank's column
        // the two shortcuts of the column-to-column hand-over (pre-solve tile parked for the next diagonal tile) need the
        // neighbouring column on this rank
        const bool next_local = !dist || (j + 1 < T && col_owner(a, j + 1) == a.rank);
        const bool prev_local = !dist || (j > 0 && col_owner(a, j - 1) == a.rank);
        const unsigned char *ts = a.ts;
        if (ts && i < T && !ts[(size_t)i * T + j]) continue;  // structurally zero tile: nothing to do, nobody waits for it
        TRACE(t, 0);
        // the update steps of this tile: k < j with L_ik and L_jk both structurally non-zero, ascending (dense: all of them)
        if (ts) {
            if (tid < 32) {
                int cnt = 0;
                for (int k0 = 0; k0 < j; k0 += 32) {
                    const int k = k0 + tid;
                    const bool need = k < j && (i == T || ts[(size_t)i * T + k]) && ts[(size_t)j * T + k];
                    const unsigned m = __ballot_sync(0xffffffffu, need);
                    if (need) klist[cnt + __popc(m & ((1u << tid) - 1u))] = (unsigned short)k;
                    cnt += __popc(m);
                }
                if (tid == 0) sh_nk = cnt;
            }
            __syncthreads();
        }
        const int nk = ts ? sh_nk : j;
        auto kof = [&](int idx) -> int { return ts ? (int)klist[idx] : idx; };

        if (i == T) {
            // ---- right-hand side tile: y_j = L_jj^-1 (b_j - sum_k L_jk y_k)
            double part = 0.0;
            const int c = tid & 63, q = tid >> 6;  // 4 threads per entry, each takes 16 of the 64 k's
            for (int idx = 0; idx < nk; idx++) {
                const int k = kof(idx);
                wait_flag(a.flags + (size_t)T * T + k, a.epoch, dist);
                wait_flag(a.flags + (size_t)j * T + k, a.epoch, dist);
                const double *Lrow = a.H + (size_t)(j0 + c) * ld + k * TB + q * 16;
                const double *yk = a.b + k * TB + q * 16;
                for (;;) {
                    double add = 0.0;
                    bool bad = false;
#pragma unroll
                    for (int p = 0; p < 16; p += 2) {
                        const double2 l = __ldcg(reinterpret_cast<const double2 *>(Lrow + p));
                        const double2 y = __ldcg(reinterpret_cast<const double2 *>(yk + p));
                        bad |= is_hole(l.x) | is_hole(l.y) | is_hole(y.x) | is_hole(y.y);
                        add = fma(l.x, y.x, fma(l.y, y.y, add));
                    }
                    if (!dist || !bad) {  // (distributed: the flags are hints, the words validate themselves)
                        part += add;
                        break;
                    }
                    __nanosleep(100);
                }
            }
            As[q * TB + c] = part;
            wait_flag(a.flags + (size_t)j * T + j, a.epoch, dist);
            if (dist) load_tile_T_valid(Bs, a.H + (size_t)j0 * ld + j0, ld);
            else load_tile_T(Bs, a.H + (size_t)j0 * ld + j0, ld);  // Bs[k][r] = L_jj[r][k]
            if (tid < TB) dinv[tid] = __ldcg(a.dinv + j0 + tid);
            __syncthreads();
            if (tid < 32) {
                // forward substitution by one warp: lane owns entries lane and lane+32
                const double b0 = a.Ain ? mc_load_sum(a.Ain + (size_t)ld * ld + j0 + tid) : __ldcg(a.b + j0 + tid);
                const double b1 = a.Ain ? mc_load_sum(a.Ain + (size_t)ld * ld + j0 + 32 + tid) : __ldcg(a.b + j0 + 32 + tid);
                double v0 = b0 - (As[tid] + As[TB + tid] + As[2 * TB + tid] + As[3 * TB + tid]);
                double v1 = b1 - (As[32 + tid] + As[TB + 32 + tid] + As[2 * TB + 32 + tid] + As[3 * TB + 32 + tid]);
                for (int c2 = 0; c2 < TB; c2++) {
                    const double src = (c2 < 32) ? v0 : v1;
                    const double yc = __shfl_sync(0xffffffffu, src, c2 & 31) * dinv[c2];
                    if (tid == (c2 & 31)) {
                        if (c2 < 32) v0 = yc; else v1 = yc;
                    }
                    // Bs[c2][r] = L[r][c2]
                    if (tid > c2) v0 = fma(-Bs[c2 * LD + tid], yc, v0);
                    if (tid + 32 > c2) v1 = fma(-Bs[c2 * LD + tid + 32], yc, v1);
                    if (tid == (c2 & 31)) {  // keep the solved entry
                        if (c2 < 32) v0 = yc; else v1 = yc;
                    }
                }
                if (dist) {
                    mc_store(a.bmc + j0 + tid, v0);
                    mc_store(a.bmc + j0 + 32 + tid, v1);
                } else {
                    a.b[j0 + tid] = v0;
                    a.b[j0 + 32 + tid] = v1;
                }
            }
            if (dist) publish_flag_mc(a.flags_mc + (size_t)T * T + j, a.epoch);
            else publish_flag(a.flags + (size_t)T * T + j, a.epoch);
            continue;
        }

        const int i0 = i * TB;
        // ---- load A_ij into the accumulator fragments (damping on the diagonal of diagonal tiles: geom_kernels.cu:1176)
        double acc[8][2];
        const int cw = (i == j) ? diag_col(tid >> 5) : (tid >> 5);  // this warp's column block (see tile_gemm_sub)
        const int mt0 = (i == j) ? cw : 0;                          // diagonal tiles: lower blocks only
        {
            const int lane = tid & 31;
#pragma unroll
            for (int mt = 0; mt < 8; mt++) {
                const int r = 8 * mt + (lane >> 2), c = 8 * cw + 2 * (lane & 3);
                if (a.Ain) {
                    const double *src = a.Ain + (size_t)(i0 + r) * ld + j0 + c;
                    acc[mt][0] = mc_load_sum(src);
                    acc[mt][1] = mc_load_sum(src + 1);
                } else {
                    const double2 v = __ldcg(reinterpret_cast<const double2 *>(a.H + (size_t)(i0 + r) * ld + j0 + c));
                    acc[mt][0] = v.x, acc[mt][1] = v.y;
                }
                if (i == j && i0 + r < a.n && (r == c || r == c + 1)) {
                    // (static indices only: a run-time index would move the whole accumulator array to local memory)
                    const bool first = (r == c);
                    double dd = first ? acc[mt][0] : acc[mt][1];
                    if (a.dampdiag) dd = a.Ain ? mc_load_sum(a.Ain + (size_t)ld * ld + ld + i0 + r) : a.dampdiag[i0 + r];
                    const double add = (double)a.ep + (double)a.lm * dd;
                    acc[mt][0] += first ? add : 0.0;
                    acc[mt][1] += first ? 0.0 : add;
                }
            }
        }
        TRACE(t, 1);
        // ---- left-looking updates
        int kstart = 0;
        if constexpr (MINB == 1) {
            // Software-pipelined part of the k-loop (all steps but the diagonal tile's fused last one): the ready flags of
            // up to 32 steps are polled in one round trip by warp 0, and the tiles of step k+1 are copied asynchronously
            // into the second buffer pair while the tensor cores work on step k.
            // (in list positions: the diagonal tile's step k = j - 1, if it exists, is the fused last one below)
            const int kend = (prev_local && i == j && nk > 0 && kof(nk - 1) == j - 1) ? nk - 1 : nk;
            const int lane = tid & 31, warp = tid >> 5;
            int ready = 0;  // steps at list positions < ready have both of their tiles published
            auto ensure = [&](int k) {
                if (k < ready) return;
                if (warp == 0) {
                    int r = ready;
                    for (;;) {
                        const int kp = r + lane;
                        bool ok = kp < kend;
                        const int kk = ok ? kof(kp) : 0;
                        if (ok) ok = flag_is(a.flags + (size_t)i * T + kk, a.epoch, dist);
                        if (ok && i != j) ok = flag_is(a.flags + (size_t)j * T + kk, a.epoch, dist);
                        const unsigned m = __ballot_sync(0xffffffffu, ok);
                        r += (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);
                        if (r > k) break;
                        __nanosleep(32);
                    }
                    if (lane == 0) sh_ready = r;
                }
                __syncthreads();
                ready = sh_ready;
            };
            auto issue = [&](int kp) {
                const int k = kof(kp);
                double *Ad = (kp & 1) ? As2 : As, *Bd = (kp & 1) ? Bs2 : Bs;
                tile_cp_async(Ad, a.H + (size_t)i0 * ld + k * TB, ld);
                if (i != j) tile_cp_async(Bd, a.H + (size_t)j0 * ld + k * TB, ld);
            };
            if (kend > 0) {
                ensure(0);
                issue(0);
                cp_async_commit();
            }
            // two copies of the loop (off-diagonal / diagonal tile) so that the hot off-diagonal one stays a tight body
            auto kloop = [&](auto is_diag) {
                for (int k = 0; k < kend; k++) {
                    if (k + 1 < kend) {
                        ensure(k + 1);
                        issue(k + 1);  // its buffers were last read by step k-1, which ended with a barrier
                    }
                    cp_async_commit();
                    cp_async_wait<1>();
                    if (dist) {  // the hint flags may have overtaken (part of) the tiles
                        const int kk = kof(k);
                        fix_holes((k & 1) ? As2 : As, a.H + (size_t)i0 * ld + kk * TB, ld);
                        if (i != j) fix_holes((k & 1) ? Bs2 : Bs, a.H + (size_t)j0 * ld + kk * TB, ld);
                    }
                    __syncthreads();
                    if (k == kend - 1) TRACE(t, 7);
                    const double *Ak = (k & 1) ? As2 : As, *Bk = (k & 1) ? Bs2 : Bs;
                    if constexpr (decltype(is_diag)::value) tile_gemm_sub(acc, Ak, Ak, cw, mt0);
                    else tile_gemm_sub(acc, Ak, Bk);
                    __syncthreads();
                }
            };
            if (i == j) kloop(std::true_type{});
            else kloop(std::false_type{});
            cp_async_wait<0>();
            kstart = kend > 0 ? kend : 0;
        }
        for (int kp = kstart; kp < nk; kp++) {
            const int k = kof(kp);
            if (MINB == 1 && prev_local && i == j && k == j - 1) {
                // Critical path: the diagonal tile's last update needs L_{j,j-1}.  Instead of waiting for the CTA that
                // owns that tile to solve and publish it (one more trip through L2), take its pre-solve copy -- ready
                // long before -- and do the tile solve here as soon as L_{j-1,j-1} appears.
                const int kk0 = k * TB;
                wait_flag(a.preflags + k, a.epoch);
                load_tile_R(As, a.H + (size_t)kk0 * ld + j0, ld);  // pre-solve copy of tile (j, j-1), parked at (j-1, j)
                __syncthreads();
                TRACE(t, 7);
                tile_trsm_fast(As, Bs, dinv, linv8, tmpw, a.ldiag + (size_t)k * kFastTile);  // L_kk as its halves arrive
                tile_gemm_sub(acc, As, As, cw, mt0);
                __syncthreads();
                continue;
            }
            wait_flag(a.flags + (size_t)i * T + k, a.epoch, dist);
            if (i != j) wait_flag(a.flags + (size_t)j * T + k, a.epoch, dist);
            TRACE(t, 7);
            if (dist) {
                load_tile_R_valid(As, a.H + (size_t)i0 * ld + k * TB, ld);
                if (i != j) load_tile_R_valid(Bs, a.H + (size_t)j0 * ld + k * TB, ld);
            } else {
                load_tile_R(As, a.H + (size_t)i0 * ld + k * TB, ld);
                if (i != j) load_tile_R(Bs, a.H + (size_t)j0 * ld + k * TB, ld);
            }
            __syncthreads();
            if (i != j) tile_gemm_sub(acc, As, Bs);
            else tile_gemm_sub(acc, As, As, cw, mt0);
            __syncthreads();
        }
        acc_to_smem_rs(acc, As, cw);
        __syncthreads();
        TRACE(t, 2);
        if (i == j) {
            unsigned long long *ft = a.ldiag + (size_t)j * kFastTile;  // the latency-critical readers take this copy
            const bool ok = tile_potrf_mma(As, dinv, col, linv8, tmpw, &sh_ok, [&](int t2) { if (t2 < 64) fast_half_store(ft, As, dinv, 0, t2, 64); });  // two warps: more of them slow the pivot chain
            fast_half_store(ft, As, dinv, 1, tid, CT);
            TRACE(t, 4);
            if (!ok && tid == 0) {
                *a.fail = a.epoch;
                if (dist) asm volatile("multimem.st.relaxed.sys.global.b32 [%0], %1;" ::"l"(a.fail_mc), "r"(a.epoch) : "memory");
            }
            // store L_jj (zero above the diagonal) and 1/diag
            for (int idx = tid; idx < TB * TB / 2; idx += CT) {
                const int r = idx >> 5, c = (idx & 31) * 2;
                const double2 v = *reinterpret_cast<const double2 *>(As + r * RS + c);
                const double2 w = make_double2(c <= r ? v.x : 0.0, c + 1 <= r ? v.y : 0.0);
                if (dist) {
                    mc_store2(a.Hmc + (size_t)(i0 + r) * ld + j0 + c, w.x, w.y);
                } else {
                    *reinterpret_cast<double2 *>(a.H + (size_t)(i0 + r) * ld + j0 + c) = w;
                }
            }
            if (tid < TB) a.dinv[j0 + tid] = dinv[tid];  // (read by this column's own rhs tile only: stays local)
            TRACE(t, 5);
            if (dist) publish_flag_mc(a.flags_mc + (size_t)i * T + j, a.epoch);
            else publish_flag(a.flags + (size_t)i * T + j, a.epoch);
            TRACE(t, 6);
            // Off the critical path (the tile is already published): L_jj^-T = I * L_jj^-T on the tensor cores, so that
            // the backward substitution's per-block triangular solve becomes a 64x64 matrix-vector product.
            for (int idx = tid; idx < TB * TB; idx += CT) {
                const int r = idx >> 6, c = idx & 63;
                Bs[r * RS + c] = (r == c) ? 1.0 : 0.0;
                if (c > r) As[r * RS + c] = 0.0;
            }
            __syncthreads();
            tile_trsm_mma(Bs, As, dinv, linv8, tmpw);
            if (dist) {
                for (int idx = tid; idx < TB * TB / 2; idx += CT) {
                    const int r = idx >> 5, c = (idx & 31) * 2;
                    mc_store2(a.linvT_mc + (size_t)j * TB * TB + r * TB + c, Bs[r * RS + c], Bs[r * RS + c + 1]);
                }
                publish_flag_mc(a.lflags_mc + j, a.epoch);
            } else {
                for (int idx = tid; idx < TB * TB; idx += CT) {
                    const int r = idx >> 6, c = idx & 63;
                    a.linvT[(size_t)j * TB * TB + idx] = Bs[r * RS + c];
                }
            }
            continue;
        } else {
            if (MINB == 1 && next_local && i == j + 1) {  // park the pre-solve tile in the upper triangle for the next diagonal tile
                for (int idx = tid; idx < TB * TB; idx += CT) {
                    const int r = idx >> 6, c = idx & 63;
                    a.H[(size_t)(j0 + r) * ld + i0 + c] = As[r * RS + c];
                }
                publish_flag(a.preflags + j, a.epoch);
            }
            TRACE(t, 3);
            tile_trsm_fast(As, Bs, dinv, linv8, tmpw, a.ldiag + (size_t)j * kFastTile);  // L_jj as its halves arrive
            TRACE(t, 4);
            if (dist) {
                for (int idx = tid; idx < TB * TB / 2; idx += CT) {
                    const int r = idx >> 5, c = (idx & 31) * 2;
                    mc_store2(a.Hmc + (size_t)(i0 + r) * ld + j0 + c, As[r * RS + c], As[r * RS + c + 1]);
                }
            } else {
                for (int idx = tid; idx < TB * TB; idx += CT) {
                    const int r = idx >> 6, c = idx & 63;
                    a.H[(size_t)(i0 + r) * ld + j0 + c] = As[r * RS + c];
                }
            }
        }
        TRACE(t, 5);
        if (dist) publish_flag_mc(a.flags_mc + (size_t)i * T + j, a.epoch);
        else publish_flag(a.flags + (size_t)i * T + j, a.epoch);
        TRACE(t, 6);
    }
}

// backward substitution x = L^-T y, one CTA per column block (claimed in descending order).
// The diagonal tile and 1/diag are staged before the first wait so that only the x_i chain is exposed.
template <bool DIST>
__global__ void __launch_bounds__(CT) chol_backward_kernel(const CholArgs a) {
    extern __shared__ __align__(16) double bsm[];
    double *Ld = bsm;                    // [64][65] L_jj
    double *Ls = Ld + TB * (TB + 1);     // [64][65] L_ij being applied
    double *part = Ls + TB * (TB + 1);   // [4][64]
    double *xi = part + 4 * TB;          // [64]
    __shared__ int sh_j;
    const int T = a.T, ld = a.ld, tid = threadIdx.x;
    constexpr bool dist = DIST;
    const bool failed0 = !DIST && *a.fail == a.epoch;  // (one GPU: the factorisation is complete when this kernel starts)
    for (;;) {
        __syncthreads();
        if (tid == 0) sh_j = T - 1 - atomicAdd(a.counter + 1, 1);
        __syncthreads();
        const int j = sh_j;
        if (j < 0) return;
        const int j0 = j * TB;
        const int c = tid & 63, q = tid >> 6;
        if (dist) {  // this column's y_j and L_jj^-T come from its owner (the local factor kernel may have finished long before)
            wait_flag(a.flags + (size_t)T * T + j, a.epoch, true);
            wait_flag(a.lflags + j, a.epoch, true);
        }
        {   // Ld <- L_jj^-T (upper triangular); all loads in flight before the first store
            double2 v[TB * TB / 2 / CT];
#pragma unroll
            for (int u = 0; u < TB * TB / 2 / CT; u++) {
                const int idx = tid + u * CT;
                const double2 *src = reinterpret_cast<const double2 *>(a.linvT + (size_t)j * TB * TB + (idx >> 5) * TB + (idx & 31) * 2);
                v[u] = __ldcg(src);
                while (dist && (is_hole(v[u].x) | is_hole(v[u].y))) v[u] = __ldcg(src);
            }
#pragma unroll
            for (int u = 0; u < TB * TB / 2 / CT; u++) {
                const int idx = tid + u * CT;
                const int r = idx >> 5, k2 = (idx & 31) * 2;
                Ld[r * (TB + 1) + k2] = v[u].x;
                Ld[r * (TB + 1) + k2 + 1] = v[u].y;
            }
        }
        double yj = (tid < TB) ? __ldcg(a.b + j0 + tid) : 0.0;  // y_j (from the factor kernel), fetched off the x chain
        while (dist && tid < TB && is_hole(yj)) yj = __ldcg(a.b + j0 + tid);
        double s = 0.0;
        for (int i = T - 1; i > j; i--) {
            if (a.ts && !a.ts[(size_t)i * T + j]) continue;  // L_ij is structurally zero
            if (dist) wait_flag(a.flags + (size_t)i * T + j, a.epoch, true);
            // stage L_ij while x_i may still be in flight
            double2 v[TB * TB / 2 / CT];
#pragma unroll
            for (int u = 0; u < TB * TB / 2 / CT; u++) {
                const int idx = tid + u * CT;
                const int r = idx >> 5, k2 = (idx & 31) * 2;
                const double2 *src = reinterpret_cast<const double2 *>(a.H + (size_t)(i * TB + r) * ld + j0 + k2);
                v[u] = __ldcg(src);
                while (dist && (is_hole(v[u].x) | is_hole(v[u].y))) v[u] = __ldcg(src);
            }
#pragma unroll
            for (int u = 0; u < TB * TB / 2 / CT; u++) {
                const int idx = tid + u * CT;
                const int r = idx >> 5, k2 = (idx & 31) * 2;
                Ls[r * (TB + 1) + k2] = v[u].x;
                Ls[r * (TB + 1) + k2 + 1] = v[u].y;
            }
            // x_i arrives as self-validating 8-byte words (zero bits = not written yet; a true +0.0 travels as -0.0): every
            // thread polls its own entry, one trip to L2 after the data lands instead of fence + flag + flag poll + load
            if (tid < TB) {
                unsigned long long v;
                do {
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.xs + i * TB + tid) : "memory");
                } while (v == 0ull);
                xi[tid] = __longlong_as_double((long long)v);
            }
            __syncthreads();
#pragma unroll
            for (int p = 0; p < 16; p++) s = fma(Ls[(q * 16 + p) * (TB + 1) + c], xi[q * 16 + p], s);
            __syncthreads();
        }
        part[q * TB + c] = s;
        __syncthreads();
        if (tid < TB) xi[tid] = yj - (part[tid] + part[TB + tid] + part[2 * TB + tid] + part[3 * TB + tid]);
        __syncthreads();
        {   // x_j = L_jj^-T v : thread (c, q) sums a quarter of row c, quarters combined through shared memory
            double acc = 0.0;
#pragma unroll
            for (int p = 0; p < 16; p++) acc = fma(Ld[c * (TB + 1) + q * 16 + p], xi[q * 16 + p], acc);
            part[q * TB + c] = acc;
        }
        __syncthreads();
        if (tid < TB) {
            const double x = part[tid] + part[TB + tid] + part[2 * TB + tid] + part[3 * TB + tid];
            unsigned long long bits = (unsigned long long)__double_as_longlong(x);
            if (bits == 0ull) bits = 0x8000000000000000ull;
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(a.xs + j0 + tid), "l"(bits) : "memory");
            a.b[j0 + tid] = x;
            // (a failure anywhere was written into every rank's flag before the failing column's tiles were published)
            const bool failed = DIST ? *reinterpret_cast<volatile int *>(a.fail) == a.epoch : failed0;
            if (j0 + tid < a.n) a.dx[a.rowmap ? a.rowmap[j0 + tid] : j0 + tid] = failed ? 0.0f : (float)x;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Systems of at most two tiles (6P <= 128: the frontend window and the inner filler) are solved by ONE CTA with
// everything resident in shared memory: no ready flags, no hops through L2, one launch.
//   [A00 .  ]   potrf(A00) -> L10 = A10 L00^-T -> A11 -= L10 L10^T -> potrf(A11) -> two triangular solves
//   [A10 A11]
template <int S>
__device__ void smem_trsv_fwd(const double *L, const double *dinv, double *v, int lane) {
    // L row-major [64][S]; solves L y = v in place (one warp)
    double v0 = v[lane], v1 = v[lane + 32];
    for (int c = 0; c < TB; c++) {
        const double yc = __shfl_sync(0xffffffffu, (c < 32) ? v0 : v1, c & 31) * dinv[c];
        if (lane > c) v0 = fma(-L[lane * S + c], yc, v0);
        if (lane + 32 > c) v1 = fma(-L[(lane + 32) * S + c], yc, v1);
        if (lane == (c & 31)) {
            if (c < 32) v0 = yc; else v1 = yc;
        }
    }
    v[lane] = v0;
    v[lane + 32] = v1;
}
template <int S>
__device__ void smem_trsv_bwd(const double *L, const double *dinv, double *v, int lane) {
    // solves L^T x = v in place (one warp)
    double v0 = v[lane], v1 = v[lane + 32];
    for (int c = TB - 1; c >= 0; c--) {
        const double xc = __shfl_sync(0xffffffffu, (c < 32) ? v0 : v1, c & 31) * dinv[c];
        if (lane < c) v0 = fma(-L[c * S + lane], xc, v0);
        if (lane + 32 < c) v1 = fma(-L[c * S + lane + 32], xc, v1);
        if (lane == (c & 31)) {
            if (c < 32) v0 = xc; else v1 = xc;
        }
    }
    v[lane] = v0;
    v[lane + 32] = v1;
}

// Systems of at most two tiles (frontend windows, inner filler: 6P <= 128) entirely in shared memory, one CTA, with the
// same tensor-core tile kernels as the dataflow solver.
// Batched form: blockIdx.x selects the problem when `prob_hoff` is non-null (many small independent problems, e.g.
// 64 motion-only clips); otherwise the explicit H / b / n / ld arguments describe the one problem.
__global__ void __launch_bounds__(CT) chol_small_kernel(const double *__restrict__ H, const double *__restrict__ b, int n,
                                                        int ld, int T, float lm, float ep, float *__restrict__ dx,
                                                        const double *__restrict__ dampdiag,
                                                        const long long *__restrict__ prob_hoff, const int *__restrict__ prob_n,
                                                        const int *__restrict__ prob_npad, const int *__restrict__ prob_row0,
                                                        int damp_on_A) {
    extern __shared__ __align__(16) double sm[];
    if (prob_hoff) {
        const int c = blockIdx.x;
        ld = prob_npad[c];
        n = prob_n[c];
        T = ld / TB;
        H = H + prob_hoff[c];
        b = H + (size_t)ld * ld;
        dampdiag = damp_on_A ? b + ld : nullptr;
        dx = dx + 6 * (size_t)prob_row0[c];
    }
    double *A00 = sm;                   // [64][RS]
    double *A10 = A00 + TB * RS;        // [64][RS]
    double *A11 = A10 + TB * RS;        // [64][RS]
    double *dinv0 = A11 + TB * RS;      // [64]
    double *dinv1 = dinv0 + TB;         // [64]
    double *col = dinv1 + TB;           // [64]
    double *linv8 = col + TB;           // [8][96]
    double *tmpw = linv8 + 8 * 96;      // [8][160]
    double *rhs = tmpw + 8 * 160;       // [128]
    __shared__ int sh_ok;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sh_ok = 1;
    {
        // All loads of the (at most three) tiles leave before the first one is used: a load-use-store loop serialises on
        // the trip to L2 (issue is in order), which was a third of this kernel's time at frontend size.
        constexpr int U = TB * TB / 2 / CT;  // double2 per thread per tile
        double2 v0[U], v1[U], v2[U];
        double dd0[U], dd1[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int idx = tid + u * CT;
            const int r = idx >> 5, c = (idx & 31) * 2;
            v0[u] = __ldg(reinterpret_cast<const double2 *>(H + (size_t)r * ld + c));
            const bool on_diag = (r == c || r == c + 1);
            dd0[u] = (on_diag && dampdiag) ? __ldg(dampdiag + r) : 0.0;
            if (T == 2) {
                v1[u] = __ldg(reinterpret_cast<const double2 *>(H + (size_t)(TB + r) * ld + c));
                v2[u] = __ldg(reinterpret_cast<const double2 *>(H + (size_t)(TB + r) * ld + TB + c));
                dd1[u] = (on_diag && dampdiag) ? __ldg(dampdiag + TB + r) : 0.0;
            }
        }
        auto damp = [&](double2 v, int r, int c, int row, double dd) {  // lower triangle + diag += ep + lm * diag (geom_kernels.cu:1176)
            double x = (c <= r) ? v.x : 0.0, y = (c + 1 <= r) ? v.y : 0.0;
            if (row < n) {
                if (r == c) x += (double)ep + (double)lm * (dampdiag ? dd : x);
                if (r == c + 1) y += (double)ep + (double)lm * (dampdiag ? dd : y);
            }
            return make_double2(x, y);
        };
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int idx = tid + u * CT;
            const int r = idx >> 5, c = (idx & 31) * 2;
            *reinterpret_cast<double2 *>(A00 + r * RS + c) = damp(v0[u], r, c, r, dd0[u]);
            if (T == 2) {
                *reinterpret_cast<double2 *>(A10 + r * RS + c) = v1[u];
                *reinterpret_cast<double2 *>(A11 + r * RS + c) = damp(v2[u], r, c, TB + r, dd1[u]);
            }
        }
    }
    for (int i = tid; i < T * TB; i += CT) rhs[i] = b[i];
    __syncthreads();
    bool ok = tile_potrf_mma(A00, dinv0, col, linv8, tmpw, &sh_ok);
    if (T == 2) {
        tile_trsm_mma(A10, A00, dinv0, linv8, tmpw);  // L10 = A10 L00^-T
        __syncthreads();
        {   // A11 -= L10 L10^T on the tensor cores
            double acc[8][2];
#pragma unroll
            for (int mt = 0; mt < 8; mt++) {
                const double2 v = *reinterpret_cast<const double2 *>(A11 + (8 * mt + (lane >> 2)) * RS + 8 * warp + 2 * (lane & 3));
                acc[mt][0] = v.x, acc[mt][1] = v.y;
            }
            tile_gemm_sub(acc, A10, A10);
            __syncthreads();
            acc_to_smem_rs(acc, A11);
        }
        __syncthreads();
        if (n - TB <= 32) {
            // at most 32 real rows in the second tile: the rest is the identity padding (its rows of L10 are zero), so only
            // the leading 32x32 block needs factorising -- half of the pivot chain
            if (warp == 0) {
                const bool ok1 = warp_potrf32_smem(A11, dinv1, col);
                if (!ok1 && lane == 0) sh_ok = 0;
            } else if (warp == 1) {
                dinv1[32 + lane] = 1.0;
            }
            __syncthreads();
            ok = (sh_ok != 0) && ok;
        } else {
            ok = tile_potrf_mma(A11, dinv1, col, linv8, tmpw, &sh_ok) && ok;
        }
    }
    if (!ok) {  // failed factorisation => zero update (geom_kernels.cu:1186-1188)
        for (int i = tid; i < n; i += CT) dx[i] = 0.0f;
        return;
    }
    // forward: y0 = L00^-1 b0 ; y1 = L11^-1 (b1 - L10 y0)
    if (warp == 0) smem_trsv_fwd<RS>(A00, dinv0, rhs, lane);
    __syncthreads();
    if (T == 2) {
        if (tid < TB) {
            double s = 0.0;
            for (int k = 0; k < TB; k++) s = fma(A10[tid * RS + k], rhs[k], s);
            rhs[TB + tid] -= s;
        }
        __syncthreads();
        if (warp == 0) {
            smem_trsv_fwd<RS>(A11, dinv1, rhs + TB, lane);
            smem_trsv_bwd<RS>(A11, dinv1, rhs + TB, lane);
        }
        __syncthreads();
        if (tid < TB) {
            double s = 0.0;
            for (int k = 0; k < TB; k++) s = fma(A10[k * RS + tid], rhs[TB + k], s);
            rhs[tid] -= s;
        }
        __syncthreads();
    }
    if (warp == 0) smem_trsv_bwd<RS>(A00, dinv0, rhs, lane);
    __syncthreads();
    for (int i = tid; i < n; i += CT) dx[i] = (float)rhs[i];
}

static cudaError_t launch_small_solve(double *H, double *b, int n, int npad, float lm, float ep, float *dx,
                                      const double *dampdiag, cudaStream_t st, int *launches) {
    const size_t sm = (size_t)(3 * TB * RS + 3 * TB + 8 * 96 + 8 * 160 + 2 * TB) * sizeof(double);
    cudaError_t err = cudaFuncSetAttribute(chol_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (err != cudaSuccess) return err;
    chol_small_kernel<<<1, CT, sm, st>>>(H, b, n, npad, npad / TB, lm, ep, dx, dampdiag, nullptr, nullptr, nullptr, nullptr, 0);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_small_solve_batch(double *sys, const long long *prob_hoff, const int *prob_n, const int *prob_npad,
                                     const int *prob_row0, int n_prob, float lm, float ep, float *dx, bool damp_on_A,
                                     cudaStream_t st, int *launches) {
    const size_t sm = (size_t)(3 * TB * RS + 3 * TB + 8 * 96 + 8 * 160 + 2 * TB) * sizeof(double);
    cudaError_t err = cudaFuncSetAttribute(chol_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (err != cudaSuccess) return err;
    chol_small_kernel<<<n_prob, CT, sm, st>>>(sys, nullptr, 0, 0, 0, lm, ep, dx, nullptr, prob_hoff, prob_n, prob_npad,
                                             prob_row0, damp_on_A ? 1 : 0);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_damped_solve(double *H, double *b, int n, int npad, float lm, float ep, float *dx, int *scratch,
                                double *dinv, double *linvT, const double *dampdiag, const double *Ain, int epoch,
                                const unsigned char *tstruct, const int *rowmap, cudaStream_t st, int *launches,
                                const CholDist *dist) {
    // scratch (ints): [0..1] counters, [2] fail, [16 .. 16+T) unused, [16+T .. 16+2T) preflags, (T+1)*T tile flags, then
    // (16-byte aligned) npad 64-bit words for the backward substitution's x exchange and T * (64*64 + 64) words for the
    // diagonal tiles' fast copies; all zeroed by the memset below
    const int T = npad / TB;
    (void)epoch;
    if (T <= 2) {
        if (Ain) return cudaErrorNotSupported;  // the fused multi-GPU input needs the tiled solver (callers check)
        return launch_small_solve(H, b, n, npad, lm, ep, dx, dampdiag, st, launches);
    }
    // ready flags, counters and the failure flag are reset by one small memset per solve (graph-capturable, and the
    // flags never carry state from one solve to the next)
    // (distributed: the tile flags carry the solve's epoch and are never reset -- a reset could wipe what a faster rank has
    // already published; only the rank-local parts are cleared: counters, the x words, the diagonal tiles' fast copies)
    const size_t flag_lo = 16, flag_hi = (16 + 2 * (size_t)T + (size_t)(T + 1) * T + 3) & ~(size_t)3;
    cudaError_t err;
    if (dist) {
        if (!Ain || !dist->Hmc || !dist->scratch_mc || !dist->linvT_mc) return cudaErrorInvalidValue;
        err = cudaMemsetAsync(scratch, 0, 2 * sizeof(int), st);
        if (err != cudaSuccess) return err;
        err = cudaMemsetAsync(scratch + flag_hi, 0, (chol_scratch_ints(npad) - flag_hi) * sizeof(int), st);
    } else {
        err = cudaMemsetAsync(scratch, 0, chol_scratch_ints(npad) * sizeof(int), st);
    }
    (void)flag_lo;
    if (err != cudaSuccess) return err;
    CholArgs a;
    a.H = H;
    a.b = b;
    a.ld = npad;
    a.T = T;
    a.n = n;
    a.lm = lm;
    a.ep = ep;
    a.counter = scratch;
    a.fail = scratch + 2;
    a.preflags = scratch + 16 + T;
    a.flags = scratch + 16 + 2 * T;
    a.xs = reinterpret_cast<unsigned long long *>(scratch + ((16 + 2 * (size_t)T + (size_t)(T + 1) * T + 3) & ~(size_t)3));  // 16-byte aligned
    a.ldiag = a.xs + npad;
    a.epoch = dist ? dist->epoch : 1;
    if (dist) {
        a.rank = dist->rank, a.world = dist->world;
        a.Hmc = dist->Hmc;
        a.bmc = dist->Hmc + (size_t)npad * npad;
        a.linvT_mc = dist->linvT_mc;
        a.fail_mc = dist->scratch_mc + 2;
        a.colblk = dist->colblk > 0 ? dist->colblk : 1;

        a.lflags = scratch + 16;
        a.lflags_mc = dist->scratch_mc + 16;
        a.flags_mc = dist->scratch_mc + 16 + 2 * T;
    }
    a.dx = dx;
    a.dinv = dinv;
    a.dampdiag = dampdiag;
    a.linvT = linvT;
    a.Ain = Ain;
    a.ts = (T <= kMaxSparseT) ? tstruct : nullptr;
    a.rowmap = rowmap;
    // MINB = 1 double-buffers the tiles of its k-loop (4 tile buffers); MINB = 2 keeps two so that two CTAs fit an SM
    const size_t sm = (size_t)(4 * TB * RS + 2 * TB + 8 * 96 + 8 * 160) * sizeof(double);
    const size_t sm2 = (size_t)(2 * TB * RS + 2 * TB + 8 * 96 + 8 * 160) * sizeof(double);
    err = cudaFuncSetAttribute(chol_factor_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(chol_factor_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(chol_factor_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(chol_factor_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int total = T * (T + 1) / 2 + T;
    // small systems are latency-bound: one CTA per SM gives the critical-path tiles a whole fp64 pipe;
    // large ones are throughput-bound: two CTAs per SM overlap tile loads with the tile GEMMs
    // (measured at C4, T = 94: two CTAs per SM instead of one change the solve by 2 %, 3.62 vs 3.70 ms)
    constexpr int t2 = 400;  // tile count from which two CTAs share an SM
    const int ctas = (T >= t2 ? 2 : 1) * sms;
    const int grid = total < ctas ? total : ctas;
#ifdef VBA_CHOL_TRACE
    static long long *trace_dev = nullptr;
    if (dist && std::getenv("VIPE_BA_CHOL_TRACE")) {
        if (!trace_dev) cudaMalloc(&trace_dev, (size_t)total * 8 * sizeof(long long));
        cudaMemsetAsync(trace_dev, 0, (size_t)total * 8 * sizeof(long long), st);
        cudaMemcpyToSymbolAsync(g_trace, &trace_dev, sizeof(trace_dev), 0, cudaMemcpyHostToDevice, st);
    }
#endif
    if (dist) {
        if (T >= t2) chol_factor_kernel<2, true><<<grid, CT, sm2, st>>>(a);
        else chol_factor_kernel<1, true><<<grid, CT, sm, st>>>(a);
    } else {
        if (T >= t2) chol_factor_kernel<2, false><<<grid, CT, sm2, st>>>(a);
        else chol_factor_kernel<1, false><<<grid, CT, sm, st>>>(a);
    }
#ifdef VBA_CHOL_TRACE
    if (dist && trace_dev && std::getenv("VIPE_BA_CHOL_TRACE")) {  // developer builds: dump the last solve's stamps
        cudaStreamSynchronize(st);
        std::vector<long long> host((size_t)total * 8);
        cudaMemcpy(host.data(), trace_dev, host.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        char name[256];
        snprintf(name, sizeof(name), "%s_r%d.txt", std::getenv("VIPE_BA_CHOL_TRACE"), dist->rank);
        if (FILE *f = fopen(name, "w")) {
            fprintf(f, "%d %d %d %d\n", T, dist->rank, dist->world, a.colblk);
            for (int t = 0; t < total; t++) {
                if (!host[(size_t)t * 8]) continue;
                fprintf(f, "%d", t);
                for (int q = 0; q < 8; q++) fprintf(f, " %lld", host[(size_t)t * 8 + q]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
#endif
    const size_t smb = (size_t)(2 * TB * (TB + 1) + 6 * TB) * sizeof(double);
    err = cudaFuncSetAttribute(chol_backward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(chol_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
    if (err != cudaSuccess) return err;
    if (dist) chol_backward_kernel<true><<<T < sms ? T : sms, CT, smb, st>>>(a);
    else chol_backward_kernel<false><<<T < sms ? T : sms, CT, smb, st>>>(a);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// NVLS all-reduce of the reduced camera system in two steps (reduce-scatter + all-gather through the switch).
// Work items: the T (T + 1) / 2 lower 64 x 64 tiles, then 2 T pieces of 64 entries of [b ; diag(A)]; item t belongs to rank
// t mod world.  Every element is read by exactly one rank (its owner) and written by it into all instances of `out`.
__global__ void __launch_bounds__(256) peer_reduce_kernel(const double *__restrict__ mc_in, double *__restrict__ mc_out, int ld, int T,
                                                          int rank, int world) {
    const int ntile = T * (T + 1) / 2, nitem = ntile + 2 * T;
    for (int item = rank + (int)blockIdx.x * world; item < nitem; item += (int)gridDim.x * world) {
        if (item < ntile) {
            int i = (int)((sqrtf(8.0f * (float)item + 1.0f) - 1.0f) * 0.5f);
            while ((i + 1) * (i + 2) / 2 <= item) i++;
            while (i * (i + 1) / 2 > item) i--;
            const int j = item - i * (i + 1) / 2;
            for (int idx = threadIdx.x; idx < TB * TB; idx += 256) {
                const size_t off = (size_t)(i * TB + (idx >> 6)) * ld + j * TB + (idx & 63);
                const double v = mc_load_sum(mc_in + off);
                asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc_out + off), "d"(v) : "memory");
            }
        } else if (threadIdx.x < TB) {
            const size_t off = (size_t)ld * ld + (size_t)(item - ntile) * TB + threadIdx.x;
            const double v = mc_load_sum(mc_in + off);
            asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc_out + off), "d"(v) : "memory");
        }
    }
    __threadfence_system();
}

cudaError_t launch_peer_reduce(const double *accum_mc, double *reduced_mc, int npad, int rank, int world, cudaStream_t st) {
    const int T = npad / TB;
    const int nitem = T * (T + 1) / 2 + 2 * T;
    int grid = (nitem + world - 1) / world;
    if (grid > 592) grid = 592;
    peer_reduce_kernel<<<grid, 256, 0, st>>>(accum_mc, reduced_mc, npad, T, rank, world);
    return cudaGetLastError();
}

size_t chol_scratch_ints(int npad) {
    const int T = npad / TB;
    return ((16 + 2 * (size_t)T + (size_t)(T + 1) * T + 3) & ~(size_t)3) + 2 * (size_t)npad + 2 * (size_t)T * (TB * TB + TB);
}

}  // namespace vba
