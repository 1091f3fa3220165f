// C ABI of libvipe_ba.so (include/vipe_ba.h): plan construction (index bookkeeping on the host, once per
// graph), workspace layout, and the stream-ordered Gauss-Newton loop.  No torch, no allocation, no sync.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/vipe_ba.h"
#include "ba_launch.h"

using namespace vba;

static thread_local std::string g_err;
static int fail(const std::string &msg) {
    g_err = msg;
    return 1;
}
namespace vba {
int set_last_error(const char *msg) { return fail(msg ? msg : "unknown error"); }
}  // namespace vba
#define VBA_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static int device_sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached_dev = dev;
        cached = n;
    }
    return cached;
}

struct vipe_ba_plan {
    int64_t E = 0, N = 0;
    int ht = 0, wd = 0, HW = 0, t0 = 0, t1 = 0, P = 0, K = 0;
    int rank = 0, world = 1;
    int n = 0, npad = 0;
    int NT = 256, PPT = 1, ntile = 0;
    bool packed = false;  // linearize2 (pixel-pair) kernel
    int NTm = 256, PPTm = 1, ntile_m = 0;  // motion-only tile shape (no staging buffer => always the widest)
    int k_lo = 0, k_hi = 0, dmax = 0;
    // source frames whose staging buffer does not fit shared memory go through global memory in a launch of their own
    std::vector<int> flist_hub, flist_bulk;
    int dmax_bulk = 0;
    size_t off_flist_hub = 0, off_flist_bulk = 0, off_uglobal = 0, off_fscratch = 0;
    bool use_lin3 = false;  // TMA-fed motion-only pipeline (ba_lin3.cu)
    size_t off_econst = 0, off_slot_src = 0;
    std::vector<int> slot_src;
    // second-generation pipeline (ba_lin4.cu): every owned frame has 1..kLin4MaxDeg edges
    bool use_lin4 = false;
    int lin4_grid = 0;
    std::vector<Lin4Item> lin4_items;
    std::vector<Lin4Unit> lin4_units;
    std::vector<int> lin4_cta_item, lin4_cta_unit;
    size_t off_lin4_items = 0, off_lin4_units = 0, off_lin4_cta_item = 0, off_lin4_cta_unit = 0, off_econst2 = 0, off_sq = 0,
           off_sqw = 0;
    int64_t n_triples = 0;
    std::vector<int64_t> kx, kk_exp;
    std::vector<int> kx32, fptr, fedge, e_jj;
    std::vector<long long> gbase, mbase;
    // problems (C == 1 unless created with vipe_ba_plan_create_batch)
    int C = 1;
    bool any_padding = false;
    size_t sys_doubles = 0;
    std::vector<int> pose_slot, pose_row, frame_prob, prob_npad, prob_n, prob_row0;
    // pose_sys[f] = position of pose f in the reduced system (the natural order, see compute_elimination_order), rowmap[s] = index
    // of system unknown s in the dx output, tstruct[i * T + j] = 1 when tile (i, j) of the Cholesky factor can be non-zero
    std::vector<int> pose_sys, rowmap;
    // deterministic assembly (see assemble_kernel): contribution slots per frame and, per destination, the slots to add
    std::vector<long long> cbase, vbase, bdst_off, vdst_off;
    std::vector<int> bdst_ld, bsrc_ptr, bsrc, vdst_adiag, vsrc_ptr, vsrc;
    long long n_cblk = 0, n_cvec = 0;
    size_t off_cbase = 0, off_vbase = 0, off_bdst_off = 0, off_bdst_ld = 0, off_bsrc_ptr = 0, off_bsrc = 0, off_vdst_off = 0,
           off_vdst_adiag = 0, off_vsrc_ptr = 0, off_vsrc = 0, off_cblk = 0, off_cvec = 0, off_cvec2 = 0;
    std::vector<unsigned char> tstruct;
    bool ordered = false;
    size_t off_psys = 0, off_rowmap = 0, off_tstruct = 0;
    std::vector<long long> prob_hoff;
    size_t off_pslot = 0, off_prow = 0, off_fprob = 0, off_phoff = 0, off_pnpad = 0, off_pn = 0, off_prow0 = 0;
    size_t off_pn_focal = 0, off_fpart = 0, off_ffpart = 0, off_uf = 0;  // focal-length variable
    int ntile_f = 0;
    std::vector<int64_t> own_lo, own_hi;
    // workspace layout (byte offsets)
    size_t off_kx = 0, off_fptr = 0, off_fedge = 0, off_ejj = 0, off_gbase = 0, off_mbase = 0, idx_bytes = 0;
    size_t off_epart = 0, off_gpart = 0, off_msc = 0, off_q = 0, off_qw = 0, off_sys = 0, off_dx = 0, off_flag = 0;
    size_t off_gdx = 0, off_gdz = 0;
    size_t total = 0, flag_bytes = 0;
    mutable int epoch = 0;
    std::vector<unsigned char> blob;  // host image of the index tables
    mutable int64_t launches = 0;
    // CUDA-graph replay of whole runs (small problems are launch-bound): one entry per distinct argument set
    struct GraphEntry {
        vipe_ba_tensors t;
        void *ws;
        int iterations, motion_only;
        float lm, ep;
        int seen;
        int64_t launches;
        cudaGraphExec_t exec;
    };
    mutable std::vector<GraphEntry> graphs;
    mutable cudaStream_t capture_stream = nullptr;
    mutable bool use_graphs = true;
    Options opt;  // semantic switches (vipe_ba_set_options)
    // multi-GPU fused reduction (vipe_ba_set_peer_system): where this rank accumulates its partial system, and the
    // multicast address through which the solve reads the sum over ranks
    mutable double *peer_accum = nullptr;
    mutable double *solve_buf = nullptr;  // where solve_update reads and factorises the system (null: the workspace)
    mutable int rows_by_slot = 0;         // targets/weights hold only this rank's edges, in CSR slot order
    mutable const double *peer_mc = nullptr;
    // distributed factorisation (vipe_ba_set_dist_solve): this rank's instances and the multicast aliases of the factor
    // buffer [L ; y] and of the solver's auxiliary buffer [scratch ints ; 1/diag ; L_jj^-T tiles]
    mutable double *dist_factor = nullptr, *dist_factor_mc = nullptr;
    mutable unsigned char *dist_aux = nullptr, *dist_aux_mc = nullptr;
    mutable int dist_rank = 0, dist_world = 1, dist_epoch = 0;
    // optional stage timing
    bool profile = false;
    mutable std::vector<cudaEvent_t> events;  // 5 per iteration
    mutable int prof_iters = 0;
    ~vipe_ba_plan() {
        for (auto e : events) cudaEventDestroy(e);
        for (auto &g : graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        if (capture_stream) cudaStreamDestroy(capture_stream);
    }
};
static constexpr int kMaxProfIters = 64;

extern "C" int vipe_ba_abi_version(void) { return 3; }

extern "C" void vipe_ba_options_default(vipe_ba_options *o) {
    if (!o) return;
    const Options d;
    o->min_depth = d.min_depth;
    o->depth_strict = d.depth_strict;
    o->alpha = d.alpha;
    o->sensor_mode = d.sensor_mode;
    o->eta_scale = d.eta_scale;
    o->eta_bias = d.eta_bias;
    o->dz_max = d.dz_max;
    o->renorm_quat = d.renorm_quat;
    o->damp_on_pose_hessian = d.damp_on_pose_hessian;
    o->backsub_all_poses = d.backsub_all_poses;
    o->frame_flags = nullptr;
    o->optimize_focal = d.optimize_focal;
    o->focal_jscale = d.focal_jscale;
    o->focal_lm = d.focal_lm;
    o->focal_ep = d.focal_ep;
}

extern "C" int vipe_ba_set_options(vipe_ba_plan *p, const vipe_ba_options *o) {
    if (!p || !o) return fail("null argument");
    Options n;
    n.min_depth = o->min_depth;
    n.depth_strict = o->depth_strict;
    n.alpha = o->alpha;
    n.sensor_mode = o->sensor_mode;
    n.eta_scale = o->eta_scale;
    n.eta_bias = o->eta_bias;
    n.dz_max = o->dz_max;
    n.renorm_quat = o->renorm_quat;
    n.damp_on_pose_hessian = o->damp_on_pose_hessian;
    n.backsub_all_poses = o->backsub_all_poses;
    n.frame_flags = o->frame_flags;
    n.optimize_focal = o->optimize_focal;
    n.focal_jscale = o->focal_jscale;
    n.focal_lm = o->focal_lm;
    n.focal_ep = o->focal_ep;
    if (n.optimize_focal) {
        if (p->C != 1 || p->world != 1) return fail("optimize_focal needs a single-problem, single-rank plan");
        if (!n.damp_on_pose_hessian) return fail("optimize_focal follows the Python solver: set damp_on_pose_hessian");
    }
    p->opt = n;
    // captured graphs have the old options baked in
    for (auto &g : p->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    p->graphs.clear();
    return 0;
}
extern "C" const char *vipe_ba_last_error(void) { return g_err.c_str(); }


// Tile structure of the Cholesky factor of the reduced camera system.  The structure of A - S is known from the graph alone:
// the free poses among {i} + targets(i) of one source frame i form a clique.  The solver works on 64 x 64 tiles (10.7 poses),
// so the symbolic factorisation is done at tile level and handed to the kernels as a byte map; structurally zero tiles are
// never touched.  Measured on the synthetic backends (banded + 10 % random closures): 412 of 435 tiles / 0.87 of the dense
// tile updates at C3, 0.90 at C4.  A fill-reducing POSE order (greedy minimum degree, the role of SimplicialLLT's AMD
// ordering, geom_kernels.cu:1178-1182) cuts the scalar flops to 0.24 of dense, but it interleaves unrelated poses inside a
// tile and leaves the TILE structure fully dense (435 of 435); ordering whole tile-sized groups of consecutive poses gives
// 0.87-0.94.  At this tile size the natural (temporal) order is the best of the three, so it is kept: pose_sys == pose_slot.
static void compute_elimination_order(vipe_ba_plan *p) {
    p->pose_sys = p->pose_slot;
    const int P = p->P, npad = p->npad, T = npad / kCholBlock;
    p->rowmap.resize(npad);
    for (int i = 0; i < npad; i++) p->rowmap[i] = i;
    p->tstruct.clear();
    p->ordered = false;
    const char *env = std::getenv("VIPE_BA_TILE_SKIP");
    const bool want = !(env && env[0] == '0');
    if (p->C != 1 || T <= 2 || !want) return;
    std::vector<unsigned char> ts((size_t)T * T, 0);
    std::vector<int> mem;
    for (int k = 0; k < p->K; k++) {
        mem.clear();
        const int f = (int)p->kx[k];
        if (p->pose_slot[f] >= 0) mem.push_back(p->pose_slot[f]);
        for (int s2 = p->fptr[k]; s2 < p->fptr[k + 1]; s2++) {
            const int j = p->e_jj[p->fedge[s2]];
            if (p->pose_slot[j] >= 0) mem.push_back(p->pose_slot[j]);
        }
        for (int a : mem)
            for (int b : mem) {
                const int ra = 6 * a, rb = 6 * b;
                for (int ti = ra / kCholBlock; ti <= (ra + 5) / kCholBlock; ti++)
                    for (int tj = rb / kCholBlock; tj <= (rb + 5) / kCholBlock; tj++)
                        if (ti >= tj) ts[(size_t)ti * T + tj] = 1;
            }
    }
    // the spare unknown 6P (focal length, Options::optimize_focal) couples with every pose: its tile row is dense; padding
    // rows only carry the identity
    for (int tj = 0; tj <= (6 * P) / kCholBlock; tj++) ts[(size_t)((6 * P) / kCholBlock) * T + tj] = 1;
    for (int t = 0; t < T; t++) ts[(size_t)t * T + t] = 1;
    for (int k = 0; k < T; k++)
        for (int i = k + 1; i < T; i++)
            if (ts[(size_t)i * T + k])
                for (int j = k + 1; j <= i; j++)
                    if (ts[(size_t)j * T + k]) ts[(size_t)i * T + j] = 1;
    p->tstruct = ts;
    p->ordered = true;
}


// Contribution slots and destination lists of the deterministic assembly.  Frame k (owned, d edges, source pose i, targets
// j_m) owns block slots [cbase, cbase + npairs + d + 1): the pair blocks M(m, m') in pair order, the d blocks T(m') of the
// source row, the block Z of (i, i); and vector slots [vbase, vbase + d + 1): g_m per edge, then the source's.  Destinations
// list their slots in ascending slot order (= by source frame, then by position), which fixes the summation order.
static void build_assembly_lists(vipe_ba_plan *p) {
    const int K = p->K;
    p->cbase.assign(K + 1, 0);
    p->vbase.assign(K + 1, 0);
    for (int k = 0; k < K; k++) {
        const long long d = p->fptr[k + 1] - p->fptr[k];
        const bool owned = k >= p->k_lo && k < p->k_hi && d > 0;
        p->cbase[k + 1] = p->cbase[k] + (owned ? d * (d + 1) / 2 + d + 1 : 0);
        p->vbase[k + 1] = p->vbase[k] + (owned ? d + 1 : 0);
    }
    p->n_cblk = p->cbase[K];
    p->n_cvec = p->vbase[K];
    struct BSrc { long long key; int src; };
    std::vector<BSrc> bl, vl;
    // key of a block destination: (problem, row pose, column pose) with row >= column
    auto add_block = [&](int prob, int pa, int pb, long long slot, bool transposed_if_swapped) {
        (void)transposed_if_swapped;
        int hi = pa, lo = pb, tr = 0;
        if (pa < pb) hi = pb, lo = pa, tr = 1;
        const long long key = ((long long)prob << 48) | ((long long)hi << 24) | (long long)lo;
        bl.push_back({key, (int)(slot * 2 + tr)});
    };
    std::vector<int> aj;
    for (int k = p->k_lo; k < p->k_hi; k++) {
        const int s0 = p->fptr[k], d = p->fptr[k + 1] - s0;
        if (d == 0) continue;
        const int f = (int)p->kx[k];
        const int prob = p->frame_prob[k], ai = p->pose_sys[f];
        aj.assign(d, -1);
        for (int m = 0; m < d; m++) {
            const int j = p->e_jj[p->fedge[s0 + m]];
            if (j != f) aj[m] = p->pose_sys[j];  // stereo edges (i == j) carry no pose block
        }
        const long long cb = p->cbase[k], vb = p->vbase[k];
        const int npairs = d * (d + 1) / 2;
        for (int mp = 0; mp < d; mp++)
            for (int m = 0; m <= mp; m++) {
                const long long slot = cb + (long long)mp * (mp + 1) / 2 + m;
                const int pa = aj[m], pb = aj[mp];
                if (pa < 0 || pb < 0) continue;
                add_block(prob, pa, pb, slot, true);
                if (m != mp && pa == pb) bl.push_back({((long long)prob << 48) | ((long long)pa << 24) | pa, (int)(slot * 2 + 1)});  // M + M^T
            }
        if (ai >= 0) {
            for (int mp = 0; mp < d; mp++)
                if (aj[mp] >= 0) add_block(prob, ai, aj[mp], cb + npairs + mp, true);
            add_block(prob, ai, ai, cb + npairs + d, false);
        }
        for (int m = 0; m < d; m++)
            if (aj[m] >= 0) vl.push_back({((long long)prob << 32) | aj[m], (int)(vb + m)});
        if (ai >= 0) vl.push_back({((long long)prob << 32) | ai, (int)(vb + d)});
    }
    auto by_key = [](const BSrc &x, const BSrc &y) { return x.key != y.key ? x.key < y.key : x.src < y.src; };
    std::sort(bl.begin(), bl.end(), by_key);
    std::sort(vl.begin(), vl.end(), by_key);
    p->bdst_off.clear(), p->bdst_ld.clear(), p->bsrc_ptr.assign(1, 0), p->bsrc.clear();
    for (size_t q = 0; q < bl.size(); q++) {
        if (q == 0 || bl[q].key != bl[q - 1].key) {
            if (q) p->bsrc_ptr.push_back((int)p->bsrc.size());
            const int prob = (int)(bl[q].key >> 48), hi = (int)((bl[q].key >> 24) & 0xFFFFFF), lo = (int)(bl[q].key & 0xFFFFFF);
            const long long n = p->prob_npad[prob];
            p->bdst_off.push_back(p->prob_hoff[prob] + (long long)(6 * hi) * n + 6 * lo);
            p->bdst_ld.push_back((int)n);
        }
        p->bsrc.push_back(bl[q].src);
    }
    if (!bl.empty()) p->bsrc_ptr.push_back((int)p->bsrc.size());
    p->vdst_off.clear(), p->vdst_adiag.clear(), p->vsrc_ptr.assign(1, 0), p->vsrc.clear();
    for (size_t q = 0; q < vl.size(); q++) {
        if (q == 0 || vl[q].key != vl[q - 1].key) {
            if (q) p->vsrc_ptr.push_back((int)p->vsrc.size());
            const int prob = (int)(vl[q].key >> 32), pose = (int)(vl[q].key & 0xFFFFFFFF);
            const long long n = p->prob_npad[prob];
            p->vdst_off.push_back(p->prob_hoff[prob] + n * n + 6 * pose);
            p->vdst_adiag.push_back((int)n);
        }
        p->vsrc.push_back(vl[q].src);
    }
    if (!vl.empty()) p->vsrc_ptr.push_back((int)p->vsrc.size());
}

// C independent problems share one plan: problem c owns frames [frame_ptr[c], frame_ptr[c+1]) and optimises the poses
// of its window [t0s[c], t1s[c]) (global frame ids).  C == 1 is the reference operator.
static int plan_create_impl(const int64_t *ii, const int64_t *jj, int64_t n_edges, int64_t n_frames, int ht, int wd, int C,
                            const int64_t *frame_ptr, const int64_t *t0s, const int64_t *t1s, int rank, int world,
                            vipe_ba_plan **out) {
    if (!out) return fail("out is null");
    *out = nullptr;
    if (n_edges < 0 || n_frames <= 0 || ht <= 0 || wd <= 0 || C < 1) return fail("bad sizes");
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank/world");
    if (C > 1 && world != 1) return fail("batched plans are not sharded: give every rank its own clips");
    if (n_edges > 0 && (!ii || !jj)) return fail("ii/jj null");
    if (frame_ptr[0] != 0 || frame_ptr[C] != n_frames) return fail("frame_ptr must run from 0 to n_frames");
    std::vector<int> prob_of_frame(n_frames, -1);
    for (int c = 0; c < C; c++) {
        if (frame_ptr[c + 1] < frame_ptr[c]) return fail("frame_ptr must be non-decreasing");
        if (t0s[c] < frame_ptr[c] || t1s[c] < t0s[c] || t1s[c] > frame_ptr[c + 1])
            return fail("need frame_ptr[c] <= t0 <= t1 <= frame_ptr[c+1] for every problem");
        for (int64_t f = frame_ptr[c]; f < frame_ptr[c + 1]; f++) prob_of_frame[f] = c;
    }
    for (int64_t e = 0; e < n_edges; e++) {
        if (ii[e] < 0 || ii[e] >= n_frames || jj[e] < 0 || jj[e] >= n_frames)
            return fail("edge " + std::to_string(e) + " references a frame outside [0, n_frames)");
        if (prob_of_frame[ii[e]] != prob_of_frame[jj[e]]) return fail("edge " + std::to_string(e) + " connects two different problems");
    }
    const int t0 = (int)t0s[0], t1 = (int)t1s[0];
    auto *p = new vipe_ba_plan();
    p->C = C;
    p->E = n_edges;
    p->N = n_frames;
    p->ht = ht;
    p->wd = wd;
    p->HW = ht * wd;
    p->t0 = t0;
    p->t1 = t1;
    p->rank = rank;
    p->world = world;
    // per-problem reduced systems
    p->pose_slot.assign(n_frames, -1);
    p->pose_row.assign(n_frames, -1);
    p->prob_n.resize(C);
    p->prob_npad.resize(C);
    p->prob_row0.resize(C);
    p->prob_hoff.resize(C);
    {
        int row = 0;
        long long hoff = 0;
        for (int c = 0; c < C; c++) {
            const int Pc = (int)(t1s[c] - t0s[c]);
            p->prob_row0[c] = row;
            p->prob_n[c] = 6 * Pc;
            // a single problem keeps one spare row for the focal-length variable (Options::optimize_focal)
            const int nmax = 6 * Pc + (C == 1 ? 1 : 0);
            p->prob_npad[c] = std::max(kCholBlock, (nmax + kCholBlock - 1) / kCholBlock * kCholBlock);
            p->prob_hoff[c] = hoff;
            if (p->prob_npad[c] > p->prob_n[c]) p->any_padding = true;
            if (C > 1 && p->prob_npad[c] > 2 * kCholBlock) {
                delete p;
                return fail("batched plans need small problems: 6 * (t1 - t0) <= 128 per problem");
            }
            hoff += (long long)p->prob_npad[c] * p->prob_npad[c] + 2LL * p->prob_npad[c];
            for (int64_t f = t0s[c]; f < t1s[c]; f++) {
                p->pose_slot[f] = (int)(f - t0s[c]);
                p->pose_row[f] = row + (int)(f - t0s[c]);
            }
            row += Pc;
        }
        p->P = row;
        p->sys_doubles = (size_t)hoff;
    }
    p->n = p->prob_n[0];
    p->npad = p->prob_npad[0];
    const int P = p->P;
    const int64_t E = n_edges;

    // kx = sorted unique of cat(arange(t0,t1), ii);  kk_exp = inverse   (geom_kernels.cu:1301-1308)
    std::vector<char> present(n_frames, 0);
    for (int c = 0; c < C; c++)
        for (int64_t t = t0s[c]; t < t1s[c]; t++) present[t] = 1;
    for (int64_t e = 0; e < E; e++) present[ii[e]] = 1;
    std::vector<int> slot(n_frames, -1);
    for (int64_t f = 0; f < n_frames; f++)
        if (present[f]) {
            slot[f] = (int)p->kx.size();
            p->kx.push_back(f);
        }
    p->K = (int)p->kx.size();
    p->kk_exp.resize(P + E);
    for (int64_t f = 0; f < n_frames; f++)
        if (p->pose_row[f] >= 0) p->kk_exp[p->pose_row[f]] = slot[f];
    for (int64_t e = 0; e < E; e++) p->kk_exp[P + e] = slot[ii[e]];
    p->frame_prob.resize(p->K);
    for (int k = 0; k < p->K; k++) p->frame_prob[k] = prob_of_frame[p->kx[k]];

    // CSR of edges by source frame, ascending edge id inside a frame (accum_cuda's ptrs/idxs, :946-981)
    const int K = p->K;
    p->fptr.assign(K + 1, 0);
    for (int64_t e = 0; e < E; e++) p->fptr[slot[ii[e]] + 1]++;
    for (int k = 0; k < K; k++) p->fptr[k + 1] += p->fptr[k];
    p->fedge.resize(E);
    {
        std::vector<int> cur(p->fptr.begin(), p->fptr.end() - 1);
        for (int64_t e = 0; e < E; e++) p->fedge[cur[slot[ii[e]]]++] = (int)e;
    }
    p->e_jj.resize(E);
    for (int64_t e = 0; e < E; e++) p->e_jj[e] = (int)jj[e];
    p->kx32.resize(K);
    for (int k = 0; k < K; k++) p->kx32[k] = (int)p->kx[k];

    // number of schur_block triples (:1209-1240): per frame k, (rows whose target pose is in the window)^2
    for (int k = 0; k < K; k++) {
        int64_t rows = 0;
        const int f = (int)p->kx[k];
        if (p->pose_slot[f] >= 0) rows++;  // the Ei row of pose f
        for (int s = p->fptr[k]; s < p->fptr[k + 1]; s++) {
            const int j = p->e_jj[p->fedge[s]];
            if (p->pose_slot[j] >= 0) rows++;
        }
        p->n_triples += rows * rows;
    }

    // keyframe sharding: contiguous kx ranges balanced by (out-degree + 1)
    p->own_lo.assign(world, 0);
    p->own_hi.assign(world, 0);
    {
        std::vector<int64_t> cost(K + 1, 0);
        for (int k = 0; k < K; k++) cost[k + 1] = cost[k] + (p->fptr[k + 1] - p->fptr[k]) + 1;
        int k = 0;
        for (int r = 0; r < world; r++) {
            p->own_lo[r] = k;
            const int64_t target = cost[K] * (r + 1) / world;
            while (k < K && cost[k + 1] <= target) k++;
            if (r == world - 1) k = K;
            p->own_hi[r] = k;
        }
    }
    p->k_lo = (int)p->own_lo[rank];
    p->k_hi = (int)p->own_hi[rank];
    p->dmax = 0;
    for (int k = p->k_lo; k < p->k_hi; k++) p->dmax = std::max(p->dmax, p->fptr[k + 1] - p->fptr[k]);

    compute_elimination_order(p);
    build_assembly_lists(p);

    // Tile shape from the out-degrees.  Up to kHubDegree edges per source frame it follows the largest degree, as it always
    // did.  Beyond that the tile follows kHubDegree and the frames whose staging buffer would not fit shared memory ("hub"
    // frames; the reference has no out-degree limit) run in a second launch that stages in global memory -- a single hub no
    // longer forces the narrowest tile on every frame, and no degree makes plan creation fail on the packed path.
    constexpr int kHubDegree = 32;
    int tile_degree = std::min(p->dmax, kHubDegree);  // the degree the tile shape is chosen for
    if (p->HW % 2 == 0 && p->dmax > 1) {
        // The widest tile (512 pixels: one partial record per 512 pixels, the least traffic for frame_reduce) holds up to d512
        // edges.  If all but a few frames fit it, they get it and the few run as hub frames (C4: 996 of 1000 frames have <= 16
        // edges, the largest 18 -- which would put every frame on 128-pixel tiles, 4x the Gram partials).
        int d512 = 0;
        while (lin2_staging_bytes(d512 + 1, 256) <= lin2_staging_cap()) d512++;
        if (p->dmax > d512 && d512 >= 1) {
            int64_t e_fit = 0, e_all = 0;
            for (int k = p->k_lo; k < p->k_hi; k++) {
                const int d = p->fptr[k + 1] - p->fptr[k];
                e_all += d;
                if (d <= d512) e_fit += d;
            }
            if (e_fit * 10 >= e_all * 9) tile_degree = d512;
        }
    }
    if (tile_config2(p->HW, std::max(tile_degree, 1), false, p->NT)) {
        p->packed = true;
        p->PPT = 2;
        if (p->dmax > tile_degree) {
            for (int k = p->k_lo; k < p->k_hi; k++) {
                const int d = p->fptr[k + 1] - p->fptr[k];
                if (lin2_staging_bytes(d, p->NT) > lin2_staging_cap()) {
                    p->flist_hub.push_back(k);
                } else {
                    p->flist_bulk.push_back(k);
                    p->dmax_bulk = std::max(p->dmax_bulk, d);
                }
            }
            if (p->flist_hub.empty()) p->flist_bulk.clear();  // everything fits: one launch
        }
    } else if (!tile_config(p->HW, std::max(p->dmax, 1), false, p->NT, p->PPT)) {
        delete p;
        return fail("a source frame has too many outgoing edges for the shared-memory staging buffer (odd ht*wd: scalar kernels)");
    }
    {
        // VIPE_BA_LIN4: "1" = the two-kernel Blackwell pipeline of ba_lin4.cu (disparity blocks first, then asynchronously
        // fed J warps + tcgen05 Gram in tensor memory) whenever every owned frame has 1..kLin4MaxDeg edges; "0" (default) = the
        // frame-major FMA kernel, which is still the faster of the two on B200.  The pipeline works on
        // 256-pixel chunks, so the partial layout of the whole plan follows.
        const char *env4 = std::getenv("VIPE_BA_LIN4");
        const int mode4 = env4 ? std::atoi(env4) : 0;
        int dmin = 1 << 30;
        for (int k = p->k_lo; k < p->k_hi; k++) dmin = std::min(dmin, p->fptr[k + 1] - p->fptr[k]);
        p->use_lin4 = mode4 >= 1 && p->k_hi > p->k_lo && dmin >= 1 && p->HW % 4 == 0 && lin4_supported(p->HW, p->dmax);
        if (p->use_lin4) {
            p->packed = true;
            p->NT = kLin4ItemPx / 2;
            p->PPT = 2;
        }
    }
    p->ntile = (p->HW + p->NT * p->PPT - 1) / (p->NT * p->PPT);
    {
        // VIPE_BA_LIN3: "0" = frame-major FMA kernels only; "1" (default) = the TMA-fed pipeline for motion-only runs
        const char *env = std::getenv("VIPE_BA_LIN3");
        const int mode = env ? std::atoi(env) : 1;
        p->use_lin3 = mode >= 1 && p->packed;  // (the motion-only tile shape is fixed further down; checked there)
    }
    // motion-only shares the partial layout (ntile), so it uses the same tile shape
    p->NTm = p->NT;
    p->PPTm = p->PPT;
    p->ntile_m = p->ntile;
    if (p->use_lin4 && p->HW % 512 == 0) {
        // motion-only runs keep their own, wide tile (they have no staging buffer): the partial layout is chosen per call
        p->NTm = 256;
        p->PPTm = 2;
        p->ntile_m = p->HW / 512;
    }
    p->use_lin3 = p->use_lin3 && lin3_supported(p->HW, p->NTm * p->PPTm);

    // partial-buffer offsets
    p->gbase.assign(K + 1, 0);
    p->mbase.assign(K + 1, 0);
    for (int k = 0; k < K; k++) {
        const long long d = p->fptr[k + 1] - p->fptr[k];
        const bool owned = k >= p->k_lo && k < p->k_hi;
        const long long rec = owned ? d * (d + 1) / 2 * 36 + 6 * d : 0;
        p->gbase[k + 1] = p->gbase[k] + rec * p->ntile;
        p->mbase[k + 1] = p->mbase[k] + (owned ? d * (d + 1) / 2 * 36 : 0);
    }

    // workspace layout
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes);
        return o;
    };
    p->off_kx = take(sizeof(int) * K);
    p->off_fptr = take(sizeof(int) * (K + 1));
    p->off_fedge = take(sizeof(int) * std::max<int64_t>(E, 1));
    p->off_ejj = take(sizeof(int) * std::max<int64_t>(E, 1));
    p->off_gbase = take(sizeof(long long) * (K + 1));
    p->off_mbase = take(sizeof(long long) * (K + 1));
    p->off_pslot = take(sizeof(int) * n_frames);
    p->off_prow = take(sizeof(int) * n_frames);
    p->off_fprob = take(sizeof(int) * K);
    p->off_phoff = take(sizeof(long long) * C);
    p->off_pnpad = take(sizeof(int) * C);
    p->off_pn = take(sizeof(int) * C);
    p->off_prow0 = take(sizeof(int) * C);
    p->off_pn_focal = take(sizeof(int) * C);
    if (p->use_lin4) {
        // items: frames by descending degree (the contiguous CTA ranges below then hold items of similar cost), chunk-major inside
        std::vector<int> order;
        for (int k = p->k_lo; k < p->k_hi; k++) order.push_back(k);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return p->fptr[x + 1] - p->fptr[x] > p->fptr[y + 1] - p->fptr[y]; });
        const int nchunk = p->HW / kLin4ItemPx, nsub = kLin4ItemPx / 64;
        std::vector<long long> ucum(1, 0);
        for (int k : order) {
            const int s0 = p->fptr[k], d = p->fptr[k + 1] - s0, src = p->kx32[k];
            for (int ch = 0; ch < nchunk; ch++) {
                const int item = (int)p->lin4_items.size();
                p->lin4_items.push_back(Lin4Item{k, d, ch, 0});
                for (int t = 0; t < nsub; t++)
                    for (int e = 0; e < d; e++) {
                        const int px0 = ch * kLin4ItemPx + t * 64;
                        p->lin4_units.push_back(Lin4Unit{s0 + e, p->fedge[s0 + e], src * p->HW + px0, k * p->HW + px0, e | (d << 8),
                                                          ((px0 / p->wd) << 16) | (px0 % p->wd), px0, item * nsub + t});
                    }
                ucum.push_back((long long)p->lin4_units.size());
            }
        }
        const int nitems = (int)p->lin4_items.size();
        p->lin4_grid = std::min(device_sm_count(), nitems);
        const long long total_units = ucum.back();
        p->lin4_cta_item.assign(p->lin4_grid + 1, 0);
        p->lin4_cta_unit.assign(p->lin4_grid + 1, 0);
        int it = 0;
        for (int b = 1; b <= p->lin4_grid; b++) {
            const long long target = total_units * b / p->lin4_grid;
            const int lo = it;  // at least one item per CTA, and enough left for the CTAs behind
            while (it < nitems - (p->lin4_grid - b) && (it == lo || ucum[it + 1] <= target)) it++;
            if (b == p->lin4_grid) it = nitems;
            p->lin4_cta_item[b] = it;
            p->lin4_cta_unit[b] = (int)ucum[it];
        }
    }
    p->off_lin4_items = take(sizeof(Lin4Item) * std::max<size_t>(p->lin4_items.size(), 1));
    p->off_lin4_units = take(sizeof(Lin4Unit) * std::max<size_t>(p->lin4_units.size(), 1));
    p->off_lin4_cta_item = take(sizeof(int) * std::max<size_t>(p->lin4_cta_item.size(), 1));
    p->off_lin4_cta_unit = take(sizeof(int) * std::max<size_t>(p->lin4_cta_unit.size(), 1));
    p->off_flist_hub = take(sizeof(int) * std::max<size_t>(p->flist_hub.size(), 1));
    p->off_flist_bulk = take(sizeof(int) * std::max<size_t>(p->flist_bulk.size(), 1));
    p->slot_src.assign(std::max<int64_t>(E, 1), 0);
    for (int k = 0; k < K; k++)
        for (int s2 = p->fptr[k]; s2 < p->fptr[k + 1]; s2++) p->slot_src[s2] = p->kx32[k];
    p->off_slot_src = take(sizeof(int) * p->slot_src.size());
    p->off_psys = take(sizeof(int) * n_frames);
    p->off_cbase = take(sizeof(long long) * (K + 1));
    p->off_vbase = take(sizeof(long long) * (K + 1));
    p->off_bdst_off = take(sizeof(long long) * std::max<size_t>(p->bdst_off.size(), 1));
    p->off_bdst_ld = take(sizeof(int) * std::max<size_t>(p->bdst_ld.size(), 1));
    p->off_bsrc_ptr = take(sizeof(int) * std::max<size_t>(p->bsrc_ptr.size(), 1));
    p->off_bsrc = take(sizeof(int) * std::max<size_t>(p->bsrc.size(), 1));
    p->off_vdst_off = take(sizeof(long long) * std::max<size_t>(p->vdst_off.size(), 1));
    p->off_vdst_adiag = take(sizeof(int) * std::max<size_t>(p->vdst_adiag.size(), 1));
    p->off_vsrc_ptr = take(sizeof(int) * std::max<size_t>(p->vsrc_ptr.size(), 1));
    p->off_vsrc = take(sizeof(int) * std::max<size_t>(p->vsrc.size(), 1));
    p->off_rowmap = take(sizeof(int) * p->npad);
    p->off_tstruct = take(std::max<size_t>(p->tstruct.size(), 1));
    p->idx_bytes = off;
    p->ntile_f = focal_tiles(p->HW);
    p->off_econst = take(sizeof(float) * 16 * (p->use_lin3 ? (size_t)std::max<int64_t>(E, 1) : 1));
    p->off_fpart = take(sizeof(float) * (size_t)std::max<int64_t>(E, 1) * p->ntile_f * kFocalStride);
    p->off_ffpart = take(sizeof(float) * (size_t)K * p->ntile_f * 2);
    p->off_uf = take(sizeof(float) * (size_t)K * p->HW);
    p->off_econst2 = take(sizeof(float) * 32 * (p->use_lin4 ? (size_t)std::max<int64_t>(E, 1) : 1));
    p->off_sq = take(sizeof(float) * (p->use_lin4 ? (size_t)K * p->HW : 1));
    p->off_sqw = take(sizeof(float) * (p->use_lin4 ? (size_t)K * p->HW : 1));
    p->off_epart = take(sizeof(float) * (size_t)std::max<int64_t>(E, 1) * std::max(p->use_lin4 ? std::max(p->ntile, p->HW / 64) : p->ntile, p->ntile_m) * kEdgeStride);
    {
        const size_t own_slots = (size_t)(p->fptr[p->k_hi] - p->fptr[p->k_lo]);
        p->off_uglobal = take(sizeof(float) * (p->flist_hub.empty() ? 1 : own_slots * p->ntile * 6 * (size_t)(p->NT * p->PPT)));
        // frame_reduce keeps 126 doubles per edge; beyond ~240 edges per frame that no longer fits shared memory
        p->off_fscratch = take(sizeof(double) * (p->dmax > 200 ? own_slots * kReduceDoubles : 1));
    }
    p->off_gpart = take(sizeof(float) * (size_t)std::max<long long>(p->gbase[K], 1));
    p->off_msc = take(sizeof(double) * (size_t)std::max<long long>(p->mbase[K], 1));
    p->off_q = take(sizeof(float) * (size_t)K * p->HW);
    p->off_qw = take(sizeof(float) * (size_t)K * p->HW);
    p->off_sys = take(sizeof(double) * p->sys_doubles);  // per problem [H ; b ; diag(A)]
    p->off_dx = take(sizeof(double) * ((size_t)p->npad + (size_t)p->npad * kCholBlock));  // 1/diag(L), then L_jj^-T tiles
    p->off_cblk = take(sizeof(double) * 36 * (size_t)std::max<long long>(p->n_cblk, 1));
    p->off_cvec = take(sizeof(double) * 6 * (size_t)std::max<long long>(p->n_cvec, 1));
    p->off_cvec2 = take(sizeof(double) * 6 * (size_t)std::max<long long>(p->n_cvec, 1));
    p->off_gdx = take(sizeof(float) * ((size_t)6 * std::max(P, 1) + 1));  // graph-owned dx / dz (see vipe_ba_run)
    p->off_gdz = take(sizeof(float) * (size_t)std::max(K, 1) * p->HW);
    p->off_flag = take(sizeof(int) * chol_scratch_ints(p->npad));
    p->flag_bytes = sizeof(int) * chol_scratch_ints(p->npad);
    p->total = off;

    p->blob.assign(p->idx_bytes, 0);
    std::memcpy(p->blob.data() + p->off_kx, p->kx32.data(), sizeof(int) * K);
    std::memcpy(p->blob.data() + p->off_fptr, p->fptr.data(), sizeof(int) * (K + 1));
    if (E > 0) {
        std::memcpy(p->blob.data() + p->off_fedge, p->fedge.data(), sizeof(int) * E);
        std::memcpy(p->blob.data() + p->off_ejj, p->e_jj.data(), sizeof(int) * E);
    }
    std::memcpy(p->blob.data() + p->off_gbase, p->gbase.data(), sizeof(long long) * (K + 1));
    std::memcpy(p->blob.data() + p->off_mbase, p->mbase.data(), sizeof(long long) * (K + 1));
    std::memcpy(p->blob.data() + p->off_pslot, p->pose_slot.data(), sizeof(int) * n_frames);
    std::memcpy(p->blob.data() + p->off_prow, p->pose_row.data(), sizeof(int) * n_frames);
    std::memcpy(p->blob.data() + p->off_fprob, p->frame_prob.data(), sizeof(int) * K);
    std::memcpy(p->blob.data() + p->off_phoff, p->prob_hoff.data(), sizeof(long long) * C);
    std::memcpy(p->blob.data() + p->off_pnpad, p->prob_npad.data(), sizeof(int) * C);
    std::memcpy(p->blob.data() + p->off_pn, p->prob_n.data(), sizeof(int) * C);
    std::memcpy(p->blob.data() + p->off_prow0, p->prob_row0.data(), sizeof(int) * C);
    std::memcpy(p->blob.data() + p->off_psys, p->pose_sys.data(), sizeof(int) * n_frames);
    auto put = [&](size_t off, const void *src, size_t bytes) {
        if (bytes) std::memcpy(p->blob.data() + off, src, bytes);
    };
    put(p->off_cbase, p->cbase.data(), sizeof(long long) * p->cbase.size());
    put(p->off_vbase, p->vbase.data(), sizeof(long long) * p->vbase.size());
    put(p->off_bdst_off, p->bdst_off.data(), sizeof(long long) * p->bdst_off.size());
    put(p->off_bdst_ld, p->bdst_ld.data(), sizeof(int) * p->bdst_ld.size());
    put(p->off_bsrc_ptr, p->bsrc_ptr.data(), sizeof(int) * p->bsrc_ptr.size());
    put(p->off_bsrc, p->bsrc.data(), sizeof(int) * p->bsrc.size());
    put(p->off_vdst_off, p->vdst_off.data(), sizeof(long long) * p->vdst_off.size());
    put(p->off_vdst_adiag, p->vdst_adiag.data(), sizeof(int) * p->vdst_adiag.size());
    put(p->off_vsrc_ptr, p->vsrc_ptr.data(), sizeof(int) * p->vsrc_ptr.size());
    put(p->off_vsrc, p->vsrc.data(), sizeof(int) * p->vsrc.size());
    std::memcpy(p->blob.data() + p->off_rowmap, p->rowmap.data(), sizeof(int) * p->npad);
    if (!p->tstruct.empty()) std::memcpy(p->blob.data() + p->off_tstruct, p->tstruct.data(), p->tstruct.size());
    std::memcpy(p->blob.data() + p->off_slot_src, p->slot_src.data(), sizeof(int) * p->slot_src.size());
    put(p->off_flist_hub, p->flist_hub.data(), sizeof(int) * p->flist_hub.size());
    put(p->off_flist_bulk, p->flist_bulk.data(), sizeof(int) * p->flist_bulk.size());
    put(p->off_lin4_items, p->lin4_items.data(), sizeof(Lin4Item) * p->lin4_items.size());
    put(p->off_lin4_units, p->lin4_units.data(), sizeof(Lin4Unit) * p->lin4_units.size());
    put(p->off_lin4_cta_item, p->lin4_cta_item.data(), sizeof(int) * p->lin4_cta_item.size());
    put(p->off_lin4_cta_unit, p->lin4_cta_unit.data(), sizeof(int) * p->lin4_cta_unit.size());
    {
        std::vector<int> nf(p->prob_n);
        for (auto &v : nf) v += 1;
        std::memcpy(p->blob.data() + p->off_pn_focal, nf.data(), sizeof(int) * C);
    }
    *out = p;
    return 0;
}

extern "C" int vipe_ba_plan_create(const int64_t *ii, const int64_t *jj, int64_t n_edges, int64_t n_frames, int ht,
                                   int wd, int t0, int t1, int rank, int world, vipe_ba_plan **out) {
    if (t0 < 0 || t1 < t0 || t1 > n_frames) return fail("need 0 <= t0 <= t1 <= n_frames");
    const int64_t fp[2] = {0, n_frames}, a0[1] = {t0}, a1[1] = {t1};
    return plan_create_impl(ii, jj, n_edges, n_frames, ht, wd, 1, fp, a0, a1, rank, world, out);
}

extern "C" int vipe_ba_plan_create_batch(const int64_t *ii, const int64_t *jj, int64_t n_edges, int64_t n_frames, int ht,
                                         int wd, int n_problems, const int64_t *frame_ptr, const int64_t *t0s,
                                         const int64_t *t1s, vipe_ba_plan **out) {
    if (n_problems < 1 || !frame_ptr || !t0s || !t1s) return fail("bad batch description");
    return plan_create_impl(ii, jj, n_edges, n_frames, ht, wd, n_problems, frame_ptr, t0s, t1s, 0, 1, out);
}
extern "C" int64_t vipe_ba_plan_num_free_poses(const vipe_ba_plan *p) { return p ? p->P : -1; }

extern "C" void vipe_ba_plan_destroy(vipe_ba_plan *plan) { delete plan; }
extern "C" int64_t vipe_ba_plan_num_kx(const vipe_ba_plan *p) { return p ? p->K : -1; }
extern "C" int vipe_ba_plan_copy_kx(const vipe_ba_plan *p, int64_t *o) {
    if (!p || !o) return fail("null argument");
    std::copy(p->kx.begin(), p->kx.end(), o);
    return 0;
}
extern "C" int vipe_ba_plan_copy_kk_exp(const vipe_ba_plan *p, int64_t *o) {
    if (!p || !o) return fail("null argument");
    std::copy(p->kk_exp.begin(), p->kk_exp.end(), o);
    return 0;
}
extern "C" int vipe_ba_plan_copy_csr(const vipe_ba_plan *p, int64_t *ptrs, int64_t *idxs) {
    if (!p || !ptrs || !idxs) return fail("null argument");
    for (int k = 0; k <= p->K; k++) ptrs[k] = p->fptr[k];
    for (int64_t e = 0; e < p->E; e++) idxs[e] = p->fedge[e];
    return 0;
}
extern "C" int vipe_ba_plan_owned_range(const vipe_ba_plan *p, int rank, int64_t *lo, int64_t *hi) {
    if (!p || !lo || !hi || rank < 0 || rank >= p->world) return fail("bad argument");
    *lo = p->own_lo[rank];
    *hi = p->own_hi[rank];
    return 0;
}
extern "C" int vipe_ba_plan_copy_sys_order(const vipe_ba_plan *p, int64_t *o) {
    if (!p || !o) return fail("null argument");
    for (int64_t f = 0; f < p->N; f++)
        if (p->pose_row[f] >= 0) o[p->pose_row[f]] = p->pose_sys[f];
    return 0;
}
extern "C" int64_t vipe_ba_plan_num_schur_triples(const vipe_ba_plan *p) { return p ? p->n_triples : -1; }
extern "C" int vipe_ba_plan_max_degree(const vipe_ba_plan *p) { return p ? p->dmax : -1; }
extern "C" size_t vipe_ba_workspace_bytes(const vipe_ba_plan *p) { return p ? p->total : 0; }
extern "C" int64_t vipe_ba_launch_count(const vipe_ba_plan *p) { return p ? p->launches : -1; }

extern "C" int vipe_ba_plan_upload(const vipe_ba_plan *p, void *ws, void *stream) {
    if (!p || !ws) return fail("null argument");
    VBA_CUDA(cudaMemcpyAsync(ws, p->blob.data(), p->idx_bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    VBA_CUDA(cudaMemsetAsync((unsigned char *)ws + p->off_flag, 0, p->flag_bytes, (cudaStream_t)stream));
    return 0;
}

static Tables make_tables(const vipe_ba_plan *p, void *ws) {
    unsigned char *w = (unsigned char *)ws;
    Tables tb;
    tb.kx = (const int *)(w + p->off_kx);
    tb.fptr = (const int *)(w + p->off_fptr);
    tb.fedge = (const int *)(w + p->off_fedge);
    tb.e_jj = (const int *)(w + p->off_ejj);
    tb.gbase = (const long long *)(w + p->off_gbase);
    tb.mbase = (const long long *)(w + p->off_mbase);
    tb.K = p->K;
    tb.E = (int)p->E;
    tb.N = (int)p->N;
    tb.HW = p->HW;
    tb.wd = p->wd;
    tb.pose_slot = (const int *)(w + p->off_pslot);
    tb.pose_row = (const int *)(w + p->off_prow);
    tb.pose_sys = (const int *)(w + p->off_psys);
    tb.frame_prob = (const int *)(w + p->off_fprob);
    tb.prob_hoff = (const long long *)(w + p->off_phoff);
    tb.prob_npad = (const int *)(w + p->off_pnpad);
    tb.prob_n = (const int *)(w + (p->opt.optimize_focal ? p->off_pn_focal : p->off_pn));
    tb.prob_row0 = (const int *)(w + p->off_prow0);
    tb.C = p->C;
    tb.ntile = p->ntile;
    tb.ntile_e = p->ntile;
    tb.k_lo = p->k_lo;
    tb.k_hi = p->k_hi;
    tb.rows_by_slot = p->rows_by_slot;
    tb.slot_lo = p->fptr[p->k_lo];
    return tb;
}

extern "C" void *vipe_ba_system_buffer(const vipe_ba_plan *p, void *ws, int64_t *n_out, int64_t *count_out) {
    if (!p || !ws) return nullptr;
    if (n_out) *n_out = p->npad;
    if (count_out) *count_out = (int64_t)p->npad * p->npad + 2 * (int64_t)p->npad;
    return (unsigned char *)ws + p->off_sys;
}
extern "C" float *vipe_ba_debug_q(const vipe_ba_plan *p, void *ws) { return (float *)((unsigned char *)ws + p->off_q); }
extern "C" float *vipe_ba_debug_qw(const vipe_ba_plan *p, void *ws) { return (float *)((unsigned char *)ws + p->off_qw); }

static int check_tensors(const vipe_ba_plan *p, const vipe_ba_tensors *t, int motion_only) {
    if (!p || !t) return fail("null plan/tensors");
    if (!t->poses || !t->disps || !t->intrinsics || !t->dx_out) return fail("null tensor pointer");
    // a graph without edges has empty targets/weights, whose storage pointer is null
    const bool no_rows = p->rows_by_slot && p->fptr[p->k_hi] == p->fptr[p->k_lo];  // a shard that owns no edge
    if (p->E > 0 && !no_rows && (!t->targets || !t->weights)) return fail("null tensor pointer");
    if (!motion_only && (!t->disps_sens || !t->eta || !t->dz_out)) return fail("disps_sens/eta/dz_out required unless motion_only");
    // the kernels read and write the pixel arrays with 8/16-byte vector accesses and 16-byte bulk copies whenever ht*wd is
    // even: a misaligned base would be a sticky device fault, so it is refused here (the Python operator re-aligns by copy)
    if (p->HW % 2 == 0) {
        const void *px[] = {t->disps, t->disps_sens, t->targets, t->weights, t->eta, t->dz_out};
        for (const void *q : px)
            if (q && (reinterpret_cast<uintptr_t>(q) & 15u)) return fail("pixel arrays (disps, disps_sens, targets, weights, eta, dz_out) must be 16-byte aligned");
    }
    return 0;
}

static int linearize_impl(const vipe_ba_plan *p, const vipe_ba_tensors *t, void *ws, int motion_only, cudaStream_t st,
                          cudaEvent_t mid, cudaEvent_t after_clear = nullptr) {
    if (check_tensors(p, t, motion_only)) return 1;
    if (!ws) return fail("null workspace");
    unsigned char *w = (unsigned char *)ws;
    if (p->dist_factor) {
        // Distributed solve: what the owners multicast validates itself against ZERO words, so this rank's receiving copies
        // ([L ; y] and the L_jj^-T tiles) are cleared here -- before the cross-rank barrier that precedes the solve, so that no
        // rank can be writing into them yet, and after this rank's own reads of the previous solve (stream order).
        const size_t soff = align_up(chol_scratch_ints(p->npad) * sizeof(int));
        VBA_CUDA(cudaMemsetAsync(p->dist_factor, 0, sizeof(double) * ((size_t)p->npad * p->npad + p->npad), st));
        VBA_CUDA(cudaMemsetAsync(p->dist_aux + soff + sizeof(double) * p->npad, 0, sizeof(double) * (size_t)p->npad * kCholBlock, st));
    }
    double *H = p->peer_accum ? p->peer_accum : (double *)(w + p->off_sys);
    {
        const Tables t0b = make_tables(p, ws);
        VBA_CUDA(launch_system_clear(H, p->sys_doubles, t0b.prob_hoff, t0b.prob_n, t0b.prob_npad, p->C, p->any_padding, st));
    }
    p->launches += p->any_padding ? 2 : 1;
    if (after_clear) VBA_CUDA(cudaEventRecord(after_clear, st));  // stage timing starts here: the system clear is not the linearisation
    const int nframes = p->k_hi - p->k_lo;
    if (nframes <= 0 || p->P <= 0) return 0;

    LinArgs la;
    la.tb = make_tables(p, ws);
    la.opt = p->opt;
    la.poses = t->poses;
    la.disps = t->disps;
    la.intr = t->intrinsics;
    la.dsens = t->disps_sens;
    la.targets = t->targets;
    la.weights = t->weights;
    la.eta = t->eta;
    la.epart = (float *)(w + p->off_epart);
    la.gpart = (float *)(w + p->off_gpart);
    la.qbuf = (float *)(w + p->off_q);
    la.qwbuf = (float *)(w + p->off_qw);
    if (motion_only) la.tb.ntile = la.tb.ntile_e = p->ntile_m;  // the motion-only kernels tile a frame their own way
    if (p->use_lin4 && !motion_only) {
        la.tb.ntile_e = p->HW / 64;  // one edge record per 64-pixel unit
        Lin4Launch l4;
        l4.items = (const Lin4Item *)(w + p->off_lin4_items);
        l4.units = (const Lin4Unit *)(w + p->off_lin4_units);
        l4.cta_item = (const int *)(w + p->off_lin4_cta_item);
        l4.cta_unit = (const int *)(w + p->off_lin4_cta_unit);
        l4.grid = p->lin4_grid;
        l4.nframes = nframes;
        l4.dmax = std::max(p->dmax, 1);
        l4.slot_src = (const int *)(w + p->off_slot_src);
        l4.slot_lo = p->fptr[p->k_lo];
        l4.nslots = p->fptr[p->k_hi] - l4.slot_lo;
        l4.econst2 = (float2 *)(w + p->off_econst2);
        l4.sqbuf = (float *)(w + p->off_sq);
        l4.sqwbuf = (float *)(w + p->off_sqw);
        VBA_CUDA(launch_lin4(la, l4, st));
        p->launches += 3;
    } else if (p->packed && motion_only && p->use_lin3) {
        const int slot_lo = p->fptr[p->k_lo], nslots = p->fptr[p->k_hi] - slot_lo;
        VBA_CUDA(launch_lin3_motion(la, (const int *)(w + p->off_slot_src), slot_lo, nslots, p->NTm * p->PPTm, (float *)(w + p->off_econst),
                                    device_sm_count(), st));
        p->launches += 2;
    } else {
        if (p->packed && !motion_only && !p->flist_hub.empty()) {
            la.uglobal = (float *)(w + p->off_uglobal);
            la.flist = (const int *)(w + p->off_flist_hub);  // the long CTAs first
            VBA_CUDA(launch_linearize2(la, (int)p->flist_hub.size(), std::max(p->dmax, 1), false, p->NT, st, true));
            la.flist = (const int *)(w + p->off_flist_bulk);
            VBA_CUDA(launch_linearize2(la, (int)p->flist_bulk.size(), std::max(p->dmax_bulk, 1), false, p->NT, st));
            la.flist = nullptr;
            p->launches++;
        } else if (p->packed)
            VBA_CUDA(launch_linearize2(la, nframes, std::max(p->dmax, 1), motion_only != 0, motion_only ? p->NTm : p->NT, st));
        else
            VBA_CUDA(launch_linearize(la, nframes, std::max(p->dmax, 1), motion_only != 0, p->NT, p->PPT, st));
        p->launches++;
    }
    if (p->opt.optimize_focal) {
        FocalArgs fa;
        fa.tb = la.tb;
        fa.opt = p->opt;
        fa.poses = t->poses;
        fa.disps = t->disps;
        fa.intr = t->intrinsics;
        fa.targets = t->targets;
        fa.weights = t->weights;
        fa.qbuf = la.qbuf;
        fa.qwbuf = la.qwbuf;
        fa.fpart = (float *)(w + p->off_fpart);
        fa.ffpart = (float *)(w + p->off_ffpart);
        fa.ufbuf = (float *)(w + p->off_uf);
        fa.ntile_f = p->ntile_f;
        fa.motion_only = motion_only;
        VBA_CUDA(launch_focal(fa, nframes, std::max(p->dmax, 1), st));
        p->launches++;
    }
    if (mid) VBA_CUDA(cudaEventRecord(mid, st));

    ReduceArgs ra;
    if (p->opt.optimize_focal) {
        ra.fpart = (const float *)(w + p->off_fpart);
        ra.ffpart = (const float *)(w + p->off_ffpart);
        ra.ntile_f = p->ntile_f;
        ra.focal_row = p->n;
        ra.focal_lm = p->opt.focal_lm;
    }
    ra.tb = la.tb;
    ra.poses = t->poses;
    ra.epart = la.epart;
    ra.gpart = la.gpart;
    ra.msc = (double *)(w + p->off_msc);
    ra.hsys = H;
    ra.motion_only = motion_only;
    ra.fscratch = p->dmax > 200 ? (double *)(w + p->off_fscratch) : nullptr;
    ra.cblk = (double *)(w + p->off_cblk);
    ra.cvec = (double *)(w + p->off_cvec);
    ra.cvec2 = (double *)(w + p->off_cvec2);
    ra.cbase = (const long long *)(w + p->off_cbase);
    ra.vbase = (const long long *)(w + p->off_vbase);
    VBA_CUDA(launch_frame_reduce(ra, nframes, std::max(p->dmax, 1), st));
    p->launches++;
    AssembleArgs aa;
    aa.cblk = ra.cblk, aa.cvec = ra.cvec, aa.cvec2 = ra.cvec2;
    aa.sys = H;
    aa.bdst_off = (const long long *)(w + p->off_bdst_off);
    aa.bdst_ld = (const int *)(w + p->off_bdst_ld);
    aa.bsrc_ptr = (const int *)(w + p->off_bsrc_ptr);
    aa.bsrc = (const int *)(w + p->off_bsrc);
    aa.nb = (int)p->bdst_off.size();
    aa.vdst_off = (const long long *)(w + p->off_vdst_off);
    aa.vdst_adiag = (const int *)(w + p->off_vdst_adiag);
    aa.vsrc_ptr = (const int *)(w + p->off_vsrc_ptr);
    aa.vsrc = (const int *)(w + p->off_vsrc);
    aa.nv = (int)p->vdst_off.size();
    if (aa.nb + aa.nv > 0) {
        VBA_CUDA(launch_assemble(aa, st));
        p->launches++;
    }
    return 0;
}

extern "C" int vipe_ba_linearize(const vipe_ba_plan *p, const vipe_ba_tensors *t, void *ws, int motion_only,
                                 void *stream) {
    return linearize_impl(p, t, ws, motion_only, (cudaStream_t)stream, nullptr);
}

static int solve_update_impl(const vipe_ba_plan *p, const vipe_ba_tensors *t, void *ws, float lm, float ep,
                             int motion_only, cudaStream_t st, cudaEvent_t mid) {
    if (check_tensors(p, t, motion_only)) return 1;
    if (!ws) return fail("null workspace");
    if (p->P <= 0) return 0;
    unsigned char *w = (unsigned char *)ws;
    double *H = p->dist_factor ? p->dist_factor : p->solve_buf ? p->solve_buf : (double *)(w + p->off_sys);
    double *b = H + (size_t)p->npad * p->npad;
    int *scratch = (int *)(w + p->off_flag);
    double *dinv = (double *)(w + p->off_dx);
    CholDist cd;
    const CholDist *dist = nullptr;
    if (p->dist_factor) {
        if (!p->peer_mc) return fail("the distributed solve reads its input through vipe_ba_set_peer_system's multicast address");
        if (p->opt.optimize_focal) return fail("optimize_focal is not supported with the distributed solve");
        const size_t soff = align_up(chol_scratch_ints(p->npad) * sizeof(int));
        scratch = (int *)p->dist_aux;
        dinv = (double *)(p->dist_aux + soff);
        cd.rank = p->dist_rank, cd.world = p->dist_world, cd.epoch = ++p->dist_epoch;
        {
            const char *env = std::getenv("VIPE_BA_DIST_COLBLK");
            // blocks of two columns from 4 GPUs on (measured at C4: 207 vs 210 it/s on 2 GPUs, 300 vs 296 on 8)
            cd.colblk = env ? std::max(1, std::atoi(env)) : (p->dist_world >= 4 ? 2 : 1);
        }
        cd.Hmc = p->dist_factor_mc;
        cd.scratch_mc = (int *)p->dist_aux_mc;
        cd.linvT_mc = (double *)(p->dist_aux_mc + soff) + p->npad;
        dist = &cd;
    }
    int cnt = 0;
    p->epoch++;
    const Tables tbs = make_tables(p, ws);
    const bool focal = p->opt.optimize_focal != 0;
    if (p->C == 1) {
        // the solver adds ep + lm * diag(A) to every diagonal entry; diag(A) is 0 in the focal row (its LM term went in
        // with the assembly), so the focal row ends up with focal_ep
        if (focal) VBA_CUDA(launch_add_scalar(H + (size_t)p->n * p->npad + p->n, (double)p->opt.focal_ep - (double)ep, st));
        if ((p->peer_mc || p->solve_buf) && p->npad <= 2 * kCholBlock)
            return fail("the fused multi-GPU reduction needs the tiled solver (more than 128 unknowns); use the all-reduce path");
        VBA_CUDA(launch_damped_solve(H, b, p->n + (focal ? 1 : 0), p->npad, lm, ep, t->dx_out, scratch, dinv, dinv + p->npad,
                                     p->opt.damp_on_pose_hessian ? b + p->npad : nullptr,
                                     (p->solve_buf && !dist) ? nullptr : p->peer_mc, p->epoch,
                                     p->ordered ? (const unsigned char *)(w + p->off_tstruct) : nullptr, nullptr, st, &cnt, dist));
    } else {  // many small independent problems: one CTA each
        VBA_CUDA(launch_small_solve_batch(H, tbs.prob_hoff, tbs.prob_n, tbs.prob_npad, tbs.prob_row0, p->C, lm, ep, t->dx_out,
                                          p->opt.damp_on_pose_hessian != 0, st, &cnt));
    }
    p->launches += cnt;
    if (mid) VBA_CUDA(cudaEventRecord(mid, st));
    const int nframes = p->k_hi - p->k_lo;
    if (!motion_only && nframes > 0) {
        BackArgs ba;
        ba.tb = make_tables(p, ws);
        ba.opt = p->opt;
        ba.poses = t->poses;
        ba.intr = t->intrinsics;
        ba.weights = t->weights;
        ba.disps = t->disps;
        ba.qbuf = (const float *)(w + p->off_q);
        ba.qwbuf = (const float *)(w + p->off_qw);
        ba.dx = t->dx_out;
        ba.dz_out = t->dz_out;
        if (focal) {
            ba.ufbuf = (const float *)(w + p->off_uf);
            ba.focal_row = p->n;
        }
        VBA_CUDA(launch_backsub(ba, nframes, std::max(p->dmax, 1), st));
        p->launches++;
    }
    // with optimize_focal the intrinsics are a variable too: fx, fy are updated in place
    VBA_CUDA(launch_pose_retr(t->poses, t->dx_out, tbs.pose_row, (int)p->N, p->opt.renorm_quat,
                              focal ? const_cast<float *>(t->intrinsics) : nullptr, p->n, p->opt.focal_jscale, st));
    p->launches++;
    return 0;
}

extern "C" int vipe_ba_solve_update(const vipe_ba_plan *p, const vipe_ba_tensors *t, void *ws, float lm, float ep,
                                    int motion_only, void *stream) {
    return solve_update_impl(p, t, ws, lm, ep, motion_only, (cudaStream_t)stream, nullptr);
}

static int run_iteration(const vipe_ba_plan *p, const vipe_ba_tensors *t, void *ws, float lm, float ep, int motion_only,
                         cudaStream_t st, cudaEvent_t *ev) {
    // stage 0 = the linearisation kernel(s) alone: its start event is recorded after the system clear
    if (linearize_impl(p, t, ws, motion_only, st, ev ? ev[1] : nullptr, ev ? ev[0] : nullptr)) return 1;
    if (ev) VBA_CUDA(cudaEventRecord(ev[2], st));
    if (solve_update_impl(p, t, ws, lm, ep, motion_only, st, ev ? ev[3] : nullptr)) return 1;
    if (ev) VBA_CUDA(cudaEventRecord(ev[4], st));
    return 0;
}

extern "C" int vipe_ba_profile_enable(vipe_ba_plan *p, int on) {
    if (!p) return fail("null plan");
    p->profile = on != 0;
    if (p->profile && p->events.empty()) {
        p->events.resize(5 * kMaxProfIters);
        for (auto &e : p->events) VBA_CUDA(cudaEventCreate(&e));
    }
    p->prof_iters = 0;
    return 0;
}

extern "C" int vipe_ba_profile_read(const vipe_ba_plan *p, float ms_out[4], int *iters_out) {
    if (!p || !ms_out) return fail("null argument");
    for (int s = 0; s < 4; s++) ms_out[s] = 0.0f;
    if (iters_out) *iters_out = p->prof_iters;
    for (int it = 0; it < p->prof_iters; it++)
        for (int s = 0; s < 4; s++) {
            float ms = 0.0f;
            VBA_CUDA(cudaEventElapsedTime(&ms, p->events[5 * it + s], p->events[5 * it + s + 1]));
            ms_out[s] += ms;
        }
    return 0;
}

// the stages of one iteration, with optional event records between them
static int run_iteration(const vipe_ba_plan *p, const vipe_ba_tensors *t, void *ws, float lm, float ep, int motion_only,
                         cudaStream_t st, cudaEvent_t *ev);

static bool same_args(const vipe_ba_plan::GraphEntry &g, const vipe_ba_tensors *t, void *ws, int iterations, float lm,
                      float ep, int motion_only) {
    // dx_out / dz_out are NOT part of the key: a captured run writes its updates into the workspace and they are copied out
    // after the launch, so callers that keep (and therefore re-allocate) their result tensors still replay
    vipe_ba_tensors a = g.t, b = *t;
    a.dx_out = b.dx_out = nullptr;
    a.dz_out = b.dz_out = nullptr;
    return std::memcmp(&a, &b, sizeof(a)) == 0 && g.ws == ws && g.iterations == iterations && g.lm == lm && g.ep == ep &&
           g.motion_only == motion_only;
}

extern "C" int vipe_ba_set_peer_system(vipe_ba_plan *p, double *accum_local, const double *accum_multicast) {
    if (!p) return fail("null plan");
    if ((accum_local == nullptr) != (accum_multicast == nullptr)) return fail("give both pointers or neither");
    if (accum_local && p->C != 1) return fail("batched plans are not sharded");
    if (accum_local && p->opt.optimize_focal) return fail("optimize_focal needs a single-rank plan");
    p->peer_accum = accum_local;
    p->peer_mc = accum_multicast;
    return 0;
}

extern "C" int64_t vipe_ba_dist_aux_bytes(const vipe_ba_plan *p) {
    if (!p) return -1;
    return (int64_t)(align_up(chol_scratch_ints(p->npad) * sizeof(int)) + sizeof(double) * ((size_t)p->npad + (size_t)p->npad * kCholBlock));
}
extern "C" int vipe_ba_set_dist_solve(vipe_ba_plan *p, double *factor_local, double *factor_multicast, void *aux_local,
                                      void *aux_multicast, int rank, int world) {
    if (!p) return fail("null plan");
    if (factor_local && (!factor_multicast || !aux_local || !aux_multicast || world < 2 || rank < 0 || rank >= world))
        return fail("vipe_ba_set_dist_solve: all four buffers and 0 <= rank < world >= 2 are needed");
    if (factor_local && (p->C != 1 || p->npad <= 2 * kCholBlock)) return fail("the distributed solve needs one problem with more than 128 unknowns");
    p->dist_factor = factor_local;
    p->dist_factor_mc = factor_multicast;
    p->dist_aux = (unsigned char *)aux_local;
    p->dist_aux_mc = (unsigned char *)aux_multicast;
    p->dist_rank = rank, p->dist_world = world;
    for (auto &g : p->graphs)  // captured runs hold the old routing
        if (g.exec) cudaGraphExecDestroy(g.exec);
    p->graphs.clear();
    return 0;
}
extern "C" int vipe_ba_set_solve_buffer(vipe_ba_plan *p, double *local) {
    if (!p) return fail("null plan");
    if (local && p->C != 1) return fail("batched plans are not sharded");
    p->solve_buf = local;
    return 0;
}

extern "C" int vipe_ba_set_owned_rows(vipe_ba_plan *p, int on) {
    if (!p) return fail("null plan");
    p->rows_by_slot = on != 0;
    for (auto &g : p->graphs)  // captured graphs have the old row addressing baked in
        if (g.exec) cudaGraphExecDestroy(g.exec);
    p->graphs.clear();
    return 0;
}

extern "C" int64_t vipe_ba_plan_num_owned_edges(const vipe_ba_plan *p) { return p ? p->fptr[p->k_hi] - p->fptr[p->k_lo] : -1; }
extern "C" int vipe_ba_plan_copy_owned_edges(const vipe_ba_plan *p, int64_t *o) {
    if (!p || !o) return fail("null argument");
    for (int s2 = p->fptr[p->k_lo]; s2 < p->fptr[p->k_hi]; s2++) o[s2 - p->fptr[p->k_lo]] = p->fedge[s2];
    return 0;
}

extern "C" int vipe_ba_peer_reduce(const vipe_ba_plan *p, const double *accum_multicast, double *reduced_multicast, int rank,
                                   int world, void *stream) {
    if (!p || !accum_multicast || !reduced_multicast) return fail("null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank/world");
    VBA_CUDA(launch_peer_reduce(accum_multicast, reduced_multicast, p->npad, rank, world, (cudaStream_t)stream));
    p->launches++;
    return 0;
}

extern "C" int vipe_ba_set_graphs(vipe_ba_plan *p, int on) {
    if (!p) return fail("null plan");
    p->use_graphs = on != 0;
    return 0;
}

// a replayed graph leaves the last iteration's updates in the workspace: hand them to the caller's tensors
static int copy_graph_results(const vipe_ba_plan *p, const vipe_ba_tensors *t, void *ws, int motion_only, cudaStream_t st) {
    const unsigned char *w = (const unsigned char *)ws;
    const size_t ndx = (size_t)6 * p->P + (p->opt.optimize_focal ? 1 : 0);
    VBA_CUDA(cudaMemcpyAsync(t->dx_out, w + p->off_gdx, sizeof(float) * ndx, cudaMemcpyDeviceToDevice, st));
    if (!motion_only && t->dz_out)
        VBA_CUDA(cudaMemcpyAsync(t->dz_out, w + p->off_gdz, sizeof(float) * (size_t)p->K * p->HW, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int vipe_ba_run(const vipe_ba_plan *p, const vipe_ba_tensors *t, void *ws, int iterations, float lm,
                           float ep, int motion_only, void *stream) {
    if (!p) return fail("null plan");
    if (p->world != 1) return fail("vipe_ba_run needs a single-rank plan; use linearize/solve_update with an all-reduce");
    if (check_tensors(p, t, motion_only)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    p->prof_iters = 0;

    // The second call with identical arguments captures the run into a CUDA graph; later calls replay it.
    // (The first call is never captured: one-off argument sets would only pay the instantiation.)
    vipe_ba_plan::GraphEntry *entry = nullptr;
    if (p->use_graphs && !p->profile && iterations > 0) {
        for (auto &g : p->graphs)
            if (same_args(g, t, ws, iterations, lm, ep, motion_only)) entry = &g;
        if (entry && entry->exec) {
            VBA_CUDA(cudaGraphLaunch(entry->exec, st));
            p->launches = entry->launches;
            return copy_graph_results(p, t, ws, motion_only, st);
        }
        if (!entry) {
            if (p->graphs.size() >= 8) {
                if (p->graphs.front().exec) cudaGraphExecDestroy(p->graphs.front().exec);
                p->graphs.erase(p->graphs.begin());
            }
            p->graphs.push_back({*t, ws, iterations, motion_only, lm, ep, 1, 0, nullptr});
            entry = nullptr;  // run directly this time
        } else {
            // capture on a private stream (the caller's may be the legacy default stream, which cannot be captured)
            if (!p->capture_stream) VBA_CUDA(cudaStreamCreateWithFlags(&p->capture_stream, cudaStreamNonBlocking));
            cudaGraph_t graph = nullptr;
            VBA_CUDA(cudaStreamBeginCapture(p->capture_stream, cudaStreamCaptureModeThreadLocal));
            p->launches = 0;
            int rc = 0;
            vipe_ba_tensors tg = *t;  // the graph's own result buffers
            tg.dx_out = (float *)((unsigned char *)ws + p->off_gdx);
            tg.dz_out = motion_only ? nullptr : (float *)((unsigned char *)ws + p->off_gdz);
            for (int it = 0; it < iterations && !rc; it++) rc = run_iteration(p, &tg, ws, lm, ep, motion_only, p->capture_stream, nullptr);
            cudaError_t ce = cudaStreamEndCapture(p->capture_stream, &graph);
            if (rc || ce != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                p->use_graphs = false;  // fall back to direct launches for good
                if (rc) return 1;
            } else {
                cudaGraphExec_t exec = nullptr;
                ce = cudaGraphInstantiate(&exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ce == cudaSuccess && exec) {
                    entry->exec = exec;
                    entry->launches = p->launches;
                    VBA_CUDA(cudaGraphLaunch(exec, st));
                    return copy_graph_results(p, t, ws, motion_only, st);
                }
                cudaGetLastError();
                p->use_graphs = false;
            }
        }
    }
    p->launches = 0;
    for (int it = 0; it < iterations; it++) {
        cudaEvent_t *ev = (p->profile && it < kMaxProfIters) ? &p->events[5 * it] : nullptr;
        if (run_iteration(p, t, ws, lm, ep, motion_only, st, ev)) return 1;
        if (ev) p->prof_iters = it + 1;
    }
    return 0;
}
