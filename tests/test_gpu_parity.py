"""GPU parity tests: the CUDA path (through the C ABI, via vipe_b200.ext.slam_ext) against the fp64 CPU oracle on
identical seeded inputs.  Tolerances are BASELINE.json's north-star figures: pose translation relative error and
rotation angle <= 1e-4, disparity relative error <= 1e-3, after the same number of Gauss-Newton iterations."""

import ctypes as C

import pytest
import torch

from oracle import ba_oracle as O
from vipe_b200.synthetic import BAConfig, disp_error, disp_error_p999, make_problem, pose_errors

pytestmark = pytest.mark.gpu

TOL_T, TOL_R, TOL_D = 1e-4, 1e-4, 1e-3


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def slam_ext(lib_built):
    from vipe_b200.ext import slam_ext as m

    return m


def _compare(slam_ext, dev, pr, args_cpu=None, overrides=None):
    overrides = overrides or {}
    ref = args_cpu() if args_cpu else pr.args()
    for k, v in overrides.items():
        ref[k] = v
    tr = O.Trace()
    dxr, dzr = O.ba(*ref, dtype=torch.float64, trace=tr)
    a = [x.clone().to(dev) if torch.is_tensor(x) else x for x in (args_cpu() if args_cpu else pr.args())]
    for k, v in overrides.items():
        a[k] = v
    dx, dz = slam_ext.ba(*a)
    torch.cuda.synchronize()
    t0, t1 = a[9], a[10]
    te, re_ = pose_errors(a[0], ref[0], t0, t1)
    de = disp_error(a[1], ref[1], tr.bk.kx)
    return dict(te=te, re=re_, de=de, a=a, ref=ref, tr=tr, dx=dx, dz=dz, dxr=dxr, dzr=dzr)


@pytest.mark.parametrize("name", ["c1", "c2", "c3"])
def test_full_ba_parity(slam_ext, dev, name):
    pr = make_problem(name)
    r = _compare(slam_ext, dev, pr)
    assert all(r["tr"].chol_ok)
    assert r["te"] <= TOL_T and r["re"] <= TOL_R and r["de"] <= TOL_D, (r["te"], r["re"], r["de"])
    # per-pixel view of the disparity error: its 99.9th percentile stays inside the same bound as the Frobenius figure
    p999 = disp_error_p999(r["a"][1], r["ref"][1], r["tr"].bk.kx)
    print(f"[parity {name}] translation rel {r['te']:.2e}  rotation max {r['re']:.2e} rad  disparity rel {r['de']:.2e}  "
          f"disparity |rel| p99.9 {p999:.2e}")
    assert p999 <= TOL_D, p999
    # gauge: poses outside [t0,t1) bit-identical to the input; frames outside kx untouched
    assert torch.equal(r["a"][0][: pr.t0].cpu(), pr.poses[: pr.t0])
    kx = set(r["tr"].bk.kx.tolist())
    for f in range(pr.cfg.n_frames):
        if f not in kx:
            assert torch.equal(r["a"][1][f].cpu(), pr.disps[f])
    # last-iteration updates
    # last-iteration update.  At C3 the 8th update has shrunk to the fp32 noise floor of the weakly constrained (gauge)
    # directions, so it is bounded in absolute terms there: the difference is below the size of a 1e-4 relative pose change
    if name != "c3":
        assert (r["dx"].cpu().double() - r["dxr"]).norm() <= 2e-2 * r["dxr"].norm() + 1e-7
    else:
        scale = r["ref"][0][pr.t0: pr.t1, :3].double().norm()
        assert (r["dx"].cpu().double() - r["dxr"]).norm() <= 1e-4 * scale, ((r["dx"].cpu().double() - r["dxr"]).norm(), scale)
    assert r["dz"].shape == (len(kx), pr.cfg.ht * pr.cfg.wd)


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_motion_only_parity(slam_ext, dev, name):
    pr = make_problem(name)
    r = _compare(slam_ext, dev, pr, overrides={14: True})
    assert r["te"] <= TOL_T and r["re"] <= TOL_R
    assert torch.equal(r["a"][1].cpu(), pr.disps), "motion-only must not touch disparities"


def test_sensor_prior_parity(slam_ext, dev):
    """disps_sens > 0 on even frames switches the per-pixel damping to alpha and adds the prior (:1359-1369)."""
    pr = make_problem("c2", sensor_on_even_frames=True)
    r = _compare(slam_ext, dev, pr)
    assert r["te"] <= TOL_T and r["re"] <= TOL_R and r["de"] <= TOL_D


def test_stage_outputs_match_oracle(slam_ext, dev):
    """After one linearisation: reduced camera system (A - S, b - v) and the disparity blocks Q, Q*w."""
    from vipe_b200 import _lib

    pr = make_problem("c2")
    cfg = pr.cfg
    a = pr.args(dev)
    plan = slam_ext.ba_plan(pr.ii, pr.jj, cfg.n_frames, cfg.ht, cfg.wd, pr.t0, pr.t1)
    ws = plan.workspace(dev)
    dx = torch.zeros(plan.P, 6, device=dev)
    dz = torch.zeros(plan.K, cfg.ht * cfg.wd, device=dev)
    tens = slam_ext._tensors(a[0], a[1], a[2], a[3], a[4], a[5], a[6], dx, dz, False)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().vipe_ba_linearize(plan.handle, C.byref(tens), ws.data_ptr(), 0, st), "linearize")
    torch.cuda.synchronize()
    sysv, npad = plan.system_view(ws)
    n = 6 * plan.P
    Hs = sysv[: npad * npad].view(npad, npad).cpu()
    Hs = torch.tril(Hs) + torch.tril(Hs, -1).T
    idx = plan.system_index()  # the system is stored in the plan's elimination order
    H = Hs[idx][:, idx]
    b = sysv[npad * npad:].cpu()[idx]
    o = pr.args()
    o[11] = 1
    tr = O.Trace()
    O.ba(*o, dtype=torch.float64, trace=tr)
    Aref, bref = tr.A - tr.S, tr.b - tr.sv.reshape(-1)
    assert (H - Aref).norm() <= 1e-5 * Aref.norm()
    assert (b - bref).norm() <= 1e-5 * bref.norm()
    q, qw = plan.debug_q(ws)
    assert (q.cpu().double() - tr.Q).norm() <= 1e-5 * tr.Q.norm()
    assert (qw.cpu().double() - tr.Q * tr.w).norm() <= 1e-4 * (tr.Q * tr.w).norm()
    # padded part of the system is the identity
    if npad > n:
        pad = sysv[: npad * npad].view(npad, npad)[n:, n:].cpu()
        assert torch.equal(pad, torch.eye(npad - n, dtype=torch.float64))


def _small_problem(seed, n_frames=7, ht=30, wd=41, t0=2, t1=None, stereo=True, dup=True, n_rand=14):
    """Irregular graph: odd image size (tail tiles, scalar loads), partial window, edges leaving fixed poses,
    a frame without outgoing edges, stereo and duplicate edges."""
    cfg = BAConfig(f"s{seed}", 100 + seed, n_frames, n_rand, ht, wd, 3, 1e-4, 0.1)
    pr = make_problem(cfg)
    gen = torch.Generator().manual_seed(seed)
    t1 = n_frames if t1 is None else t1
    ii, jj = pr.ii.clone(), pr.jj.clone()
    if stereo:
        ii[3], jj[3] = 4, 4
    if dup:
        ii[5], jj[5] = ii[4], jj[4]
    lonely = n_frames - 1  # no outgoing edges from the last frame
    jj = torch.where(ii == lonely, jj, jj)
    src_ok = ii != lonely
    ii, jj = ii[src_ok], jj[src_ok]
    keep = (ii < t1) & (jj < t1)
    ii, jj = ii[keep], jj[keep]
    sel = torch.nonzero(src_ok)[:, 0][keep]
    bk = O.bookkeeping(ii, jj, t0, t1)
    eta = 0.01 * torch.rand(bk.kx.numel(), ht, wd, generator=gen) + 1e-6

    def args():
        return [pr.poses.clone(), pr.disps.clone(), pr.intrinsics.clone(), pr.disps_sens.clone(),
                pr.targets[sel].clone().contiguous(), pr.weights[sel].clone().contiguous(), eta.clone(), ii.clone(), jj.clone(),
                t0, t1, 3, 1e-4, 0.1, False]

    return pr, args


def test_window_from_frame_zero(slam_ext, dev):
    """t0 = 0: no pose is held fixed (the damping alone removes the gauge freedom) and the quirk Q4 of the reference --
    the first pose of the window is skipped in the back-substitution -- hits frame 0 itself."""
    pr, args = _small_problem(5, t0=0, stereo=False)
    r = _compare(slam_ext, dev, pr, args_cpu=args)
    assert r["te"] <= TOL_T and r["re"] <= TOL_R and r["de"] <= TOL_D, (r["te"], r["re"], r["de"])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_irregular_graphs(slam_ext, dev, seed):
    pr, args = _small_problem(seed, t1=None if seed != 2 else 6)
    r = _compare(slam_ext, dev, pr, args_cpu=args)
    assert r["te"] <= TOL_T and r["re"] <= TOL_R and r["de"] <= TOL_D, (r["te"], r["re"], r["de"])
    a, ref = r["a"], r["ref"]
    assert torch.equal(a[0][: a[9]].cpu(), pr.poses[: a[9]]) and torch.equal(a[0][a[10]:].cpu(), pr.poses[a[10]:])
    kx = set(r["tr"].bk.kx.tolist())
    for f in range(pr.cfg.n_frames):
        if f not in kx:
            assert torch.equal(a[1][f].cpu(), pr.disps[f])
    # a frame inside the window without outgoing edges only sees its damping prior: dz = Q * w = 0 with no sensor
    assert (r["dz"].cpu().double() - r["dzr"]).norm() <= 1e-3 * r["dzr"].norm() + 1e-9


def test_invalid_depth_pixels(slam_ext, dev):
    """Points that land behind MIN_DEPTH get zero weight and zero Jacobians (:301-305)."""
    pr = make_problem("c1")

    def args():
        a = pr.args()
        a[1][2, 10:20, :] = 6.0  # huge disparity => z = 1 + h*t_z can fall below 0.25 for edges with t_z < 0
        a[0][3, 2] -= 0.2
        return a

    ref = args()
    h = ref[1].double().view(8, -1)[pr.ii]
    tij, qij, _ = O.relative_poses(ref[0].double(), pr.ii, pr.jj)
    assert bool(((1 + h * tij[:, 2:3]) < 0.3).any()), "test must actually exercise invalid pixels"
    r = _compare(slam_ext, dev, pr, args_cpu=args, overrides={11: 1})
    assert r["te"] <= TOL_T and r["re"] <= TOL_R and r["de"] <= TOL_D, (r["te"], r["re"], r["de"])


def test_failed_factorisation_is_silent_zero_update(slam_ext, dev):
    """NaN in the system => LLT fails => dx = 0, poses unchanged (:1181-1188)."""
    pr = make_problem("c1")
    a = pr.args(dev)
    a[5][0, 0, 0, 0] = float("nan")
    a[14] = True
    dx, _ = slam_ext.ba(*a)
    torch.cuda.synchronize()
    assert torch.equal(dx.cpu(), torch.zeros(7, 6))
    assert torch.equal(a[0].cpu(), pr.poses)


@pytest.mark.timeout(120)
def test_failed_factorisation_tiled_solver(slam_ext, dev):
    """Same on the tile-dataflow solver (6P > 128): a NaN weight poisons the system; the factorisation must finish (its
    tiles and the backward substitution's x travel as self-validating words -- NaN is a valid word, nothing may spin on
    it), report failure and leave the poses alone; the next call on clean inputs works."""
    cfg = BAConfig("tiled_fail", 41, 48, 300, 24, 32, 2, 1e-4, 0.1)
    pr = make_problem(cfg)
    a = pr.args(dev)
    a[5][3, 0, 0, 0] = float("nan")
    a[14] = True
    dx, _ = slam_ext.ba(*a)
    torch.cuda.synchronize()
    assert torch.equal(dx.cpu(), torch.zeros(pr.t1 - pr.t0, 6))
    assert torch.equal(a[0].cpu(), pr.poses)
    r = _compare(slam_ext, dev, pr)
    assert r["te"] <= TOL_T and r["re"] <= TOL_R and r["de"] <= TOL_D, (r["te"], r["re"], r["de"])


@pytest.mark.parametrize("name", ["c3", "c4"])
def test_full_size_properties(slam_ext, dev, name):
    """Size-independent properties at BASELINE.json's full sizes (the oracle is too slow at C4):
    (1) noise-free targets + ground-truth start is a fixed point, (2) from a perturbed start BA converges
    to the ground truth, (3) a second call from the converged state barely moves."""
    pr = make_problem(name, noise_px=0.0, perturb=0.0)
    a = pr.args(dev)
    dx, dz = slam_ext.ba(*a)
    torch.cuda.synchronize()
    assert dx.abs().max() < 2e-5 and dz.abs().median() < 1e-5
    te, re_ = pose_errors(a[0], pr.poses_gt, pr.t0, pr.t1)
    assert te < 1e-4 and re_ < 1e-4

    pr2 = make_problem(name, noise_px=0.0, perturb=1.0)
    a = pr2.args(dev)
    te0, re0 = pose_errors(a[0], pr2.poses_gt, pr2.t0, pr2.t1)
    slam_ext.ba(*a)
    te1, re1 = pose_errors(a[0], pr2.poses_gt, pr2.t0, pr2.t1)
    assert re1 < 0.05 * re0, (re0, re1)
    # translation/disparity share the monocular scale gauge: compare after removing the best common scale
    t_est, t_gt = a[0][1:, :3].double().cpu(), pr2.poses_gt[1:, :3].double()
    s = (t_est * t_gt).sum() / (t_gt * t_gt).sum()
    assert ((t_est - s * t_gt).norm() / t_gt.norm()) < 0.05 * te0, (te0, te1)
    before = a[0].clone()
    dx, dz = slam_ext.ba(*a)
    torch.cuda.synchronize()
    assert dx.abs().max() < 1e-3
    assert pose_errors(a[0], before, pr2.t0, pr2.t1)[0] < 1e-3


@pytest.mark.parametrize("name", ["c2", "c3"])
def test_plan_cache_and_repeat_calls_are_deterministic(slam_ext, dev, name):
    pr = make_problem(name)
    outs = []
    for _ in range(2):
        a = pr.args(dev)
        slam_ext.ba(*a)
        outs.append((a[0].cpu(), a[1].cpu()))
    # the reduced system is assembled in a fixed order (assemble_kernel) and every other sum is a fixed-shape tree:
    # repeated calls are bit-identical, like the reference's Eigen assembly
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("iters", [2, 8])
def test_c4_vs_reference_run(slam_ext, dev, iters):
    """Full-size C4 (1000 keyframes, 12000 edges, 64x112): two and all eight Gauss-Newton iterations against the REFERENCE's
    own CUDA slam_ext.ba (oracle/_ref) on the same inputs -- the fp64 oracle needs >10 GB and minutes at this size."""
    from oracle import build_ref

    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref/vipe_ref_ext.so not built")
    pr = make_problem("c4")
    a, b = pr.args(dev), pr.args(dev)
    a[11] = b[11] = iters
    dxr, dzr = mod.slam_ext.ba(*a)
    dx, dz = slam_ext.ba(*b)
    torch.cuda.synchronize()
    te, re_ = pose_errors(b[0], a[0], pr.t0, pr.t1)
    kx = torch.unique(torch.cat([torch.arange(pr.t0, pr.t1), pr.ii]))
    de = disp_error(b[1], a[1], kx)
    print("C4 vs reference run:", te, re_, de, float((dx - dxr).norm() / dxr.norm()), float((dz - dzr).norm() / dzr.norm()))
    assert te <= TOL_T and re_ <= TOL_R and de <= TOL_D, (te, re_, de)
    if iters == 2:  # (after 8 iterations the update itself is at the noise floor)
        assert (dx - dxr).norm() <= 5e-2 * dxr.norm()


@pytest.mark.parametrize("motion_only", [True, False])
def test_batched_clips_match_per_clip_oracle(slam_ext, dev, motion_only):
    """ba_batch: independent clips in one set of launches; every clip must equal its own oracle run (BASELINE config 5
    is the motion-only case)."""
    clips = [make_problem("c1", clip=c) for c in range(5)]
    N, E = 8, 24
    cat = lambda xs: torch.cat(xs, dim=0)
    poses, disps = cat([p.poses for p in clips]), cat([p.disps for p in clips])
    dsens = cat([p.disps_sens for p in clips])
    targets, weights = cat([p.targets for p in clips]), cat([p.weights for p in clips])
    eta = cat([p.eta for p in clips])  # every clip's kx is all 8 frames
    ii = cat([p.ii + c * N for c, p in enumerate(clips)])
    jj = cat([p.jj + c * N for c, p in enumerate(clips)])
    frame_ptr = [c * N for c in range(6)]
    t0s = [c * N + 1 for c in range(5)]
    t1s = [(c + 1) * N for c in range(5)]
    a = [x.to(dev) for x in (poses, disps, clips[0].intrinsics, dsens, targets, weights, eta, ii, jj)]
    dx, dz = slam_ext.ba_batch(*a, frame_ptr, t0s, t1s, 2, 1e-4, 0.1, motion_only)
    torch.cuda.synchronize()
    assert dx.shape == (35, 6)
    for c, p in enumerate(clips):
        ref = p.args()
        ref[14] = motion_only
        tr = O.Trace()
        dxr, dzr = O.ba(*ref, dtype=torch.float64, trace=tr)
        pc = a[0][c * N:(c + 1) * N]
        te, re_ = pose_errors(pc, ref[0], 1, N)
        assert te <= TOL_T and re_ <= TOL_R, (c, te, re_)
        assert torch.equal(pc[:1].cpu(), p.poses[:1])
        if motion_only:
            assert torch.equal(a[1][c * N:(c + 1) * N].cpu(), p.disps)
        else:
            assert disp_error(a[1][c * N:(c + 1) * N], ref[1], tr.bk.kx) <= TOL_D
        assert (dx[c * 7:(c + 1) * 7].cpu().double() - dxr).norm() <= 2e-2 * dxr.norm() + 1e-7


@pytest.mark.parametrize("name", ["c2", "c3"])
def test_tensor_core_pipeline_parity(slam_ext, dev, name, monkeypatch):
    """VIPE_BA_LIN4=1 routes the full linearisation through the two-kernel Blackwell pipeline of ba_lin4.cu (disparity blocks
    first; then asynchronously fed J warps, TF32 hi/lo Schur Gram by tcgen05.mma with the accumulators in tensor memory); same
    bounds as the default path.  C2 has frames with fewer edges than J warps (the case that needs the completion counters), C3
    has frames on both sides of the 10-edge boundary (one MMA per K step / four)."""
    from vipe_b200 import plan as plan_mod

    monkeypatch.setenv("VIPE_BA_LIN4", "1")
    plan_mod._CACHE.clear()
    slam_ext._LAST_PLANS.clear()
    try:
        pr = make_problem(name)
        r = _compare(slam_ext, dev, pr)
        print(f"[tc parity {name}] translation rel {r['te']:.2e}  rotation max {r['re']:.2e} rad  disparity rel {r['de']:.2e}")
        assert r["te"] <= TOL_T and r["re"] <= TOL_R and r["de"] <= TOL_D, (r["te"], r["re"], r["de"])
    finally:
        plan_mod._CACHE.clear()
        slam_ext._LAST_PLANS.clear()


def test_c5_real_configuration(slam_ext, dev):
    """BASELINE config 5 as specified: 64 independent clips of the C2 shape (16 keyframes, 120 edges, 48x64), motion-only,
    4 Gauss-Newton iterations, ONE batched call; a sample of the clips is checked against the fp64 oracle run per clip."""
    from vipe_b200.synthetic import CONFIGS

    cfg = CONFIGS["c5"]
    clips = [make_problem(cfg, clip=c) for c in range(cfg.clips)]
    N, nc = cfg.n_frames, cfg.clips
    cat = lambda xs: torch.cat(xs, dim=0)
    pr = clips[0]
    a = [cat([p.poses for p in clips]).to(dev), cat([p.disps for p in clips]).to(dev), pr.intrinsics.to(dev),
         cat([p.disps_sens for p in clips]).to(dev), cat([p.targets for p in clips]).to(dev), cat([p.weights for p in clips]).to(dev),
         cat([p.eta for p in clips]).to(dev), cat([p.ii + c * N for c, p in enumerate(clips)]).to(dev),
         cat([p.jj + c * N for c, p in enumerate(clips)]).to(dev)]
    d0 = a[1].clone()
    dx, dz = slam_ext.ba_batch(*a, [c * N for c in range(nc + 1)], [c * N + pr.t0 for c in range(nc)],
                               [c * N + pr.t1 for c in range(nc)], cfg.iters, cfg.lm, cfg.ep, True)
    torch.cuda.synchronize()
    assert torch.equal(a[1], d0), "motion-only must not touch disparities"
    P = pr.t1 - pr.t0
    for c in (0, 1, 17, 40, 63):
        ref = clips[c].args()
        dxr, _ = O.ba(*ref, dtype=torch.float64)
        pc = a[0][c * N:(c + 1) * N]
        te, re_ = pose_errors(pc, ref[0], pr.t0, pr.t1)
        assert te <= TOL_T and re_ <= TOL_R, (c, te, re_)
        assert torch.equal(pc[: pr.t0].cpu(), clips[c].poses[: pr.t0])
        assert (dx[c * P:(c + 1) * P].cpu().double() - dxr).norm() <= 2e-2 * dxr.norm() + 1e-7


def _hub_problem(n_frames, ht=24, wd=32):
    """Star graph: frame 0 is the source of an edge to every other frame (degree n_frames - 1) and their target."""
    cfg = BAConfig(f"hub{n_frames}", 300 + n_frames, n_frames, 2 * (n_frames - 1), ht, wd, 2, 1e-4, 0.1, trajectory="orbit")
    pr = make_problem(cfg)
    others = torch.arange(1, n_frames)
    ii = torch.cat([torch.zeros(n_frames - 1, dtype=torch.int64), others])
    jj = torch.cat([others, torch.zeros(n_frames - 1, dtype=torch.int64)])
    # targets of the hub edges: the true reprojection of the (perturbed) state is fine for a parity test
    gen = torch.Generator().manual_seed(n_frames)
    E = ii.numel()
    targets = pr.targets[:1].expand(E, -1, -1, -1).clone() + torch.randn(E, 2, ht, wd, generator=gen)
    weights = torch.rand(E, 2, ht, wd, generator=gen)
    eta = 0.01 * torch.rand(n_frames, ht, wd, generator=gen) + 1e-6

    def args():
        return [pr.poses.clone(), pr.disps.clone(), pr.intrinsics.clone(), pr.disps_sens.clone(), targets.clone(), weights.clone(),
                eta.clone(), ii.clone(), jj.clone(), 1, n_frames, 1, 1e-4, 0.1, False]

    return pr, args


@pytest.mark.parametrize("n_frames,ht,wd", [(72, 24, 32), (151, 24, 32), (301, 16, 16)])
def test_high_degree_source_frame(slam_ext, dev, n_frames, ht, wd):
    """A source frame with 71 / 150 / 300 outgoing edges (the reference has no out-degree limit): its staging buffer does not
    fit shared memory, so the plan sends it through the global-memory staging launch (and, at 300 edges, frame_reduce through
    its global working set) while the other frames keep their tile; the result must still match the oracle."""
    pr, args = _hub_problem(n_frames, ht, wd)
    r = _compare(slam_ext, dev, pr, args_cpu=args)
    assert (r["dx"].cpu().double() - r["dxr"]).norm() <= 1e-4 * r["dxr"].norm(), float((r["dx"].cpu().double() - r["dxr"]).norm() / r["dxr"].norm())
    assert (r["dz"].cpu().double() - r["dzr"]).norm() <= 1e-3 * r["dzr"].norm()


def test_degree_beyond_every_buffer_fails_loudly(slam_ext, dev):
    """899 outgoing edges: even the per-edge constants of the hub launch no longer fit shared memory (limit ~690); the call must
    raise, not corrupt anything, and the process must stay usable."""
    pr, args = _hub_problem(900, ht=8, wd=8)
    a = [x.to(dev) if torch.is_tensor(x) else x for x in args()]
    with pytest.raises(RuntimeError):
        slam_ext.ba(*a)
    pr2 = make_problem("c1")
    r = _compare(slam_ext, dev, pr2)
    assert r["te"] <= TOL_T


def test_identity_plan_cache_sees_in_place_edits(slam_ext, dev):
    """The identity shortcut in front of the plan cache must notice an edge list edited in place."""
    pr = make_problem("c1")
    a = pr.args(dev)
    p1 = slam_ext.ba_plan(a[7], a[8], 8, 48, 64, 1, 8)
    assert slam_ext.ba_plan(a[7], a[8], 8, 48, 64, 1, 8) is p1
    a[8][0] = (int(a[8][0]) + 3) % 8  # in-place edit: version counter moves
    p2 = slam_ext.ba_plan(a[7], a[8], 8, 48, 64, 1, 8)
    assert p2 is not p1
    assert slam_ext.ba_plan(a[7].clone(), a[8].clone(), 8, 48, 64, 1, 8) is p2  # same content, other objects: hash cache


def test_host_feed_matches_direct_calls(slam_ext, dev):
    """HostFeed (uploads on a copy stream, two alternating device argument sets) gives the results of calling
    slam_ext.ba on device copies, call after call, including slot reuse (not bit-for-bit: the assembly of the reduced
    system uses fp64 atomics, whose order varies from run to run)."""
    from vipe_b200.host_feed import HostFeed

    problems_ = [make_problem("c1"), make_problem("c2"), make_problem("c1"), make_problem("c2"), make_problem("c1")]
    want = []
    for pr in problems_:
        a = pr.args(dev)
        slam_ext.ba(*a)
        want.append((a[0].cpu(), a[1].cpu()))
    feed = HostFeed(dev)
    host = [[x.pin_memory() if torch.is_tensor(x) else x for x in pr.args()] for pr in problems_]
    outs = [(torch.empty_like(h[0]).pin_memory(), torch.empty_like(h[1]).pin_memory()) for h in host]
    feed.prefetch(host[0])
    for q in range(len(host)):
        if q + 1 < len(host):
            feed.prefetch(host[q + 1])
        feed.run(out_poses=outs[q][0], out_disps=outs[q][1])
    feed.synchronize()
    for (p, d), (wp, wd) in zip(outs, want):
        assert torch.allclose(p, wp, rtol=1e-5, atol=1e-6) and torch.allclose(d, wd, rtol=1e-5, atol=1e-6)
    with pytest.raises(RuntimeError):
        feed.run()  # nothing prefetched


def test_zero_iterations_and_empty_graph(slam_ext, dev):
    """Empty inputs: `iterations = 0` touches nothing; a graph without edges has a zero right-hand side, so one
    iteration returns dx = 0, dz = 0 and leaves poses and disparities as they were."""
    pr = make_problem("c1")
    a = pr.args(dev)
    a[11] = 0
    dx, dz = slam_ext.ba(*a)
    torch.cuda.synchronize()
    assert torch.count_nonzero(dx) == 0 and torch.count_nonzero(dz) == 0
    assert torch.equal(a[0].cpu(), pr.poses) and torch.equal(a[1].cpu(), pr.disps)

    a = pr.args(dev)
    ht, wd = pr.cfg.ht, pr.cfg.wd
    a[4] = torch.zeros(0, 2, ht, wd, device=dev)
    a[5] = torch.zeros(0, 2, ht, wd, device=dev)
    a[7] = torch.zeros(0, dtype=torch.int64, device=dev)
    a[8] = torch.zeros(0, dtype=torch.int64, device=dev)
    K = pr.t1 - pr.t0  # kx = the window's frames when there are no edges
    a[6] = a[6][:K].contiguous()
    a[11] = 1
    dx, dz = slam_ext.ba(*a)
    torch.cuda.synchronize()
    assert dx.shape == (K, 6) and torch.count_nonzero(dx) == 0
    assert torch.count_nonzero(dz) == 0
    assert torch.equal(a[0].cpu(), pr.poses) and torch.equal(a[1].cpu(), pr.disps)
