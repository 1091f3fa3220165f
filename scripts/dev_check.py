"""Developer check on a GPU box: stage-level and end-to-end parity of slam_ext.ba vs the fp64 oracle,
plus rough timings.  Usage: python scripts/dev_check.py [c1 c2 c3 ...]"""

import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import ba_oracle as O  # noqa: E402
from vipe_b200 import _lib  # noqa: E402
from vipe_b200.ext import slam_ext  # noqa: E402
from vipe_b200.synthetic import CONFIGS, disp_error, make_problem, pose_errors  # noqa: E402
import ctypes as C  # noqa: E402


def stage_check(pr, motion_only=False):
    dev = torch.device("cuda:0")
    a = pr.args(dev)
    cfg = pr.cfg
    plan = slam_ext.ba_plan(pr.ii, pr.jj, cfg.n_frames, cfg.ht, cfg.wd, pr.t0, pr.t1)
    ws = plan.workspace(dev)
    P, K, HW = plan.P, plan.K, cfg.ht * cfg.wd
    dx = torch.zeros(P, 6, device=dev)
    dz = torch.zeros(K, HW, device=dev)
    tens = slam_ext._tensors(a[0], a[1], a[2], a[3], a[4], a[5], a[6], dx, dz, motion_only)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().vipe_ba_linearize(plan.handle, C.byref(tens), ws.data_ptr(), int(motion_only), st), "lin")
    torch.cuda.synchronize()
    sysv, npad = plan.system_view(ws)
    H = sysv[: npad * npad].view(npad, npad).cpu()
    n = 6 * P
    idx = plan.system_index()  # the system is stored in the plan's elimination order
    Hl = torch.tril(H)
    Hfull = (Hl + torch.tril(Hl, -1).T)[idx][:, idx]
    b = sysv[npad * npad:].cpu()[idx]

    # oracle, one iteration
    o = pr.args()
    o[11] = 1
    o[14] = motion_only
    tr = O.Trace()
    O.ba(*o, dtype=torch.float64, trace=tr)
    Aref = tr.A if motion_only else tr.A - tr.S
    bref = tr.b if motion_only else tr.b - tr.sv.reshape(-1)
    print(f"  [stage] H rel err {((Hfull - Aref).norm() / Aref.norm()).item():.3e}  max abs {((Hfull - Aref).abs().max()).item():.3e} (scale {Aref.abs().max().item():.3e})")
    print(f"  [stage] b rel err {((b[:n] - bref).norm() / bref.norm()).item():.3e}")
    if not motion_only:
        q, qw = plan.debug_q(ws)
        print(f"  [stage] Q rel err {((q.cpu().double() - tr.Q).norm() / tr.Q.norm()).item():.3e}  Qw rel err {((qw.cpu().double() - tr.Q * tr.w).norm() / (tr.Q * tr.w).norm()).item():.3e}")
    _lib.check(_lib.lib().vipe_ba_solve_update(plan.handle, C.byref(tens), ws.data_ptr(), cfg.lm, cfg.ep, int(motion_only), st), "solve")
    torch.cuda.synchronize()
    print(f"  [stage] dx rel err {((dx.cpu().double() - tr.dx).norm() / tr.dx.norm()).item():.3e}")
    if not motion_only:
        print(f"  [stage] dz rel err {((dz.cpu().double() - tr.dz).norm() / tr.dz.norm()).item():.3e}")
        print("  [stage] disp err", disp_error(a[1], o[1], tr.bk.kx))
    print("  [stage] pose err", pose_errors(a[0], o[0], pr.t0, pr.t1))


def full_check(pr):
    dev = torch.device("cuda:0")
    cfg = pr.cfg
    o = pr.args()
    t = time.time()
    tr = O.Trace()
    O.ba(*o, dtype=torch.float64, trace=tr)
    t_or = time.time() - t
    a = pr.args(dev)
    dx, dz = slam_ext.ba(*a)
    torch.cuda.synchronize()
    te, re_ = pose_errors(a[0], o[0], pr.t0, pr.t1)
    de = disp_error(a[1], o[1], tr.bk.kx) if not cfg.motion_only else 0.0
    print(f"  [full] iters={cfg.iters} trans_rel={te:.3e} rot_max={re_:.3e} disp_rel={de:.3e} chol_ok={tr.chol_ok} oracle_s={t_or:.2f}")
    fixed_same = torch.equal(a[0][: pr.t0].cpu(), pr.poses[: pr.t0])
    print(f"  [full] fixed poses bit-identical: {fixed_same}")
    # timing
    ts = []
    for _ in range(5):
        a = pr.args(dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        slam_ext.ba(*a)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"  [time] {ms:.3f} ms per call, {ms / cfg.iters * 1e3:.1f} us per iteration, {pr.edge_pixels * cfg.iters / ms / 1e6:.2f} G edge-px/s")


if __name__ == "__main__":
    names = sys.argv[1:] or ["c1", "c2"]
    for name in names:
        print("==", name)
        pr = make_problem(name)
        if name != "c5":
            stage_check(pr)
        stage_check(pr, motion_only=True)
        full_check(pr)
