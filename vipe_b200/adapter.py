"""Adapter from the conventions of the BA that vipe/slam actually calls to this repo's operator.

`bundle_adjustment` mirrors `GraphBuffer.bundle_adjustment` (vipe/slam/components/buffer.py:373-525), which today runs
the pure-Python sparse solver (`Solver.run_inplace`, vipe/slam/ba/solver.py:117-197).  The reference itself names the
CUDA kernels as the intended accelerator for that term (vipe/slam/ba/terms.py:160).  Supported: one view, pinhole
camera, fixed rig, with or without the sparse-track flow term -- i.e. the frontend (frontend.py:100-115), the inner filler
(inner_filler.py:110-116) and the backend (backend.py:48-72) with or without `optimize_intrinsics` (for a pinhole
camera that is one extra variable, the focal length, retractor.py:51-62).  Everything else raises.

Differences between the two reference BAs that the adapter maps onto `vipe_ba_options` (SURVEY.md section 8(a')):
  * target / weight arrive channel-last `[E, ht*wd, 2]` (factor_graph.py:292-293);
  * a point is valid iff its depth is > 0.1 (cameras.py:48, geom.py:263) instead of >= 0.25;
  * disparity damping is `0.2*eta + 1e-7` (+ ep 1e-7) on every pixel (buffer.py:482-489);
  * the sensor prior (`alpha` from the config) is gated per frame and then applies to every pixel
    (buffer.py:470-479, terms.py:244-300);
  * only frames that are the source of an edge own a disparity variable; `limited_disp` additionally fixes the
    disparities of frames outside `[t0, t1)` (buffer.py:490-491);
  * LM damping scales the pose Hessian's own diagonal, before the Schur complement (solver.py:161-164), and every
    free pose takes part in the back-substitution (the CUDA BA drops pose index 0, geom_kernels.cu:1089);
  * `dz > 10 -> 0` (retractor.py:41), final `disps.clamp_(min=0.001)` (buffer.py:525), quaternion renormalised;
  * `optimize_intrinsics`: the focal length (fx and fy move together) joins the reduced system with damping
    1e-6 / 1e-6 (buffer.py:496-498); its Jacobian is scaled by 1/8 because the BA runs at 1/8 resolution
    (terms.py:186,224); `intrinsics` is updated in place.
The reduced system is solved in fp64 Cholesky here; the Python path uses fp32 SuperLU (solver.py:33-44).
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .ext import slam_ext
from .plan import cached_plan

INTRINSICS_FACTOR = 8.0  # buffer.py:413


def bundle_adjustment(poses, disps, disps_sens, intrinsics, target, weight, disp_damping, ii, jj, t0, t1, n_iters,
                      pose_damping, pose_ep, motion_only, limited_disp, optimize_intrinsics=False,
                      optimize_rig_rotation=False, dense_disp_alpha=0.001, sparse_target=None, sparse_weight=None, n_views=1):
    """poses[N,7], disps[N,ht,wd], disps_sens[N,ht,wd] (all updated/read in place, CUDA fp32); intrinsics[4] at FULL
    resolution (the Python path scales by 1/8, terms.py:186), updated in place when `optimize_intrinsics`;
    target/weight[E, ht*wd, 2]; disp_damping[N,ht,wd].  `sparse_target/sparse_weight[E, ht*wd, 2]`: the optional second
    flow term from sparse tracks (what `sparse_tracks.compute_dense_disp_target_weight` returns, buffer.py:422-449): the
    same edges with their own targets and weights, i.e. E more edges for the kernels."""
    if optimize_rig_rotation or n_views != 1:
        raise NotImplementedError("vipe_b200.adapter covers the single-view pinhole BA with a fixed rig")
    assert t0 <= t1
    dev = poses.device
    N, ht, wd = disps.shape
    HW = ht * wd
    E = ii.numel()
    if tuple(target.shape) != (E, HW, 2) or tuple(weight.shape) != (E, HW, 2):
        raise RuntimeError("target/weight must be channel-last [E, ht*wd, 2]")
    # channel-first copies in the operator's layout
    tgt = target.reshape(E, ht, wd, 2).permute(0, 3, 1, 2).contiguous()
    wgt = weight.reshape(E, ht, wd, 2).permute(0, 3, 1, 2).contiguous()
    intr = (intrinsics.reshape(-1)[:4] / INTRINSICS_FACTOR).contiguous()
    if sparse_target is not None:
        if tuple(sparse_target.shape) != (E, HW, 2) or tuple(sparse_weight.shape) != (E, HW, 2):
            raise RuntimeError("sparse_target/sparse_weight must be channel-last [E, ht*wd, 2]")
        # both terms carry the same 0.001 weight factor (buffer.py:396), so the track term is just a second set of edges
        tgt = torch.cat([tgt, sparse_target.reshape(E, ht, wd, 2).permute(0, 3, 1, 2)]).contiguous()
        wgt = torch.cat([wgt, sparse_weight.reshape(E, ht, wd, 2).permute(0, 3, 1, 2)]).contiguous()
        ii, jj = torch.cat([ii, ii]), torch.cat([jj, jj])

    ii_h = ii.detach().to("cpu", torch.int64).contiguous()
    jj_h = jj.detach().to("cpu", torch.int64).contiguous()
    plan = cached_plan(ii_h, jj_h, N, ht, wd, int(t0), int(t1), tag="python-ba-focal" if optimize_intrinsics else "python-ba")
    kx = plan.kx
    K = plan.K
    # disparity variables exist only for edge sources (di_unique, buffer.py:402); limited_disp fixes those outside the window
    is_source = torch.zeros(N, dtype=torch.bool)
    is_source[ii_h] = True
    fixed = ~is_source[kx]
    if limited_disp:
        fixed |= (kx < t0) | (kx >= t1)
    kx_d = kx.to(dev)
    gate = disps_sens.reshape(N, HW)[kx_d].sum(1) > 0.0  # per-frame gate, buffer.py:472 (no host sync)
    flags = gate.to(torch.uint8) | (fixed.to(dev).to(torch.uint8) << 1)
    plan.set_options(min_depth=0.1, depth_strict=1, alpha=float(dense_disp_alpha), sensor_mode=1, eta_scale=0.2,
                     eta_bias=2e-7, dz_max=10.0, renorm_quat=1, damp_on_pose_hessian=1, backsub_all_poses=1, frame_flags=flags,
                     optimize_focal=int(bool(optimize_intrinsics)), focal_jscale=1.0 / INTRINSICS_FACTOR, focal_lm=1e-6,
                     focal_ep=1e-6)
    eta = disp_damping.reshape(N, HW)[kx_d].contiguous()
    P = int(t1) - int(t0)
    dx_all = torch.zeros(6 * P + 1, dtype=torch.float32, device=dev)  # pose steps, then the focal step
    dx = dx_all[: 6 * P].view(P, 6)
    dz = torch.zeros(K, HW, dtype=torch.float32, device=dev)
    if n_iters > 0 and P > 0:
        with torch.cuda.device(dev):
            ws = plan.workspace(dev)
            tens = slam_ext._tensors(poses, disps, intr, disps_sens, tgt, wgt, eta, dx, dz, bool(motion_only))
            _lib.check(_lib.lib().vipe_ba_run(plan.handle, C.byref(tens), ws.data_ptr(), int(n_iters), float(pose_damping),
                                              float(pose_ep), int(bool(motion_only)), torch.cuda.current_stream(dev).cuda_stream),
                       "vipe_ba_run")
    if optimize_intrinsics and n_iters > 0 and P > 0:
        intrinsics.reshape(-1)[:2] = intr[:2] * INTRINSICS_FACTOR  # exact: the factor is a power of two
    disps.clamp_(min=0.001)  # buffer.py:525
    return dx, dz
