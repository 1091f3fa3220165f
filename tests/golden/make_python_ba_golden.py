"""Golden vectors from the reference's PYTHON bundle adjustment (the solver vipe/slam really calls):
`GraphBuffer.bundle_adjustment` (vipe/slam/components/buffer.py:373-525) -> `Solver.run_inplace`
(vipe/slam/ba/solver.py:117-197) -> `DenseDepthFlowTerm` / `DispSensRegularizationTerm` (vipe/slam/ba/terms.py).

Runs IN THE BUILD CONTAINER ONLY (needs /root/reference); the outputs are committed under tests/golden/ and the
adapter tests (tests/test_adapter.py) compare the CUDA path against them.

The reference modules are imported UNMODIFIED from /root/reference.  What the image lacks is replaced by stand-ins
registered in sys.modules before the import (nothing of the reference is copied):
  * `omegaconf`, `vipe.pipeline`          -- imported by vipe/__init__.py, unused by the BA
  * `vipe_ext`                            -- the compiled extension module; its `lietorch_ext` backend is restated
                                             here in torch for SE3 (exp, inv, mul, adjT, act, act4) following
                                             csrc/lietorch_ext/se3.h / so3.h; the other submodules are empty
The body of `bundle_adjustment` (which lives in a class that needs rerun/GUI imports) is restated below line by
line for the single-view pinhole case without intrinsics / rig optimisation and without sparse tracks.

Re-running the script reproduces the committed vectors to ~2e-6 only (the Python solver's threaded fp32 sums and
SuperLU are not bit-reproducible run to run); the tolerances of the tests are 1e-4 / 1e-3.

Usage: python tests/golden/make_python_ba_golden.py
"""

from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")


# ----------------------------------------------------------------------------- lietorch backend stand-in (SE3)
def _quat_mul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                        aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], dim=-1)


def _rot(q, v):
    qv, qw = q[..., :3], q[..., 3:4]
    uv = 2.0 * torch.linalg.cross(qv, v, dim=-1)
    return v + qw * uv + torch.linalg.cross(qv, uv, dim=-1)


def _hat(v):
    o = torch.zeros_like(v[..., 0])
    return torch.stack([torch.stack([o, -v[..., 2], v[..., 1]], -1), torch.stack([v[..., 2], o, -v[..., 0]], -1),
                        torch.stack([-v[..., 1], v[..., 0], o], -1)], -2)


def _so3_exp(phi):  # csrc/lietorch_ext/so3.h Exp
    th2 = (phi * phi).sum(-1, keepdim=True)
    th = th2.sqrt()
    small = th2 < 1e-8  # EPS-style branch; both branches agree to fp32 there
    ths = torch.where(small, torch.ones_like(th), th)
    imag = torch.where(small, 0.5 - th2 / 48.0 + th2 * th2 / 3840.0, torch.sin(0.5 * ths) / ths)
    real = torch.where(small, 1.0 - th2 / 8.0 + th2 * th2 / 384.0, torch.cos(0.5 * ths))
    return torch.cat([imag * phi, real], -1)


def _left_jacobian(phi):  # so3.h left_jacobian
    th2 = (phi * phi).sum(-1)[..., None, None]
    th = th2.sqrt()
    Phi = _hat(phi)
    Phi2 = Phi @ Phi
    eye = torch.eye(3, dtype=phi.dtype).expand_as(Phi)
    small = th2 < 1e-8
    ths = torch.where(small, torch.ones_like(th), th)
    th2s = torch.where(small, torch.ones_like(th2), th2)
    a = torch.where(small, torch.full_like(th, 0.5), (1 - torch.cos(ths)) / th2s)
    b = torch.where(small, torch.full_like(th, 1.0 / 6.0), (ths - torch.sin(ths)) / (th2s * ths))
    return eye + a * Phi + b * Phi2


class _LieBackend(types.ModuleType):
    def _se3(self, gid):
        assert gid == 3, "only SE3 is needed by the BA path"

    def expm(self, gid, a):
        self._se3(gid)
        q = _so3_exp(a[:, 3:])
        t = (_left_jacobian(a[:, 3:]) @ a[:, :3, None])[..., 0]
        return torch.cat([t, q], -1)

    def inv(self, gid, X):
        self._se3(gid)
        qi = torch.cat([-X[:, 3:6], X[:, 6:7]], -1)
        return torch.cat([-_rot(qi, X[:, :3]), qi], -1)

    def mul(self, gid, X, Y):
        self._se3(gid)
        q = _quat_mul(X[:, 3:], Y[:, 3:])
        q = q / q.norm(dim=-1, keepdim=True)  # so3.h:36-38
        return torch.cat([X[:, :3] + _rot(X[:, 3:], Y[:, :3]), q], -1)

    def adjT(self, gid, X, a):  # Adj = [[R, tx R], [0, R]]; returns Adj^T a  (se3.h:60-83)
        self._se3(gid)
        qi = torch.cat([-X[:, 3:6], X[:, 6:7]], -1)
        top = _rot(qi, a[:, :3])  # R^T a1
        bot = _rot(qi, torch.linalg.cross(a[:, :3], X[:, :3], dim=-1)) + _rot(qi, a[:, 3:])  # (tx R)^T a1 + R^T a2
        return torch.cat([top, bot], -1)

    def act(self, gid, X, p):
        self._se3(gid)
        return _rot(X[:, 3:], p) + X[:, :3]

    def act4(self, gid, X, p):
        self._se3(gid)
        return torch.cat([_rot(X[:, 3:], p[:, :3]) + X[:, :3] * p[:, 3:4], p[:, 3:4]], -1)

    def __getattr__(self, name):  # every other backend entry (backward ops, log, ...) is unused here
        return None


def install_stubs():
    om = types.ModuleType("omegaconf")

    class OmegaConf:  # noqa: D401
        @staticmethod
        def has_resolver(name):
            return True

        @staticmethod
        def register_new_resolver(*a, **k):
            return None

    om.OmegaConf = OmegaConf
    sys.modules["omegaconf"] = om
    pipe = types.ModuleType("vipe.pipeline")
    pipe.make_pipeline = lambda *a, **k: None
    sys.modules["vipe.pipeline"] = pipe
    ext = types.ModuleType("vipe_ext")
    for name in ("droid_net_ext", "grounding_dino_ext", "utils_ext", "slam_ext", "scatter_ext", "corr_ext"):
        setattr(ext, name, types.ModuleType(name))
    ext.lietorch_ext = _LieBackend("lietorch_ext")
    sys.modules["vipe_ext"] = ext
    sys.path.insert(0, str(REF))


# ----------------------------------------------------------------------------- bundle_adjustment, restated call by call
def python_bundle_adjustment(poses, disps, disps_sens, intrinsics_full, target, weight, disp_damping, ii, jj, t0, t1,
                             n_iters, pose_damping, pose_ep, motion_only, limited_disp, alpha, ht, wd, optimize_intrinsics=False,
                             sparse_target=None, sparse_weight=None):
    """buffer.py:373-525 for n_views == 1, pinhole, no intrinsics / rig optimisation, no sparse tracks.
    poses[N,7], disps[N,ht,wd] (updated in place), intrinsics_full[1,4] at full resolution (factor 8),
    target/weight[E, ht*wd, 2] channel-last, disp_damping[N,ht,wd]."""
    from einops import rearrange
    from vipe.ext.lietorch import SE3
    from vipe.slam.ba.solver import Solver
    from vipe.slam.ba.terms import DenseDepthFlowTerm, DispSensRegularizationTerm
    from vipe.slam.maths.retractor import DenseDispRetractor, IntrinsicsRetractor, PoseRetractor, RigRotationOnlyRetractor
    from vipe.slam.maths.vector import SparseBlockVector
    from vipe.utils.cameras import CameraType

    weight_dense_disp = 0.001  # buffer.py:396
    pi, pj, di = ii, jj, ii    # expand_edge_multiview with one view
    qi = qj = torch.zeros_like(ii)
    di_unique = torch.unique(di)
    pi_unique = torch.unique(ii)
    solver = Solver(compute_energy=False)
    solver.add_term(DenseDepthFlowTerm(pose_i_inds=pi, pose_j_inds=pj, rig_i_inds=qi, rig_j_inds=qj, dense_disp_i_inds=di,
                                       target=target, weight=weight_dense_disp * weight, intrinsics=None,
                                       intrinsics_factor=8.0, rig=None, image_size=(ht, wd), camera_type=CameraType.PINHOLE))
    if sparse_target is not None:  # the sparse-track flow term: same edges, its own targets and weights (buffer.py:422-449)
        solver.add_term(DenseDepthFlowTerm(pose_i_inds=pi, pose_j_inds=pj, rig_i_inds=qi, rig_j_inds=qj, dense_disp_i_inds=di,
                                           target=sparse_target, weight=0.001 * sparse_weight, intrinsics=None,
                                           intrinsics_factor=8.0, rig=None, image_size=(ht, wd), camera_type=CameraType.PINHOLE))
    solver.set_fixed("pose", torch.cat([pi_unique[pi_unique < t0], pi_unique[pi_unique >= t1]]) if t0 < t1 else None)
    solver.set_retractor("pose", PoseRetractor())
    solver.set_damping("pose", damping=pose_damping, ep=pose_ep)
    if not motion_only:
        dsens = rearrange(disps_sens, "nv h w -> nv (h w)")
        sens_i_inds = di_unique[dsens[di_unique].sum(1) > 0.0]
        if len(sens_i_inds) > 0:
            solver.add_term(DispSensRegularizationTerm(i_inds=sens_i_inds, alpha=alpha, disps_sens=dsens))
        solver.set_retractor("dense_disp", DenseDispRetractor())
        dd = rearrange(disp_damping, "nv h w -> nv (h w)")
        solver.set_damping("dense_disp", damping=SparseBlockVector(inds=di_unique, data=0.2 * dd[di_unique] + 1e-7), ep=1e-7)
        if limited_disp:
            solver.set_fixed("dense_disp", torch.cat([di[pi < t0], di[pi >= t1]]))
    else:
        solver.set_fixed("dense_disp")
    solver.set_marginilized("dense_disp")
    solver.set_retractor("intrinsics", IntrinsicsRetractor(CameraType.PINHOLE))
    solver.set_damping("intrinsics", damping=1e-6, ep=1e-6)
    if not optimize_intrinsics:  # buffer.py:497-498
        solver.set_fixed("intrinsics")
    solver.set_retractor("rig", RigRotationOnlyRetractor())
    solver.set_damping("rig", damping=1e-4, ep=1e-4)
    solver.set_fixed("rig")
    rig = torch.tensor([[0.0, 0, 0, 0, 0, 0, 1]])
    disps_flat = rearrange(disps, "nv h w -> nv (h w)")
    for _ in range(n_iters):
        solver.run_inplace({"pose": SE3(poses), "dense_disp": disps_flat, "intrinsics": intrinsics_full, "rig": SE3(rig)})
    disps.clamp_(min=0.001)  # buffer.py:525


CASES = {
    # name: (config, kwargs for the problem, BA flags)
    "pyba_c1_full": ("c1", {}, dict(motion_only=False, limited_disp=False, t0=1)),
    "pyba_c2_full": ("c2", {}, dict(motion_only=False, limited_disp=False, t0=1)),
    "pyba_c1_motion": ("c1", {}, dict(motion_only=True, limited_disp=False, t0=1)),
    "pyba_c2_sensor_limited": ("c2", {"sensor_on_even_frames": True}, dict(motion_only=False, limited_disp=True, t0=4)),
    # backend default (configs/pipeline/default.yaml:26): the focal length is a variable too; it starts 3 % off
    "pyba_c2_focal": ("c2", {}, dict(motion_only=False, limited_disp=False, t0=1, optimize_intrinsics=True, focal_scale=1.03)),
    "pyba_c1_focal_motion": ("c1", {}, dict(motion_only=True, limited_disp=False, t0=1, optimize_intrinsics=True, focal_scale=0.98)),
    # second flow term from sparse tracks: 3 % of the pixels carry a (less noisy) track target with a large weight
    "pyba_c2_tracks": ("c2", {}, dict(motion_only=False, limited_disp=False, t0=1, sparse_tracks=True)),
}
STRIDE = 16


def _start_intrinsics(pr, flags):
    intr = (pr.intrinsics * 8.0)[None].clone()  # full resolution, buffer.py:413
    intr[:, :2] *= flags.get("focal_scale", 1.0)
    return intr


def case_inputs(name):
    """Inputs of a golden case in the PYTHON path's conventions, derived from the seeded synthetic problems."""
    from vipe_b200.synthetic import make_problem

    cfg_name, kw, flags = CASES[name]
    pr = make_problem(cfg_name, **kw)
    cfg = pr.cfg
    HW = cfg.ht * cfg.wd
    E = pr.ii.numel()
    target = pr.targets.reshape(E, 2, HW).permute(0, 2, 1).contiguous()  # channel-last, factor_graph.py:292-293
    weight = pr.weights.reshape(E, 2, HW).permute(0, 2, 1).contiguous()
    gen = torch.Generator().manual_seed(77)
    disp_damping = 0.01 * torch.nn.functional.softplus(torch.randn(cfg.n_frames, cfg.ht, cfg.wd, generator=gen))
    sparse = {}
    if flags.get("sparse_tracks"):
        g2 = torch.Generator().manual_seed(78)
        clean = make_problem(cfg_name, noise_px=0.0, **kw).targets.reshape(E, 2, HW).permute(0, 2, 1)
        sparse["sparse_target"] = (clean + 0.05 * torch.randn(E, HW, 2, generator=g2)).contiguous()
        hit = (torch.rand(E, HW, 1, generator=g2) < 0.03).float()
        sparse["sparse_weight"] = (hit * 5.0).expand(E, HW, 2).contiguous()
    return pr, dict(**sparse, poses=pr.poses.clone(), disps=pr.disps.clone(), disps_sens=pr.disps_sens.clone(),
                    intrinsics_full=_start_intrinsics(pr, flags), target=target, weight=weight,
                    disp_damping=disp_damping, ii=pr.ii.clone(), jj=pr.jj.clone(), t0=flags["t0"], t1=cfg.n_frames,
                    n_iters=cfg.iters, pose_damping=cfg.lm, pose_ep=cfg.ep, motion_only=flags["motion_only"],
                    limited_disp=flags["limited_disp"], alpha=0.001, ht=cfg.ht, wd=cfg.wd,
                    optimize_intrinsics=flags.get("optimize_intrinsics", False))


def main():
    assert REF.is_dir(), "needs /root/reference"
    install_stubs()
    out_dir = ROOT / "tests" / "golden"
    for name in CASES:
        pr, kw = case_inputs(name)
        python_bundle_adjustment(**kw)
        rec = {"poses": kw["poses"].numpy(), "disps_sub": kw["disps"].reshape(pr.cfg.n_frames, -1)[:, ::STRIDE].numpy(),
               "intrinsics_full": kw["intrinsics_full"].numpy()}
        np.savez_compressed(out_dir / f"{name}.npz", **rec)
        moved = float((kw["poses"] - pr.poses).abs().max())
        print(f"wrote {name}.npz  max pose change {moved:.4f}  disp change {float((kw['disps'] - pr.disps).abs().max()):.4f}")


if __name__ == "__main__":
    main()
