"""The adapter for vipe/slam's callers (vipe_b200/adapter.py) against golden vectors produced by the reference's own
PYTHON bundle adjustment (Solver.run_inplace etc., imported unmodified from /root/reference by
tests/golden/make_python_ba_golden.py).  Tolerances are the north-star ones; the Python path solves in fp32 LU."""

from pathlib import Path

import numpy as np
import pytest
import torch

from vipe_b200.synthetic import pose_errors

GOLD = Path(__file__).resolve().parent / "golden"

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["pyba_c1_full", "pyba_c2_full", "pyba_c1_motion", "pyba_c2_sensor_limited", "pyba_c2_focal",
                                  "pyba_c1_focal_motion", "pyba_c2_tracks"])
def test_adapter_matches_python_ba(lib_built, name):
    import sys

    sys.path.insert(0, str(GOLD))
    from make_python_ba_golden import STRIDE, case_inputs

    from vipe_b200 import adapter

    f = GOLD / f"{name}.npz"
    assert f.is_file(), f"{f} missing: run tests/golden/make_python_ba_golden.py where /root/reference exists"
    g = np.load(f)
    pr, kw = case_inputs(name)
    dev = torch.device("cuda:0")
    poses, disps = kw["poses"].to(dev), kw["disps"].to(dev)
    intr = kw["intrinsics_full"][0].to(dev)
    adapter.bundle_adjustment(poses, disps, kw["disps_sens"].to(dev), intr, kw["target"].to(dev),
                              kw["weight"].to(dev), kw["disp_damping"].to(dev), kw["ii"].to(dev), kw["jj"].to(dev), kw["t0"],
                              kw["t1"], kw["n_iters"], kw["pose_damping"], kw["pose_ep"], kw["motion_only"], kw["limited_disp"],
                              optimize_intrinsics=kw["optimize_intrinsics"], dense_disp_alpha=kw["alpha"],
                              sparse_target=kw["sparse_target"].to(dev) if "sparse_target" in kw else None,
                              sparse_weight=kw["sparse_weight"].to(dev) if "sparse_weight" in kw else None)
    torch.cuda.synchronize()
    gi = torch.from_numpy(g["intrinsics_full"])[0]
    if kw["optimize_intrinsics"]:  # the focal length moved by several pixels; it must land where the Python solver puts it
        assert float((gi[0] - kw["intrinsics_full"][0, 0]).abs()) > 1.0
        assert float((intr.cpu()[:2] - gi[:2]).abs().max()) <= 1e-4 * float(gi[0]), (intr.cpu(), gi)
        assert torch.equal(intr.cpu()[2:], gi[2:])
    else:
        assert torch.equal(intr.cpu(), gi)
    gp = torch.from_numpy(g["poses"])
    te, re_ = pose_errors(poses, gp, kw["t0"], kw["t1"])
    assert te <= 1e-4 and re_ <= 1e-4, (te, re_)
    assert torch.equal(poses[: kw["t0"]].cpu(), gp[: kw["t0"]])
    d = disps.cpu().reshape(pr.cfg.n_frames, -1)[:, ::STRIDE].double()
    gd = torch.from_numpy(g["disps_sub"]).double()
    assert (d - gd).norm() <= 1e-3 * gd.norm(), float((d - gd).norm() / gd.norm())
    if kw["motion_only"]:
        assert torch.equal(disps.cpu(), pr.disps.clamp(min=0.001))


def test_adapter_rejects_unsupported(lib_built):
    from vipe_b200 import adapter

    x = torch.zeros(1, device="cuda:0")
    with pytest.raises(NotImplementedError):
        adapter.bundle_adjustment(x, x, x, x, x, x, x, x, x, 0, 1, 1, 1e-3, 0.1, False, False, optimize_rig_rotation=True)


def test_focal_recovers_at_backend_size(lib_built):
    """C3-sized problem (no Python golden at this size; exercises the tiled Cholesky with the bordered system): starting
    3 % off, the focal length must come back to the value the targets were generated with, and the poses must still land
    on the ground truth.  The bounded "orbit" trajectory is used because the Python path accepts depths down to 0.1
    (cameras.py:48), which lets C3's random-walk loop closures (points almost in the camera plane) blow the step up."""
    import dataclasses

    from vipe_b200 import adapter
    from vipe_b200.synthetic import CONFIGS, make_problem

    pr = make_problem(dataclasses.replace(CONFIGS["c3"], trajectory="orbit"))
    cfg = pr.cfg
    dev = torch.device("cuda:0")
    E, HW = pr.ii.numel(), cfg.ht * cfg.wd
    target = pr.targets.reshape(E, 2, HW).permute(0, 2, 1).contiguous().to(dev)
    weight = pr.weights.reshape(E, 2, HW).permute(0, 2, 1).contiguous().to(dev)
    gt_focal = float(pr.intrinsics[0]) * 8.0
    intr = (pr.intrinsics * 8.0).to(dev)
    intr[:2] *= 1.03
    poses, disps = pr.poses.clone().to(dev), pr.disps.clone().to(dev)
    gen = torch.Generator().manual_seed(77)
    damp = (0.01 * torch.nn.functional.softplus(torch.randn(cfg.n_frames, cfg.ht, cfg.wd, generator=gen))).to(dev)  # droid_net.py:410
    adapter.bundle_adjustment(poses, disps, pr.disps_sens.to(dev), intr, target, weight, damp, pr.ii.to(dev), pr.jj.to(dev), 1,
                              cfg.n_frames, 8, cfg.lm, cfg.ep, False, False, optimize_intrinsics=True)
    torch.cuda.synchronize()
    assert abs(float(intr[0]) - gt_focal) < 2e-3 * gt_focal, (float(intr[0]), gt_focal)
    assert float(intr[0]) == float(intr[1])
    te, re_ = pose_errors(poses, pr.poses_gt.to(dev), 1, cfg.n_frames)
    te0, re0 = pose_errors(pr.poses.to(dev), pr.poses_gt.to(dev), 1, cfg.n_frames)
    assert te < 0.5 * te0 and re_ < 0.5 * re0, (te, te0, re_, re0)
