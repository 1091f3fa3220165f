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


@pytest.mark.parametrize("name", ["pyba_c1_full", "pyba_c2_full", "pyba_c1_motion", "pyba_c2_sensor_limited"])
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
    adapter.bundle_adjustment(poses, disps, kw["disps_sens"].to(dev), kw["intrinsics_full"][0].to(dev), kw["target"].to(dev),
                              kw["weight"].to(dev), kw["disp_damping"].to(dev), kw["ii"].to(dev), kw["jj"].to(dev), kw["t0"],
                              kw["t1"], kw["n_iters"], kw["pose_damping"], kw["pose_ep"], kw["motion_only"], kw["limited_disp"],
                              dense_disp_alpha=kw["alpha"])
    torch.cuda.synchronize()
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
        adapter.bundle_adjustment(x, x, x, x, x, x, x, x, x, 0, 1, 1, 1e-3, 0.1, False, False, optimize_intrinsics=True)
