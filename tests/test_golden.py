"""The oracle against golden vectors recorded from the reference's own CUDA `slam_ext.ba` on a B200
(tests/golden/make_golden.py).  The reference computes in fp32 with --use_fast_math, the oracle in fp64, so the
comparison uses the north-star tolerances (pose 1e-4, disparity 1e-3), not bit equality."""

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ba_oracle as O
from vipe_b200.synthetic import make_problem, pose_errors

GOLD = Path(__file__).resolve().parent / "golden"
STRIDE = 16
CASES = [("c1", False), ("c1", True), ("c2", False), ("c2", True)]


@pytest.mark.parametrize("name,motion_only", CASES)
def test_oracle_reproduces_reference_outputs(name, motion_only):
    f = GOLD / f"ref_{name}_{'motion' if motion_only else 'full'}.npz"
    if not f.is_file():
        pytest.skip(f"{f.name} not recorded yet")
    g = np.load(f)
    pr = make_problem(name)
    a = pr.args()
    a[14] = motion_only
    dx, dz = O.ba(*a, dtype=torch.float64)
    te, re_ = pose_errors(a[0], torch.from_numpy(g["poses"]), pr.t0, pr.t1)
    assert te <= 1e-4 and re_ <= 1e-4, (te, re_)
    assert torch.equal(a[0][: pr.t0], torch.from_numpy(g["poses"])[: pr.t0])
    gdx = torch.from_numpy(g["dx"]).double()
    assert (dx - gdx).norm() <= 1e-2 * gdx.norm() + 1e-7
    d = a[1].view(pr.cfg.n_frames, -1)[:, ::STRIDE].double()
    gd = torch.from_numpy(g["disps_sub"]).double()
    assert (d - gd).norm() <= 1e-3 * gd.norm()
    if not motion_only:
        gdz = torch.from_numpy(g["dz_sub"]).double()
        assert (dz[:, ::STRIDE] - gdz).norm() <= 1e-2 * gdz.norm()
