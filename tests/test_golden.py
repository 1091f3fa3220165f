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


def test_geom_oracle_reproduces_reference_outputs():
    """projmap / frame_distance / depth_filter / iproj: CPU oracle vs vectors recorded from the reference's CUDA run."""
    f = GOLD / "ref_geom_c2.npz"
    if not f.is_file():
        pytest.skip(f"{f.name} not recorded yet")
    from oracle import geom_oracle as G

    g = np.load(f)
    pr = make_problem("c2")
    coords, valid, z = G.projmap(pr.poses, pr.disps, pr.intrinsics, pr.ii, pr.jj)
    assert (coords.view(120, -1, 3)[:, ::STRIDE] - torch.from_numpy(g["coords_sub"]).double()).abs().max() < 2e-3
    assert (valid.view(120, -1)[:, ::STRIDE] != torch.from_numpy(g["valid_sub"]).double()).float().mean() < 1e-4
    K2 = torch.stack([pr.intrinsics, pr.intrinsics * 1.01])
    pi, pj, qi, qj = [torch.from_numpy(g[k]) for k in ("pi", "pj", "qi", "qj")]
    dist, ratio = G.frame_distance(pr.poses, pr.disps, K2, pi, pj, qi, qj, pi, 0.3)
    clear = (ratio - 0.75).abs() > 1e-3
    assert torch.allclose(dist[clear], torch.from_numpy(g["dist"]).double()[clear], rtol=2e-4, atol=1e-4)
    counter, margin = G.depth_filter(pr.poses_gt, pr.disps_gt, pr.intrinsics, torch.from_numpy(g["ix"]), torch.from_numpy(g["thresh"]))
    c, m = counter.view(6, -1)[:, ::STRIDE], margin.view(6, -1)[:, ::STRIDE]
    ok = m > 1e-4
    assert torch.equal(c[ok], torch.from_numpy(g["counter_sub"]).double()[ok])
    pts = G.iproj(pr.poses, pr.disps, pr.intrinsics)
    assert torch.allclose(pts.view(16, -1, 3)[:, ::STRIDE], torch.from_numpy(g["points_sub"]).double(), rtol=2e-5, atol=1e-5)
