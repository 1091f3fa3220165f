"""GPU parity of the other slam_ext operators (projmap, frame_distance, depth_filter, iproj) against the CPU oracle
(oracle/geom_oracle.py) and against the reference's own CUDA run (oracle/_ref), plus CPU self-checks of the oracle."""

import pytest
import torch

from oracle import build_ref
from oracle import geom_oracle as G
from vipe_b200.synthetic import make_problem


@pytest.fixture(scope="module")
def pr():
    return make_problem("c2")


def test_oracle_projmap_agrees_with_ba_geometry(pr):
    """CPU: projmap's reprojection must be the point where the BA residual vanishes (same transform, :295 vs :507)."""
    from oracle import ba_oracle as O

    coords, valid, z = G.projmap(pr.poses, pr.disps, pr.intrinsics, pr.ii, pr.jj)
    a = pr.args()
    a[4] = coords[..., :2].permute(0, 3, 1, 2).contiguous().float()  # targets := reprojection
    Hs, vs, Eii, Eij, Cii, bz = O.linearize(a[0].double(), a[1].double(), a[2].double(), a[4].double(), a[5].double(), pr.ii, pr.jj)
    assert vs.abs().max() < 1e-3 and bz.abs().max() < 1e-4  # residuals ~ fp32 rounding of the stored targets
    assert valid.min() == 1.0


def test_oracle_iproj_inverts_projection(pr):
    pts = G.iproj(pr.poses, pr.disps, pr.intrinsics)
    # pose 0 is the identity in C2: points are ((u-cx)/fx, (v-cy)/fy, 1) / d
    d = pr.disps[0].double()
    assert torch.allclose(pts[0, ..., 2], 1.0 / d, rtol=1e-12)


gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops(lib_built):
    from vipe_b200.ext import slam_ext

    return slam_ext


@pytest.fixture(scope="module")
def ref_ops():
    m = build_ref.load()
    return m.slam_ext if m is not None else None


@gpu
def test_projmap(ops, ref_ops, dev, pr):
    p, d, k, ii, jj = [x.to(dev) for x in (pr.poses, pr.disps, pr.intrinsics, pr.ii, pr.jj)]
    coords, valid = ops.projmap(p, d, k, ii, jj)
    oc, ov, z = G.projmap(pr.poses, pr.disps, pr.intrinsics, pr.ii, pr.jj)
    assert coords.shape == (120, 48, 64, 3) and valid.shape == (120, 48, 64, 1)
    assert (coords.cpu().double() - oc).abs().max() < 2e-3  # pixels; fp32 projection of coordinates up to ~100
    clear = (z - O_MIN).abs() > 1e-5
    assert torch.equal(valid.cpu()[..., 0][clear].double(), ov[..., 0][clear])
    assert torch.equal(coords[..., 2].cpu(), torch.zeros(120, 48, 64))
    if ref_ops is not None:
        rc, rv = ref_ops.projmap(p, d, k, ii, jj)
        assert (coords - rc).abs().max() < 2e-3
        assert (valid != rv).float().mean() < 1e-4


O_MIN = 0.25


@gpu
def test_frame_distance(ops, ref_ops, dev, pr):
    p, d = pr.poses.to(dev), pr.disps.to(dev)
    K = torch.stack([pr.intrinsics, pr.intrinsics * 1.01]).to(dev)
    gen = torch.Generator().manual_seed(3)
    M = 40
    pi = torch.randint(0, 16, (M,), generator=gen)
    pj = torch.randint(0, 16, (M,), generator=gen)
    qi = torch.randint(0, 2, (M,), generator=gen)
    qj = torch.randint(0, 2, (M,), generator=gen)
    di = pi.clone()
    for beta in (0.3, 1.0):
        dist = ops.frame_distance(p, d, K, pi.to(dev), pj.to(dev), qi.to(dev), qj.to(dev), di.to(dev), beta)
        od, ratio = G.frame_distance(pr.poses, pr.disps, K.cpu(), pi, pj, qi, qj, di, beta)
        clear = (ratio - 0.75).abs() > 1e-3
        assert torch.allclose(dist.cpu().double()[clear], od[clear], rtol=2e-4, atol=1e-4)
        if ref_ops is not None:
            rd = ref_ops.frame_distance(p, d, K, pi.to(dev), pj.to(dev), qi.to(dev), qj.to(dev), di.to(dev), beta)
            assert torch.allclose(dist[clear.to(dev)], rd[clear.to(dev)], rtol=2e-4, atol=1e-4)


@gpu
def test_depth_filter(ops, ref_ops, dev, pr):
    p, d, k = pr.poses_gt.to(dev), pr.disps_gt.to(dev), pr.intrinsics.to(dev)
    ix = torch.tensor([0, 1, 5, 8, 14, 15])
    thresh = torch.tensor([0.05, 0.1, 0.2, 0.4, 0.8, 1.6])
    counter = ops.depth_filter(p, d, k, ix.to(dev), thresh.to(dev))
    oc, margin = G.depth_filter(pr.poses_gt, pr.disps_gt, pr.intrinsics, ix, thresh)
    assert counter.shape == (6, 48, 64)
    clear = margin > 1e-4
    assert clear.float().mean() > 0.95
    assert torch.equal(counter.cpu().double()[clear], oc[clear])
    assert counter.max() <= 6 and counter.min() >= 0 and counter.sum() > 0
    if ref_ops is not None:
        rc = ref_ops.depth_filter(p, d, k, ix.to(dev), thresh.to(dev))
        assert (counter != rc).float().mean() < 2e-3


@gpu
def test_iproj(ops, ref_ops, dev, pr):
    p, d, k = pr.poses.to(dev), pr.disps.to(dev), pr.intrinsics.to(dev)
    pts = ops.iproj(p, d, k)
    op = G.iproj(pr.poses, pr.disps, pr.intrinsics)
    assert pts.shape == (16, 48, 64, 3)
    assert torch.allclose(pts.cpu().double(), op, rtol=2e-5, atol=1e-5)
    if ref_ops is not None:
        rp = ref_ops.iproj(p, d, k)
        assert torch.allclose(pts, rp, rtol=2e-5, atol=1e-5)


@gpu
def test_ops_reject_cpu_and_noncontiguous(ops, pr):
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.iproj(pr.poses, pr.disps, pr.intrinsics)
