"""Pins the oracle against the REFERENCE ITSELF: csrc/slam_ext/{geom_kernels.cu,slam.cpp} compiled unmodified by
oracle/build_ref.py (Eigen replaced by oracle/eigen_stub) and executed on the GPU.  The reference is fp32 with
--use_fast_math (vipe/ext/specs.py:39) and sums 256-wide in fp32, so agreement with the fp64 oracle is ~1e-5, not
bit-for-bit.  Skipped when oracle/_ref/vipe_ref_ext.so was not built."""

import pytest
import torch

from oracle import ba_oracle as O
from oracle import build_ref
from vipe_b200.synthetic import disp_error, make_problem, pose_errors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_mod():
    m = build_ref.load()
    if m is None:
        pytest.skip("oracle/_ref/vipe_ref_ext.so not built (needs /root/reference at build time)")
    return m


@pytest.mark.parametrize("name,motion_only", [("c1", False), ("c2", False), ("c1", True), ("c2", True)])
def test_oracle_matches_reference_run(ref_mod, name, motion_only):
    pr = make_problem(name)
    dev = torch.device("cuda:0")
    o = pr.args()
    o[14] = motion_only
    tr = O.Trace()
    dxo, dzo = O.ba(*o, dtype=torch.float64, trace=tr)
    a = pr.args(dev)
    a[14] = motion_only
    out = ref_mod.slam_ext.ba(*a)
    torch.cuda.synchronize()
    te, re_ = pose_errors(a[0], o[0], pr.t0, pr.t1)
    assert te <= 1e-4 and re_ <= 1e-4, (te, re_)
    assert (out[0].cpu().double() - dxo).norm() <= 1e-2 * dxo.norm() + 1e-7
    if not motion_only:
        assert disp_error(a[1], o[1], tr.bk.kx) <= 1e-3
        assert (out[1].cpu().double() - dzo).norm() <= 1e-2 * dzo.norm()


def test_ours_matches_reference_run(ref_mod, lib_built):
    from vipe_b200.ext import slam_ext

    pr = make_problem("c2")
    dev = torch.device("cuda:0")
    a, b = pr.args(dev), pr.args(dev)
    ref_mod.slam_ext.ba(*a)
    slam_ext.ba(*b)
    torch.cuda.synchronize()
    te, re_ = pose_errors(b[0], a[0], pr.t0, pr.t1)
    kx = torch.unique(torch.cat([torch.arange(pr.t0, pr.t1), pr.ii]))
    assert te <= 1e-4 and re_ <= 1e-4 and disp_error(b[1], a[1], kx) <= 1e-3
