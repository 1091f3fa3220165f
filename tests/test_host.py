"""CPU tests of the host side: the C-ABI library loads and exports what include/vipe_ba.h declares, the plan's
index bookkeeping is bit-exact against the oracle's (geom_kernels.cu:1301-1308, :946-981, :1209-1240), the keyframe
sharding covers every frame exactly once, and the operator mirrors the reference's error behaviour."""

import ctypes
import re
from pathlib import Path

import pytest
import torch

from oracle import ba_oracle as O

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(lib_built):
    header = (ROOT / "include" / "vipe_ba.h").read_text()
    declared = set(re.findall(r"\b(vipe_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 20
    L = ctypes.CDLL(str(lib_built))
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/vipe_ba.h but not exported"
    from vipe_b200 import _lib

    assert set(_lib.SIGNATURES) == declared, "ctypes table and header disagree"
    assert _lib.lib().vipe_ba_abi_version() == _lib.ABI_VERSION


def _random_graph(gen, n_frames, n_edges, stereo=False):
    ii = torch.randint(0, n_frames, (n_edges,), generator=gen)
    jj = torch.randint(0, n_frames, (n_edges,), generator=gen)
    if not stereo:
        jj = torch.where(ii == jj, (jj + 1) % n_frames, jj)
    return ii, jj


@pytest.mark.parametrize("seed", range(8))
def test_plan_bookkeeping_is_bit_exact(lib_built, seed):
    from vipe_b200.plan import BAPlan

    gen = torch.Generator().manual_seed(seed)
    n_frames = int(torch.randint(3, 14, (1,), generator=gen))
    n_edges = int(torch.randint(0, 40, (1,), generator=gen))
    ii, jj = _random_graph(gen, n_frames, n_edges, stereo=seed % 2 == 0)
    t0 = int(torch.randint(0, n_frames - 1, (1,), generator=gen))
    t1 = n_frames if seed % 3 else max(t0 + 1, n_frames - 2)
    keep = (ii < t1) & (jj < t1) if t1 < n_frames else torch.ones_like(ii, dtype=torch.bool)
    ii, jj = ii[keep], jj[keep]
    plan = BAPlan(ii, jj, n_frames, 6, 8, t0, t1)
    bk = O.bookkeeping(ii, jj, t0, t1)
    assert torch.equal(plan.kx, bk.kx)
    assert torch.equal(plan.kk_exp, bk.kk_exp)
    ptrs, idxs = plan.csr()
    rows = O.csr_by_source(ii, bk.kx)
    assert ptrs.tolist() == [0] + list(torch.tensor([len(r) for r in rows]).cumsum(0).tolist()) if rows else [0]
    for k, r in enumerate(rows):
        assert sorted(idxs[ptrs[k]: ptrs[k + 1]].tolist()) == r
    trip, _ = O.schur_triples(bk, t0, t1)
    assert plan.num_schur_triples == len(trip)
    assert plan.max_degree == max([len(r) for r in rows] + [0])


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_assignment_partitions_the_frames(lib_built, world, problems):
    from vipe_b200.plan import BAPlan

    pr = problems("c2")
    ranges = []
    for r in range(world):
        plan = BAPlan(pr.ii, pr.jj, 16, 48, 64, pr.t0, pr.t1, rank=r, world=world)
        lo, hi = plan.owned_range()
        assert (lo, hi) == plan.owned_range(r)
        ranges.append((lo, hi))
        # every rank computes the same partition
        assert [plan.owned_range(q) for q in range(world)] == [BAPlan(pr.ii, pr.jj, 16, 48, 64, pr.t0, pr.t1, rank=0, world=world).owned_range(q) for q in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == plan.K
    for a, b in zip(ranges[:-1], ranges[1:]):
        assert a[1] == b[0]
    # balanced by (out-degree + 1) within one frame's worth of work
    deg = torch.bincount(pr.ii, minlength=16)[plan.kx] + 1
    loads = [int(deg[lo:hi].sum()) for lo, hi in ranges]
    assert max(loads) - min(loads) <= 2 * int(deg.max())


def test_plan_rejects_bad_input(lib_built):
    from vipe_b200.plan import BAPlan

    with pytest.raises(RuntimeError, match="outside"):
        BAPlan(torch.tensor([0, 9]), torch.tensor([1, 0]), 4, 6, 8, 1, 4)
    with pytest.raises(RuntimeError, match="t0"):
        BAPlan(torch.tensor([0]), torch.tensor([1]), 4, 6, 8, 3, 2)
    with pytest.raises(RuntimeError):
        BAPlan(torch.tensor([0, 1]), torch.tensor([1]), 4, 6, 8, 1, 4)


def test_operator_error_behaviour_matches_reference(lib_built, problems):
    """CHECK_CONTIGUOUS -> RuntimeError (geom_kernels.cu:1287-1294); plus: no CPU fallback."""
    from vipe_b200.ext import slam_ext

    pr = problems("c1")
    a = pr.args()
    with pytest.raises(RuntimeError, match="CUDA"):
        slam_ext.ba(*a)
    a = pr.args()
    a[4] = a[4].permute(0, 1, 3, 2)
    with pytest.raises(RuntimeError, match="targets must be contiguous"):
        slam_ext.ba(*a)
    a = pr.args()
    a[7] = a[7].int()
    with pytest.raises(RuntimeError, match="scalar type"):
        slam_ext.ba(*a)


def test_no_oracle_import_in_product_path():
    """The shipped package must never import the oracle (it is test infrastructure)."""
    for p in (ROOT / "vipe_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), p


def test_synthetic_configs_match_baseline(problems):
    pr = problems("c2")
    assert pr.ii.numel() == 120 and pr.poses.shape == (16, 7) and pr.targets.shape == (120, 2, 48, 64)
    assert int((pr.ii - pr.jj).abs().max()) == 5
    pr1 = problems("c1")
    assert pr1.ii.numel() == 24 and int((pr1.ii - pr1.jj).abs().max()) == 2
    # deterministic
    from vipe_b200.synthetic import make_problem

    again = make_problem("c1")
    assert torch.equal(again.targets, pr1.targets) and torch.equal(again.poses, pr1.poses)


def test_host_feed_needs_cuda():
    """The host-memory feeder is CUDA-only like the operator itself (no CPU fallback)."""
    import pytest

    from vipe_b200.host_feed import HostFeed

    with pytest.raises(RuntimeError):
        HostFeed("cpu")


def test_identity_plan_cache_keeps_several_graphs(lib_built):
    """Callers that alternate between argument sets (HostFeed's two slots, frontend/backend graphs) must all hit the
    identity cache; an in-place edit of the edge list must miss it."""
    from vipe_b200.ext import slam_ext

    gen = torch.Generator().manual_seed(7)
    graphs = [_random_graph(gen, 9, 20) for _ in range(3)]
    plans = [slam_ext.ba_plan(ii, jj, 9, 8, 8, 1, 9) for ii, jj in graphs]
    for _ in range(2):
        for (ii, jj), p in zip(graphs, plans):
            assert slam_ext.ba_plan(ii, jj, 9, 8, 8, 1, 9) is p
    assert len(slam_ext._LAST_PLANS) <= slam_ext._LAST_PLANS_MAX
    ii, jj = graphs[0]
    keep = int(jj[0])
    jj[0] = (keep + 1) % 9 if (keep + 1) % 9 != int(ii[0]) else (keep + 2) % 9
    p2 = slam_ext.ba_plan(ii, jj, 9, 8, 8, 1, 9)
    assert p2 is not plans[0]
    # a shard's plan is its own entry
    assert slam_ext.ba_plan(ii, jj, 9, 8, 8, 1, 9, 1, 2) is not p2
    assert slam_ext.ba_plan(ii, jj, 9, 8, 8, 1, 9, 1, 2) is slam_ext.ba_plan(ii, jj, 9, 8, 8, 1, 9, 1, 2)


@pytest.mark.parametrize("deg", [40, 150, 300, 650])
def test_plan_accepts_hub_frames(lib_built, deg):
    """A source frame with far more outgoing edges than the shared-memory staging buffer holds (the reference has no out-degree
    limit) must not make plan creation fail: hub frames get a launch of their own that stages through global memory."""
    from vipe_b200.plan import BAPlan

    n = deg + 1
    others = torch.arange(1, n)
    ii = torch.cat([torch.zeros(deg, dtype=torch.int64), others])
    jj = torch.cat([others, torch.zeros(deg, dtype=torch.int64)])
    plan = BAPlan(ii, jj, n, 8, 16, 1, n)
    assert plan.P == n - 1
    bk = O.bookkeeping(ii, jj, 1, n)
    assert torch.equal(plan.kx, bk.kx)
    assert int(_max_degree(plan)) == deg


def _max_degree(plan):
    from vipe_b200 import _lib

    return _lib.lib().vipe_ba_plan_max_degree(plan.handle)
