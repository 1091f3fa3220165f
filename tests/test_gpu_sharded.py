"""Keyframe-sharded CUDA path over NCCL (needs >= 2 GPUs on the box; skipped otherwise): the sharded result must
match the single-GPU result of the same operator to fp32 rounding."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, name, out, collective, owned=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from vipe_b200.distributed import ba_sharded
    from vipe_b200.synthetic import make_problem

    pr = make_problem(name)
    a = pr.args(dev)
    prof = {}
    kw = {}
    if owned:  # this rank holds only the targets/weights rows of its own edges
        from vipe_b200.ext import slam_ext

        cfg = pr.cfg
        plan = slam_ext.ba_plan(pr.ii, pr.jj, cfg.n_frames, cfg.ht, cfg.wd, pr.t0, pr.t1, rank, world)
        own = plan.owned_edges()
        a[4], a[5] = pr.targets[own].contiguous().to(dev), pr.weights[own].contiguous().to(dev)
        kw = dict(plan=plan, owned_inputs=True)
    dx, dz = ba_sharded(*a, collective=collective, profile=prof, **kw)
    assert prof["collective"] == collective
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"poses": a[0].cpu(), "disps": a[1].cpu(), "dx": dx.cpu(), "dz": dz.cpu()}, out)
    chk = torch.cat([a[0].reshape(-1), a[1].reshape(-1)]).clone()
    ref = chk.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(chk, ref), "ranks disagree on the final state"
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("name,collective,owned", [("c2", "allreduce", False), ("c3", "allreduce", True), ("c3", "nvls", False),
                                                   ("c3", "nvls2", True), ("c3", "dist", True)])
def test_sharded_matches_single_gpu(lib_built, tmp_path, name, collective, owned):
    """"nvls": no all-reduce launch; the Cholesky kernel reads the sum of the ranks' partial systems through the NVSwitch
    (multimem.ld_reduce), see vipe_b200/distributed.py PeerSystem.  "nvls2": the in-switch sum as a reduce-scatter +
    multicast kernel of its own, then a local solve.  `owned`: every rank is given only its own edges' targets/weights."""
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 4)
    from vipe_b200.ext import slam_ext
    from vipe_b200.synthetic import disp_error, make_problem, pose_errors

    out = tmp_path / "r0.pt"
    mp.spawn(_worker, args=(world, _free_port(), name, str(out), collective, owned), nprocs=world, join=True)
    got = torch.load(out)
    pr = make_problem(name)
    a = pr.args(torch.device("cuda:0"))
    slam_ext.ba(*a)
    torch.cuda.synchronize()
    te, re_ = pose_errors(got["poses"], a[0], pr.t0, pr.t1)
    kx = torch.unique(torch.cat([torch.arange(pr.t0, pr.t1), pr.ii]))
    de = disp_error(got["disps"], a[1], kx)
    assert te <= 1e-5 and re_ <= 1e-5 and de <= 1e-4, (te, re_, de)


def _fail_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from vipe_b200.distributed import ba_sharded
    from vipe_b200.synthetic import BAConfig, make_problem

    pr = make_problem(BAConfig("tiled_fail", 41, 48, 300, 24, 32, 2, 1e-4, 0.1))
    a = pr.args(dev)
    a[5][3, 0, 0, 0] = float("nan")  # poisons the reduced system: the factorisation fails on every rank's copy
    a[14] = True
    prof = {}
    dx, _ = ba_sharded(*a, collective="dist", profile=prof)
    assert prof["collective"] == "dist"
    torch.cuda.synchronize()
    assert torch.equal(dx.cpu(), torch.zeros(pr.t1 - pr.t0, 6)), "a failed factorisation must give a zero update on every rank"
    assert torch.equal(a[0].cpu(), pr.poses)
    # and the next, clean call on the same plan / symmetric buffers works (flags carry epochs, nothing is left behind)
    b = pr.args(dev)
    b[14] = True
    dx2, _ = ba_sharded(*b, collective="dist")
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"poses": b[0].cpu(), "dx": dx2.cpu()}, out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_distributed_solve_failure_and_recovery(lib_built, tmp_path):
    """The distributed factorisation with a NaN in the system: every word it multicasts validates itself against ZERO, so NaN
    must travel like any other value (nothing may spin on it), the failure flag must reach every rank (dx = 0 everywhere,
    geom_kernels.cu:1186-1188), and the following clean call must match the single-GPU result."""
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 4)
    from vipe_b200.ext import slam_ext
    from vipe_b200.synthetic import BAConfig, make_problem, pose_errors

    out = tmp_path / "r0.pt"
    mp.spawn(_fail_worker, args=(world, _free_port(), str(out)), nprocs=world, join=True)
    got = torch.load(out)
    pr = make_problem(BAConfig("tiled_fail", 41, 48, 300, 24, 32, 2, 1e-4, 0.1))
    a = pr.args(torch.device("cuda:0"))
    a[14] = True
    slam_ext.ba(*a)
    torch.cuda.synchronize()
    te, re_ = pose_errors(got["poses"], a[0], pr.t0, pr.t1)
    assert te <= 1e-5 and re_ <= 1e-5, (te, re_)
