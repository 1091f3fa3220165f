"""world_size-2 `gloo` test of the keyframe-sharded driver (vipe_b200/distributed.py) on CPU.

The orchestration under test is the product's: plan sharding (C ABI), one all-reduce of the reduced camera system
per iteration, replicated solve, owner-local back-substitution, one exchange of owned disparity rows.  Only the
engine between the collectives is swapped: instead of the CUDA kernels it is the fp64 oracle restricted to the
rank's owned source frames.  The sharded result must equal the single-process oracle."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ba_oracle as O


class OracleShardEngine:
    def __init__(self, plan, poses, disps, intrinsics, disps_sens, targets, weights, eta, motion_only):
        self.plan, self.motion_only = plan, bool(motion_only)
        self.poses, self.disps = poses, disps
        f64 = torch.float64
        self.intr, self.dsens = intrinsics.to(f64), disps_sens.to(f64)
        self.tg, self.wt = targets.to(f64), weights.to(f64)
        self.HW = plan.ht * plan.wd
        self.eta = eta.to(f64).reshape(-1, self.HW)
        self.lo, self.hi = plan.owned_range()
        self.kx = plan.kx
        ptrs, idxs = plan.csr()
        self.own_edges = idxs[ptrs[self.lo]: ptrs[self.hi]].sort().values
        self.n = 6 * plan.P
        self.system = torch.zeros(self.n * self.n + self.n, dtype=f64)
        self.dx = torch.zeros(plan.P, 6, dtype=torch.float32)
        self.dz = torch.zeros(plan.K, self.HW, dtype=torch.float32)

    def _state(self):
        return self.poses.to(torch.float64), self.disps.to(torch.float64)

    def linearize(self):
        pl, (p, d) = self.plan, self._state()
        t0, t1, P = pl.t0, pl.t1, pl.P
        oe = self.own_edges
        self.ii, self.jj = self._edges()
        ii, jj = self.ii[oe], self.jj[oe]
        Hs, vs, Eii, Eij, Cii, bz = O.linearize(p, d, self.intr, self.tg[oe], self.wt[oe], ii, jj)
        A, b = O.assemble_pose_system(Hs, vs, ii, jj, t0, t1)
        if not self.motion_only:
            own_kx = self.kx[self.lo: self.hi]
            m = (self.dsens[own_kx] > 0).to(torch.float64).reshape(-1, self.HW)
            C = O.accum(Cii, ii, own_kx) + m * O.ALPHA + (1 - m) * self.eta[self.lo: self.hi]
            w = O.accum(bz, ii, own_kx) - m * O.ALPHA * (d[own_kx] - self.dsens[own_kx]).reshape(-1, self.HW)
            Q = 1.0 / C
            # rows: one Ei row per owned frame inside the window, one Eij row per owned edge
            rows_E, rows_pose, rows_k = [], [], []
            for kpos, f in enumerate(own_kx.tolist()):
                if t0 <= f < t1:
                    rows_E.append(Eii[ii == f].sum(0) if (ii == f).any() else torch.zeros(6, self.HW, dtype=torch.float64))
                    rows_pose.append(f - t0)
                    rows_k.append(kpos)
            pos = torch.searchsorted(own_kx, ii)
            for e in range(ii.numel()):
                rows_E.append(Eij[e])
                rows_pose.append(int(jj[e]) - t0)
                rows_k.append(int(pos[e]))
            Eall = torch.stack(rows_E) if rows_E else torch.zeros(0, 6, self.HW, dtype=torch.float64)
            rp, rk = torch.tensor(rows_pose, dtype=torch.int64), torch.tensor(rows_k, dtype=torch.int64)
            ok = (rp >= 0) & (rp < P)
            S = torch.zeros(P, 6, P, 6, dtype=torch.float64)
            sv = torch.zeros(P, 6, dtype=torch.float64)
            for kpos in range(own_kx.numel()):
                r = torch.nonzero((rk == kpos) & ok).flatten()
                if r.numel() == 0:
                    continue
                Ek = Eall[r]
                Sk = torch.einsum("aip,bjp->aibj", Ek * Q[kpos], Ek)
                pa = rp[r]
                S.index_put_((pa[:, None, None, None], torch.arange(6)[None, :, None, None], pa[None, None, :, None],
                              torch.arange(6)[None, None, None, :]), Sk, accumulate=True)
                sv.index_add_(0, pa, torch.einsum("p,nip->ni", Q[kpos] * w[kpos], Ek))
            A = A - S.reshape(self.n, self.n)
            b = b - sv.reshape(-1)
            self._back = (Eall, rp, rk, Q, w)
        self.system[: self.n * self.n] = A.reshape(-1)
        self.system[self.n * self.n:] = b
        return self.system

    def _edges(self):
        return self._ii, self._jj

    def solve_update(self, lm, ep):
        pl = self.plan
        P, t0, t1 = pl.P, pl.t0, pl.t1
        H = self.system[: self.n * self.n].reshape(self.n, self.n)
        x, _ = O.solve_damped(H, self.system[self.n * self.n:], lm, ep)
        dx = x.reshape(P, 6)
        p, d = self._state()
        if not self.motion_only:
            Eall, rp, rk, Q, w = self._back
            okb = (rp > 0) & (rp < P)  # Q4
            dxe = torch.zeros(Eall.shape[0], 6, dtype=torch.float64)
            dxe[okb] = dx[rp[okb]]
            dw = torch.einsum("nip,ni->np", Eall, dxe)
            acc = torch.zeros_like(Q)
            acc.index_add_(0, rk, dw)
            dz = Q * (w - acc)
            own_kx = self.kx[self.lo: self.hi]
            d[own_kx] = d[own_kx] + dz.reshape(-1, pl.ht, pl.wd)
            self.dz.zero_()
            self.dz[self.lo: self.hi] = dz.float()
            self.disps.copy_(d.to(self.disps.dtype))
        tn, qn = O.retr_se3(dx, p[t0:t1, :3], p[t0:t1, 3:])
        p[t0:t1, :3], p[t0:t1, 3:] = tn, qn
        self.poses.copy_(p.to(self.poses.dtype))
        self.dx.copy_(dx.float())


def _worker(rank, world, port, name, motion_only, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from vipe_b200.distributed import ba_sharded
    from vipe_b200.synthetic import make_problem

    pr = make_problem(name)
    a = pr.args()
    a[0], a[1] = a[0].double(), a[1].double()  # keep the state in fp64 so the comparison is tight
    a[14] = motion_only

    class Engine(OracleShardEngine):
        _ii, _jj = a[7], a[8]

    dx, dz = ba_sharded(*a, engine_cls=Engine)
    if rank == 0:
        torch.save({"poses": a[0], "disps": a[1], "dx": dx, "dz": dz}, out)
    # every rank must hold the same state at the end
    chk = torch.cat([a[0].reshape(-1), a[1].reshape(-1)]).clone()
    ref = chk.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(chk, ref), "ranks disagree on the final state"
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("name,motion_only", [("c1", False), ("c2", False), ("c1", True)])
def test_sharded_equals_single_process_oracle(lib_built, tmp_path, name, motion_only):
    from vipe_b200.synthetic import make_problem

    out = tmp_path / "r0.pt"
    mp.spawn(_worker, args=(2, _free_port(), name, motion_only, str(out)), nprocs=2, join=True)
    got = torch.load(out)
    pr = make_problem(name)
    a = pr.args()
    a[0], a[1] = a[0].double(), a[1].double()
    a[14] = motion_only
    dx, dz = O.ba(*a, dtype=torch.float64)
    assert torch.allclose(got["poses"], a[0], rtol=0, atol=1e-9)
    assert torch.allclose(got["disps"], a[1], rtol=0, atol=1e-9)
    assert torch.allclose(got["dx"].double(), dx, rtol=0, atol=1e-6)
    if not motion_only:
        assert torch.allclose(got["dz"].double(), dz, rtol=0, atol=1e-6)
