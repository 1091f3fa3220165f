"""Deterministic synthetic dense-BA problems (BASELINE.json `configs`).

Every tensor is produced on the CPU from `torch.Generator().manual_seed(20251018 + cfg)` so that the
CPU oracle and the GPU path see identical bits; callers copy to the device afterwards.  Layouts are
the ones `slam_ext.ba` takes (/root/reference/csrc/slam_ext/geom_kernels.cu:1283-1299):
poses[N,7] = (tx,ty,tz,qx,qy,qz,qw) world->camera, disps[N,ht,wd], intrinsics[4], disps_sens[N,ht,wd],
targets/weights[E,2,ht,wd] channel-first, eta[K,ht,wd], ii/jj[E] int64.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

BASE_SEED = 20251018


@dataclass
class BAConfig:
    name: str
    cfg_id: int
    n_frames: int
    n_edges: int
    ht: int
    wd: int
    iters: int
    lm: float
    ep: float
    motion_only: bool = False
    closure_frac: float = 0.0  # share of edges that are long-range loop closures
    max_delta: int | None = None
    clips: int = 1  # C5: number of independent clips of this shape
    trajectory: str = "walk"  # "walk": random walk; "orbit": bounded Lissajous motion, every pair covisible


CONFIGS = {
    "c1": BAConfig("c1", 1, 8, 24, 48, 64, 2, 1e-4, 0.1),
    "c2": BAConfig("c2", 2, 16, 120, 48, 64, 4, 1e-3, 0.1),
    "c3": BAConfig("c3", 3, 300, 3000, 48, 64, 8, 1e-5, 1e-2, closure_frac=0.1),
    # C4 uses the bounded trajectory: with a 1000-frame random walk the "random pairs with |i-j| > 16" of SURVEY 8(d)
    # connect frames that see nothing in common and the Gauss-Newton step explodes (in the reference too)
    "c4": BAConfig("c4", 4, 1000, 12000, 64, 112, 8, 1e-5, 1e-2, closure_frac=0.1, trajectory="orbit"),
    "c5": BAConfig("c5", 5, 16, 120, 48, 64, 4, 1e-3, 0.1, motion_only=True, clips=64),
}


@dataclass
class BAProblem:
    cfg: BAConfig
    poses: torch.Tensor
    disps: torch.Tensor
    intrinsics: torch.Tensor
    disps_sens: torch.Tensor
    targets: torch.Tensor
    weights: torch.Tensor
    eta: torch.Tensor
    ii: torch.Tensor
    jj: torch.Tensor
    t0: int
    t1: int
    poses_gt: torch.Tensor
    disps_gt: torch.Tensor

    def args(self, device=None):
        """Positional argument list of slam_ext.ba; poses/disps are fresh clones (they are mutated)."""
        def mv(x):
            return x.clone().to(device) if device is not None else x.clone()

        c = self.cfg
        return [mv(self.poses), mv(self.disps), mv(self.intrinsics), mv(self.disps_sens), mv(self.targets),
                mv(self.weights), mv(self.eta), mv(self.ii), mv(self.jj), self.t0, self.t1, c.iters, c.lm, c.ep,
                c.motion_only]

    @property
    def edge_pixels(self) -> int:
        return int(self.ii.numel()) * self.cfg.ht * self.cfg.wd


# ----------------------------------------------------------------------------- small SE3 toolkit (fp64)
def _quat_rotate(q, X):
    qv, qw = q[..., :3], q[..., 3:4]
    uv = 2.0 * torch.linalg.cross(qv.expand_as(X), X, dim=-1)
    return X + qw * uv + torch.linalg.cross(qv.expand_as(X), uv, dim=-1)


def _quat_mul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by,
                        aw * by + ay * bw + az * bx - ax * bz,
                        aw * bz + az * bw + ax * by - ay * bx,
                        aw * bw - ax * bx - ay * by - az * bz], dim=-1)


def _se3_exp(xi):
    tau, phi = xi[..., :3], xi[..., 3:]
    th2 = (phi * phi).sum(-1, keepdim=True).clamp_min(1e-30)
    th = th2.sqrt()
    q = torch.cat([torch.sin(0.5 * th) / th * phi, torch.cos(0.5 * th)], dim=-1)
    a = (1 - torch.cos(th)) / th2
    b = (th - torch.sin(th)) / (th * th2)
    c1 = torch.linalg.cross(phi, tau, dim=-1)
    c2 = torch.linalg.cross(phi, c1, dim=-1)
    return tau + a * c1 + b * c2, q


def _compose_left(xi, t, q):
    """exp(xi) * (t, q)."""
    dt, dq = _se3_exp(xi)
    return _quat_rotate(dq, t) + dt, _quat_mul(dq, q)


# ----------------------------------------------------------------------------- edges
def banded_edges(n_frames: int, n_edges: int):
    """(i,i+d),(i+d,i) for d = 1,2,... and i ascending, truncated to n_edges."""
    ii, jj = [], []
    d = 1
    while len(ii) < n_edges and d < n_frames:
        for i in range(n_frames - d):
            ii += [i, i + d]
            jj += [i + d, i]
            if len(ii) >= n_edges:
                break
        d += 1
    return ii[:n_edges], jj[:n_edges]


def make_edges(cfg: BAConfig, gen: torch.Generator):
    n_close = int(round(cfg.n_edges * cfg.closure_frac))
    n_close -= n_close % 2
    ii, jj = banded_edges(cfg.n_frames, cfg.n_edges - n_close)
    have = set(zip(ii, jj))
    while n_close > 0:
        a, b = torch.randint(0, cfg.n_frames, (2,), generator=gen).tolist()
        if abs(a - b) <= 16 or (a, b) in have:
            continue
        have.add((a, b))
        have.add((b, a))
        ii += [a, b]
        jj += [b, a]
        n_close -= 2
    return torch.tensor(ii, dtype=torch.int64), torch.tensor(jj, dtype=torch.int64)


# ----------------------------------------------------------------------------- generator
def make_problem(cfg: BAConfig | str, clip: int = 0, sensor_on_even_frames: bool = False, noise_px: float = 0.25,
                 perturb: float = 1.0) -> BAProblem:
    """`noise_px`: std of the target noise; `perturb`: scale of the initial pose/disparity perturbation (0 = start at GT)."""
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    gen = torch.Generator().manual_seed(BASE_SEED + cfg.cfg_id + 1000 * clip)
    N, ht, wd = cfg.n_frames, cfg.ht, cfg.wd
    HW = ht * wd
    f64 = torch.float64

    def randn(*shape, dtype=f64):
        return torch.randn(*shape, generator=gen, dtype=dtype)

    intr = torch.tensor([0.94 * wd, 0.94 * wd, wd / 2.0, ht / 2.0], dtype=f64)

    # ground-truth trajectory
    xi = torch.cat([randn(N - 1, 3) * 0.05 + torch.tensor([0.05, 0.0, 0.0], dtype=f64), randn(N - 1, 3) * 0.02], dim=-1)
    t = torch.zeros(N, 3, dtype=f64)
    q = torch.zeros(N, 4, dtype=f64)
    q[:, 3] = 1.0
    if cfg.trajectory == "orbit":
        amp = torch.tensor([0.8, 0.8, 0.4, 0.12, 0.12, 0.12], dtype=f64)
        period = 80.0 + 80.0 * torch.rand(6, generator=gen, dtype=f64)
        phase = 6.283185307179586 * torch.rand(6, generator=gen, dtype=f64)
        nn_ = torch.arange(N, dtype=f64)[:, None]
        t, q = _se3_exp(amp * torch.sin(6.283185307179586 * nn_ / period + phase))
    else:
        for n in range(N - 1):
            t[n + 1], q[n + 1] = _compose_left(xi[n], t[n], q[n])
    poses_gt = torch.cat([t, q], dim=-1)

    # ground-truth disparity
    coarse = randn(N, 1, 6, 8, dtype=torch.float32)
    disps_gt = 0.2 + 0.8 * torch.sigmoid(F.interpolate(coarse, size=(ht, wd), mode="bilinear", align_corners=False))[:, 0]
    disps_gt = disps_gt.to(f64)

    ii, jj = make_edges(cfg, gen)
    E = ii.numel()

    # targets = GT reprojection + noise (chunked over edges to bound memory)
    v, u = torch.meshgrid(torch.arange(ht, dtype=f64), torch.arange(wd, dtype=f64), indexing="ij")
    Xi = torch.stack([(u.reshape(HW) - intr[2]) / intr[0], (v.reshape(HW) - intr[3]) / intr[1], torch.ones(HW, dtype=f64)], -1)
    targets = torch.empty(E, 2, ht, wd, dtype=torch.float32)
    for s in range(0, E, 256):
        ie, je = ii[s:s + 256], jj[s:s + 256]
        # Tij = Tj * Ti^-1
        qi_inv = torch.cat([-q[ie, :3], q[ie, 3:]], dim=-1)
        qij = _quat_mul(q[je], qi_inv)
        tij = t[je] - _quat_rotate(qij, t[ie])
        h = disps_gt.reshape(N, HW)[ie]
        Xj = _quat_rotate(qij[:, None, :], Xi[None].expand(ie.numel(), HW, 3)) + h[..., None] * tij[:, None, :]
        z = Xj[..., 2].clamp_min(1e-3)
        targets[s:s + 256, 0] = (intr[0] * Xj[..., 0] / z + intr[2]).reshape(-1, ht, wd).float()
        targets[s:s + 256, 1] = (intr[1] * Xj[..., 1] / z + intr[3]).reshape(-1, ht, wd).float()
    targets += noise_px * randn(E, 2, ht, wd, dtype=torch.float32)

    weights = torch.sigmoid(randn(E, 2, ht, wd, dtype=torch.float32))
    drop = torch.rand(E, 1, ht, wd, generator=gen) < 0.1
    weights = weights * (~drop)

    t0, t1 = 1, N
    K = int(torch.unique(torch.cat([torch.arange(t0, t1), ii])).numel())
    eta = (0.2 * 0.01 * F.softplus(randn(K, ht, wd, dtype=torch.float32)) + 1e-7).float()

    disps_sens = torch.zeros(N, ht, wd, dtype=torch.float32)
    if sensor_on_even_frames:
        disps_sens[0::2] = disps_gt[0::2].float()

    # initial state = perturbed ground truth
    pt, pq = _compose_left(randn(N, 6) * (0.01 * perturb), t, q)
    poses = torch.cat([pt, pq], dim=-1)
    poses[0] = poses_gt[0]
    disps = (disps_gt * (1.0 + (0.05 * perturb) * randn(N, ht, wd))).clamp_min(0.01)

    return BAProblem(cfg, poses.float().contiguous(), disps.float().contiguous(), intr.float(), disps_sens,
                     targets.contiguous(), weights.contiguous(), eta.contiguous(), ii, jj, t0, t1,
                     poses_gt.float(), disps_gt.float())


def pose_errors(p, p_ref, t0, t1):
    """(relative translation error over [t0,t1), max rotation geodesic angle in rad)."""
    a, b = p[t0:t1].double().cpu(), p_ref[t0:t1].double().cpu()
    terr = float((a[:, :3] - b[:, :3]).norm() / b[:, :3].norm().clamp_min(1e-30))
    qa = a[:, 3:] / a[:, 3:].norm(dim=-1, keepdim=True)
    qb = b[:, 3:] / b[:, 3:].norm(dim=-1, keepdim=True)
    dot = (qa * qb).sum(-1).abs().clamp(max=1.0)
    # 2*acos(dot) loses precision near 1: use the sine of the half angle instead
    s = (qa - qb * torch.sign((qa * qb).sum(-1, keepdim=True))).norm(dim=-1)
    ang = 2.0 * torch.asin((0.5 * s).clamp(max=1.0))
    del dot
    return terr, float(ang.max())


def disp_error(d, d_ref, kx):
    a, b = d[kx].double().cpu(), d_ref[kx].double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def disp_error_p999(d, d_ref, kx):
    """99.9th percentile over the pixels of the frames in `kx` of |d - d_ref| / |d_ref| (reported
    beside the Frobenius figure, which a handful of bad pixels could hide behind)."""
    a, b = d[kx].double().cpu().flatten(), d_ref[kx].double().cpu().flatten()
    rel = (a - b).abs() / b.abs().clamp_min(1e-12)
    k = max(1, int(round(0.999 * rel.numel())))
    return float(rel.kthvalue(k).values)
