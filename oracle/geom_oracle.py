"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch, float64 by default) of the reference's four other `slam_ext`
operators: projmap, frame_distance, depth_filter, iproj (/root/reference/csrc/slam_ext/geom_kernels.cu:434-861).
Pinned against the reference's own CUDA run in tests/test_geom_ops.py (oracle/_ref) -- see oracle/ba_oracle.py for
the rules that apply to everything under oracle/."""

from __future__ import annotations

import torch

from .ba_oracle import MIN_DEPTH, act_so3, rel_se3


def _grid(ht, wd, intr, dt):
    v, u = torch.meshgrid(torch.arange(ht, dtype=dt), torch.arange(wd, dtype=dt), indexing="ij")
    fx, fy, cx, cy = [intr[k] for k in range(4)]
    return u.reshape(-1), v.reshape(-1), (u.reshape(-1) - cx) / fx, (v.reshape(-1) - cy) / fy


def _transform(poses, ii, jj, xn, yn, h):
    """Xj = Tij * (xn, yn, 1, h) with Tij = Tj Ti^-1 and NO stereo convention (relSE3 as is, :484)."""
    tij, qij = rel_se3(poses[ii, :3], poses[ii, 3:], poses[jj, :3], poses[jj, 3:])
    Xi = torch.stack([xn.expand_as(h), yn.expand_as(h), torch.ones_like(h)], dim=-1)
    return act_so3(qij[:, None, :], Xi) + h[..., None] * tij[:, None, :], tij


def projmap(poses, disps, intrinsics, ii, jj, dtype=torch.float64):
    """projmap_kernel, geom_kernels.cu:434-519."""
    p, d, k = poses.to(dtype), disps.to(dtype), intrinsics.to(dtype)
    N, ht, wd = d.shape
    u, v, xn, yn = _grid(ht, wd, k, dtype)
    Xj, _ = _transform(p, ii, jj, xn[None], yn[None], d.reshape(N, -1)[ii])
    z = Xj[..., 2]
    ok = z > 0.01
    zs = torch.where(ok, z, torch.ones_like(z))
    cu = torch.where(ok, k[0] * (Xj[..., 0] / zs) + k[2], u[None].expand_as(z))
    cv = torch.where(ok, k[1] * (Xj[..., 1] / zs) + k[3], v[None].expand_as(z))
    coords = torch.stack([cu, cv, torch.zeros_like(cu)], dim=-1).reshape(-1, ht, wd, 3)
    valid = (z > MIN_DEPTH).to(dtype).reshape(-1, ht, wd, 1)
    return coords, valid, z.reshape(-1, ht, wd)


def frame_distance(poses, disps, intrinsics, pi, pj, qi, qj, di, beta, dtype=torch.float64):
    """frame_distance_kernel, geom_kernels.cu:521-676."""
    p, d, K = poses.to(dtype), disps.to(dtype), intrinsics.to(dtype)
    N, ht, wd = d.shape
    HW = ht * wd
    v, u = torch.meshgrid(torch.arange(ht, dtype=dtype), torch.arange(wd, dtype=dtype), indexing="ij")
    u, v = u.reshape(1, HW), v.reshape(1, HW)
    ki, kj = K[qi], K[qj]
    xn = (u - ki[:, 2:3]) / ki[:, 0:1]
    yn = (v - ki[:, 3:4]) / ki[:, 1:2]
    h = d.reshape(N, HW)[di]
    Xj, tij = _transform(p, pi, pj, xn, yn, h)

    def flow(X):
        du = kj[:, 0:1] * (X[..., 0] / X[..., 2]) + kj[:, 2:3] - u
        dv = kj[:, 1:2] * (X[..., 1] / X[..., 2]) + kj[:, 3:4] - v
        return torch.sqrt(du * du + dv * dv), X[..., 2] > MIN_DEPTH

    d1, ok1 = flow(Xj)
    Xt = torch.stack([xn + h * tij[:, 0:1], yn + h * tij[:, 1:2], 1 + h * tij[:, 2:3]], dim=-1)
    d2, ok2 = flow(Xt)
    accum = (beta * torch.where(ok1, d1, torch.zeros_like(d1)) + (1 - beta) * torch.where(ok2, d2, torch.zeros_like(d2))).sum(-1)
    valid = (beta * ok1.to(dtype) + (1 - beta) * ok2.to(dtype)).sum(-1)
    total = torch.full_like(valid, float(HW))
    ratio = valid / (total + 1e-8)
    return torch.where(ratio < 0.75, torch.full_like(valid, 1000.0), accum / valid.clamp_min(1e-30)), ratio


def depth_filter(poses, disps, intrinsics, ix, thresh, dtype=torch.float64):
    """depth_filter_kernel, geom_kernels.cu:678-793.  Returns (counter, margin) where margin is the smallest distance
    of any comparison to its threshold (pixels with a tiny margin may legitimately differ in fp32)."""
    p, d, k = poses.to(dtype), disps.to(dtype), intrinsics.to(dtype)
    N, ht, wd = d.shape
    HW = ht * wd
    u, v, xn, yn = _grid(ht, wd, k, dtype)
    counter = torch.zeros(ix.numel(), HW, dtype=dtype)
    margin = torch.full((ix.numel(), HW), float("inf"), dtype=dtype)
    for b, i in enumerate(ix.tolist()):
        for n in range(6):
            j = i - n - 1 if n < 3 else i + n - 2  # :709
            if j < 0 or j >= N:
                continue
            di = d[i].reshape(1, HW)
            Xj, _ = _transform(p, torch.tensor([i]), torch.tensor([j]), xn[None], yn[None], di)
            X = Xj[0]
            uj = k[0] * (X[:, 0] / X[:, 2]) + k[2]
            vj = k[1] * (X[:, 1] / X[:, 2]) + k[3]
            dj = di[0] / X[:, 2]
            u0, v0 = torch.floor(uj), torch.floor(vj)
            inside = (u0 >= 0) & (v0 >= 0) & (u0 < wd - 1) & (v0 < ht - 1) & torch.isfinite(uj) & torch.isfinite(vj)
            u0c = u0.clamp(0, wd - 2).long()
            v0c = v0.clamp(0, ht - 2).long()
            hit = torch.zeros(HW, dtype=torch.bool)
            for dv_, du_ in ((0, 0), (0, 1), (1, 0), (1, 1)):
                dn = d[j][v0c + dv_, u0c + du_]
                err = (1.0 / dj - 1.0 / dn).abs()
                hit |= err < thresh[b].to(dtype)
                margin[b] = torch.minimum(margin[b], torch.where(inside, (err - thresh[b].to(dtype)).abs(), margin[b]))
            # pixels that land within 1e-3 px of a cell boundary may pick another cell in fp32
            edge = torch.minimum((uj - torch.round(uj)).abs(), (vj - torch.round(vj)).abs())
            margin[b] = torch.minimum(margin[b], torch.where(edge < 1e-3, torch.zeros_like(edge), margin[b]))
            counter[b] += (hit & inside).to(dtype)
    return counter.reshape(-1, ht, wd), margin.reshape(-1, ht, wd)


def iproj(poses, disps, intrinsics, dtype=torch.float64):
    """iproj_kernel, geom_kernels.cu:795-861."""
    p, d, k = poses.to(dtype), disps.to(dtype), intrinsics.to(dtype)
    N, ht, wd = d.shape
    u, v, xn, yn = _grid(ht, wd, k, dtype)
    h = d.reshape(N, -1)
    Xi = torch.stack([xn[None].expand_as(h), yn[None].expand_as(h), torch.ones_like(h)], dim=-1)
    X = act_so3(p[:, None, 3:], Xi) + h[..., None] * p[:, None, :3]
    return (X / h[..., None]).reshape(N, ht, wd, 3)
