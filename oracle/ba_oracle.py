"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's dense bundle adjustment.

This file restates, stage by stage, what `slam_ext.ba` does in the reference
(/root/reference/csrc/slam_ext/geom_kernels.cu:1283-1404, `ba_cuda`) with plain torch CPU ops, in a
chosen dtype (float64 = the parity oracle; float32 = the "port" CPU baseline that bench.py times).
It is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  The shipped path (vipe_b200/) never does.

Pinning status: the reference ships no tests, fixtures or golden vectors, so this
oracle is pinned against the reference itself: the UNMODIFIED reference CUDA sources are compiled by
oracle/build_ref.py (Eigen replaced by oracle/eigen_stub, nothing else) and run on a B200;
tests/test_reference_pin.py compares every output of that run with this file, and
tests/golden/ holds vectors produced by that run (tests/golden/make_golden.py).  Until such a run has
been recorded the status is "parity unpinned"; see DESIGN.md §Oracle for the current state.

Every function cites the reference lines it follows.  Reference quirks Q1-Q9 are
reproduced on purpose.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import torch

MIN_DEPTH = 0.25  # geom_kernels.cu:33
WEIGHT_SCALE = 0.001  # geom_kernels.cu:304-305
ALPHA = 0.05  # geom_kernels.cu:1359
STEREO_BASELINE = -0.1  # geom_kernels.cu:222


# ----------------------------------------------------------------------------- SE3 helpers
def act_so3(q: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """Rotate X[...,3] by quaternion q[...,4] = (x,y,z,w).  geom_kernels.cu:69-78."""
    qv, qw = q[..., :3], q[..., 3:4]
    uv = 2.0 * torch.linalg.cross(qv.expand_as(X), X, dim=-1)
    return X + qw * uv + torch.linalg.cross(qv.expand_as(X), uv, dim=-1)


def rel_se3(ti, qi, tj, qj):
    """Tij = Tj * Ti^-1.  geom_kernels.cu:104-114."""
    qij = torch.stack(
        [
            -qj[..., 3] * qi[..., 0] + qj[..., 0] * qi[..., 3] - qj[..., 1] * qi[..., 2] + qj[..., 2] * qi[..., 1],
            -qj[..., 3] * qi[..., 1] + qj[..., 1] * qi[..., 3] - qj[..., 2] * qi[..., 0] + qj[..., 0] * qi[..., 2],
            -qj[..., 3] * qi[..., 2] + qj[..., 2] * qi[..., 3] - qj[..., 0] * qi[..., 1] + qj[..., 1] * qi[..., 0],
            qj[..., 3] * qi[..., 3] + qj[..., 0] * qi[..., 0] + qj[..., 1] * qi[..., 1] + qj[..., 2] * qi[..., 2],
        ],
        dim=-1,
    )
    tij = tj - act_so3(qij, ti)
    return tij, qij


def adj_se3(t, q, X):
    """Y = Adj(T)^T-style map used for dl/dTi.  geom_kernels.cu:88-102.  X, Y: [...,6]."""
    qinv = torch.cat([-q[..., :3], q[..., 3:4]], dim=-1)
    a, b = X[..., :3], X[..., 3:]
    # u = (t2*X1 - t1*X2, t0*X2 - t2*X0, t1*X0 - t0*X1) = a x t
    u = torch.linalg.cross(a, t.expand_as(a), dim=-1)
    return torch.cat([act_so3(qinv, a), act_so3(qinv, b) + act_so3(qinv, u)], dim=-1)


def exp_so3(phi):
    """geom_kernels.cu:116-136 (small-angle branch at theta^2 < 1e-8)."""
    th2 = (phi * phi).sum(-1, keepdim=True)
    th4 = th2 * th2
    th = th2.sqrt()
    small = th2 < 1e-8
    ths = torch.where(small, torch.ones_like(th), th)
    imag = torch.where(small, 0.5 - th2 / 48.0 + th4 / 3840.0, torch.sin(0.5 * ths) / ths)
    real = torch.where(small, 1.0 - th2 / 8.0 + th4 / 384.0, torch.cos(0.5 * ths))
    return torch.cat([imag * phi, real], dim=-1)


def exp_se3(xi):
    """geom_kernels.cu:150-177 (translation Jacobian applied only when theta > 1e-4)."""
    tau, phi = xi[..., :3], xi[..., 3:]
    q = exp_so3(phi)
    th2 = (phi * phi).sum(-1, keepdim=True)
    th = th2.sqrt()
    big = th > 1e-4
    ths = torch.where(big, th, torch.ones_like(th))
    th2s = torch.where(big, th2, torch.ones_like(th2))
    a = (1 - torch.cos(ths)) / th2s
    b = (ths - torch.sin(ths)) / (ths * th2s)
    c1 = torch.linalg.cross(phi, tau, dim=-1)
    c2 = torch.linalg.cross(phi, c1, dim=-1)
    t = tau + torch.where(big, a * c1 + b * c2, torch.zeros_like(tau))
    return t, q


def retr_se3(xi, t, q):
    """T <- exp(xi) * T without quaternion renormalisation (Q6).  geom_kernels.cu:882-899."""
    dt, dq = exp_se3(xi)
    q1 = torch.stack(
        [
            dq[..., 3] * q[..., 0] + dq[..., 0] * q[..., 3] + dq[..., 1] * q[..., 2] - dq[..., 2] * q[..., 1],
            dq[..., 3] * q[..., 1] + dq[..., 1] * q[..., 3] + dq[..., 2] * q[..., 0] - dq[..., 0] * q[..., 2],
            dq[..., 3] * q[..., 2] + dq[..., 2] * q[..., 3] + dq[..., 0] * q[..., 1] - dq[..., 1] * q[..., 0],
            dq[..., 3] * q[..., 3] - dq[..., 0] * q[..., 0] - dq[..., 1] * q[..., 1] - dq[..., 2] * q[..., 2],
        ],
        dim=-1,
    )
    t1 = act_so3(dq, t) + dt
    return t1, q1


# ----------------------------------------------------------------------------- stage 1: linearisation
def relative_poses(poses, ii, jj):
    """Per-edge (tij, qij) including the stereo convention Q2.  geom_kernels.cu:219-249."""
    ti, qi = poses[ii, :3], poses[ii, 3:]
    tj, qj = poses[jj, :3], poses[jj, 3:]
    tij, qij = rel_se3(ti, qi, tj, qj)
    stereo = ii == jj
    if stereo.any():
        tij = tij.clone()
        qij = qij.clone()
        tij[stereo] = torch.tensor([STEREO_BASELINE, 0.0, 0.0], dtype=poses.dtype)
        qij[stereo] = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=poses.dtype)
    return tij, qij, stereo


def linearize(poses, disps, intrinsics, targets, weights, ii, jj):
    """projective_transform_kernel, geom_kernels.cu:178-432.

    Returns Hs[4,E,6,6] (ii,ij,ji,jj), vs[2,E,6], Eii[E,6,HW], Eij[E,6,HW], Cii[E,HW], bz[E,HW].
    """
    dt = poses.dtype
    E = ii.numel()
    ht, wd = disps.shape[1], disps.shape[2]
    HW = ht * wd
    fx, fy, cx, cy = [intrinsics[k] for k in range(4)]
    tij, qij, stereo = relative_poses(poses, ii, jj)

    v, u = torch.meshgrid(torch.arange(ht, dtype=dt), torch.arange(wd, dtype=dt), indexing="ij")
    u = u.reshape(1, HW)
    v = v.reshape(1, HW)
    Xi = torch.stack(
        [((u - cx) / fx).expand(E, HW), ((v - cy) / fy).expand(E, HW), torch.ones(E, HW, dtype=dt)], dim=-1
    )  # [E,HW,3]   :289-291
    h = disps.reshape(-1, HW)[ii]  # [E,HW]    :292
    Xj = act_so3(qij[:, None, :], Xi) + h[..., None] * tij[:, None, :]  # actSE3 :80-86,295
    x, y, z = Xj[..., 0], Xj[..., 1], Xj[..., 2]
    valid = ~(z < MIN_DEPTH)  # :301
    d = torch.where(valid, 1.0 / torch.where(valid, z, torch.ones_like(z)), torch.zeros_like(z))
    d2 = d * d
    tg = targets.reshape(E, 2, HW)
    wt = weights.reshape(E, 2, HW)
    zero = torch.zeros((), dtype=dt)
    wu = torch.where(valid, WEIGHT_SCALE * wt[:, 0], zero)  # :304
    wv = torch.where(valid, WEIGHT_SCALE * wt[:, 1], zero)  # :305
    ru = tg[:, 0] - (fx * d * x + cx)  # :308
    rv = tg[:, 1] - (fy * d * y + cy)  # :309

    o = torch.zeros_like(x)
    Jju = fx * torch.stack([h * d, o, -x * h * d2, -x * y * d2, 1 + x * x * d2, -y * d], dim=-1)  # :314-319
    Jjv = fy * torch.stack([o, h * d, -y * h * d2, -1 - y * y * d2, x * y * d2, x * d], dim=-1)  # :356-361
    Jzu = fx * (tij[:, None, 0] * d - tij[:, None, 2] * (x * d2))  # :322
    Jzv = fy * (tij[:, None, 1] * d - tij[:, None, 2] * (y * d2))  # :363

    Cii = wu * Jzu * Jzu + wv * Jzv * Jzv  # :325,364   (before the stereo zeroing)
    bz = wu * ru * Jzu + wv * rv * Jzv  # :326,365

    ns = (~stereo).to(dt)[:, None]  # :329,367
    wu = wu * ns
    wv = wv * ns

    Jiu = -adj_se3(tij[:, None, :], qij[:, None, :], Jju)  # :332-333
    Jiv = -adj_se3(tij[:, None, :], qij[:, None, :], Jjv)  # :369-370

    Jxu = torch.cat([Jiu, Jju], dim=-1)  # [E,HW,12]
    Jxv = torch.cat([Jiv, Jjv], dim=-1)
    H12 = torch.einsum("ep,epn,epm->enm", wu, Jxu, Jxu) + torch.einsum("ep,epn,epm->enm", wv, Jxv, Jxv)  # :336-342
    Hs = torch.stack([H12[:, :6, :6], H12[:, :6, 6:], H12[:, 6:, :6], H12[:, 6:, 6:]], dim=0)  # :417-426
    v12 = torch.einsum("ep,epn->en", wu * ru, Jxu) + torch.einsum("ep,epn->en", wv * rv, Jxv)  # :345-347
    vs = torch.stack([v12[:, :6], v12[:, 6:]], dim=0)
    Eii = ((wu * Jzu)[..., None] * Jiu + (wv * Jzv)[..., None] * Jiv).permute(0, 2, 1).contiguous()  # :350,384
    Eij = ((wu * Jzu)[..., None] * Jju + (wv * Jzv)[..., None] * Jjv).permute(0, 2, 1).contiguous()  # :351,385
    return Hs, vs, Eii, Eij, Cii, bz


# ----------------------------------------------------------------------------- index bookkeeping
@dataclass
class Bookkeeping:
    """Index structures of ba_cuda, geom_kernels.cu:1301-1308."""

    ts: torch.Tensor
    ii_exp: torch.Tensor
    jj_exp: torch.Tensor
    kx: torch.Tensor
    kk_exp: torch.Tensor


def bookkeeping(ii, jj, t0, t1) -> Bookkeeping:
    ts = torch.arange(t0, t1, dtype=torch.int64)
    ii_exp = torch.cat([ts, ii])
    jj_exp = torch.cat([ts, jj])
    kx, kk_exp = torch.unique(ii_exp, sorted=True, return_inverse=True)
    return Bookkeeping(ts, ii_exp, jj_exp, kx, kk_exp)


def accum(data, ix, jx):
    """accum_cuda/accum_kernel, geom_kernels.cu:863-880,946-992: out[j] = sum_{e: ix[e]==jx[j]} data[e].

    jx is sorted and unique at every call site (kx or ts)."""
    pos = torch.searchsorted(jx, ix)
    pos_c = pos.clamp(max=jx.numel() - 1)
    hit = jx[pos_c] == ix
    out = torch.zeros((jx.numel(),) + tuple(data.shape[1:]), dtype=data.dtype)
    out.index_add_(0, pos_c[hit], data[hit])
    return out


def csr_by_source(ix, jx):
    """The (ptrs, idxs) CSR accum_cuda builds on the host, geom_kernels.cu:946-981 (rows as sets)."""
    rows = []
    for j in jx.tolist():
        rows.append(sorted(torch.nonzero(ix == j).flatten().tolist()))
    return rows


def schur_triples(bk: Bookkeeping, t0, t1):
    """The (row_a, row_b, frame) triple list of schur_block, geom_kernels.cu:1209-1240, and the
    (pose_a, pose_b) block each one lands in.  Pure-Python: small cases only.  Rows whose target pose is
    outside [t0, t1) are dropped (the reference's `j <= t1` at :1216 would index out of bounds at j == t1;
    callers keep max(ii,jj) < t1, Q5)."""
    P = t1 - t0
    graph = [[] for _ in range(P)]
    index = [[] for _ in range(P)]
    for n, (j, k) in enumerate(zip(bk.jj_exp.tolist(), bk.kk_exp.tolist())):
        if t0 <= j < t1:
            graph[j - t0].append(k)
            index[j - t0].append(n)
    trip, blocks = [], []
    for i in range(P):
        for j in range(P):
            for a, ka in enumerate(graph[i]):
                for b, kb in enumerate(graph[j]):
                    if ka == kb:
                        trip.append((index[i][a], index[j][b], ka))
                        blocks.append((i, j))
    return trip, blocks


# ----------------------------------------------------------------------------- pose system + solve
def assemble_pose_system(Hs, vs, ii, jj, t0, t1):
    """SparseBlock::update_lhs/update_rhs as called at geom_kernels.cu:1343-1347 (fp64 assembly, Q7).

    Blocks whose row or column pose is < t0 are dropped (:1127,:1148).  Indices >= t1 are undefined
    behaviour in the reference (no upper-bound check); they are dropped here too."""
    P = t1 - t0
    A = torch.zeros(P, 6, P, 6, dtype=torch.float64)
    b = torch.zeros(P, 6, dtype=torch.float64)
    rows = torch.cat([ii, ii, jj, jj]) - t0
    cols = torch.cat([ii, jj, ii, jj]) - t0
    blocks = Hs.reshape(-1, 6, 6).to(torch.float64)
    ok = (rows >= 0) & (cols >= 0) & (rows < P) & (cols < P)
    A.index_put_((rows[ok][:, None, None], torch.arange(6)[None, :, None], cols[ok][:, None, None],
                  torch.arange(6)[None, None, :]), blocks[ok], accumulate=True)
    r = torch.cat([ii, jj]) - t0
    okr = (r >= 0) & (r < P)
    b.index_add_(0, r[okr], vs.reshape(-1, 6).to(torch.float64)[okr])
    return A.reshape(6 * P, 6 * P), b.reshape(6 * P)


def solve_damped(A, b, lm, ep):
    """SparseBlock::solve, geom_kernels.cu:1172-1191: diag += ep + lm*diag, fp64 LLT, failure => zeros."""
    L = A.clone()
    d = L.diagonal()
    d += float(ep) + float(lm) * d.clone()
    chol, info = torch.linalg.cholesky_ex(L)
    if int(info) != 0 or not torch.isfinite(chol).all():
        return torch.zeros_like(b), False
    x = torch.cholesky_solve(b[:, None], chol)[:, 0]
    return x, True


# ----------------------------------------------------------------------------- the driver
@dataclass
class Trace:
    """Intermediates of the LAST iteration (for per-stage kernel tests)."""

    bk: Bookkeeping | None = None
    Hs: torch.Tensor | None = None
    vs: torch.Tensor | None = None
    Eii: torch.Tensor | None = None
    Eij: torch.Tensor | None = None
    Cii: torch.Tensor | None = None
    bz: torch.Tensor | None = None
    A: torch.Tensor | None = None
    b: torch.Tensor | None = None
    C: torch.Tensor | None = None
    w: torch.Tensor | None = None
    Q: torch.Tensor | None = None
    S: torch.Tensor | None = None
    sv: torch.Tensor | None = None
    dx: torch.Tensor | None = None
    dw: torch.Tensor | None = None
    dz: torch.Tensor | None = None
    chol_ok: list = field(default_factory=list)
    energy: list = field(default_factory=list)


def ba(poses, disps, intrinsics, disps_sens, targets, weights, eta, ii, jj, t0, t1, iterations, lm, ep,
       motion_only, dtype=torch.float64, trace: Trace | None = None):
    """ba_cuda, geom_kernels.cu:1283-1404.  Same argument order and meaning as slam_ext.ba
    (csrc/slam_ext/slam.cpp:24-27).  `poses` and `disps` are updated IN PLACE when they already have
    `dtype`; otherwise converted copies are updated and written back at the end.  Returns [dx, dz]."""
    P = t1 - t0
    ht, wd = disps.shape[1], disps.shape[2]
    HW = ht * wd
    ii = ii.to(torch.int64).cpu()
    jj = jj.to(torch.int64).cpu()
    p = poses.detach().cpu().to(dtype).clone()
    dsp = disps.detach().cpu().to(dtype).clone()
    intr = intrinsics.detach().cpu().to(dtype)
    dsens = disps_sens.detach().cpu().to(dtype)
    tg = targets.detach().cpu().to(dtype)
    wt = weights.detach().cpu().to(dtype)
    E = ii.numel()

    bk = bookkeeping(ii, jj, t0, t1)  # :1301-1308
    K = bk.kx.numel()
    et = eta.detach().cpu().to(dtype).reshape(-1, HW) if not motion_only else None
    dx = torch.zeros(P, 6, dtype=dtype)
    dz = torch.zeros(K, HW, dtype=dtype)

    for _ in range(iterations):
        Hs, vs, Eii, Eij, Cii, bz = linearize(p, dsp, intr, tg, wt, ii, jj)  # :1325-1340
        A, b = assemble_pose_system(Hs, vs, ii, jj, t0, t1)  # :1343-1347
        if trace is not None:
            trace.bk, trace.Hs, trace.vs, trace.Eii, trace.Eij, trace.Cii, trace.bz = bk, Hs, vs, Eii, Eij, Cii, bz
            trace.A, trace.b = A, b
        if motion_only:
            x, ok = solve_damped(A, b, lm, ep)  # :1350
            dx = x.reshape(P, 6).to(dtype)
        else:
            m = (dsens[bk.kx] > 0).to(dtype).reshape(-1, HW)  # :1361-1363
            C = accum(Cii, ii, bk.kx) + m * ALPHA + (1 - m) * et  # :1365
            w = accum(bz, ii, bk.kx) - m * ALPHA * (dsp[bk.kx] - dsens[bk.kx]).reshape(-1, HW)  # :1367-1369
            Q = 1.0 / C  # :1370
            Ei = accum(Eii.reshape(E, 6 * HW), ii, bk.ts).reshape(P, 6, HW)  # :1373
            Eall = torch.cat([Ei, Eij], dim=0)  # :1374

            # schur_block, :1198-1281.  S(a,b) = sum over frames k of F[a,k] diag(Q[k]) F[b,k]^T where the
            # rows of Eall that belong to pose a and frame k are exactly the triples of :1225-1240.
            pose_of_row = bk.jj_exp - t0
            row_ok = (pose_of_row >= 0) & (pose_of_row < P)
            S = torch.zeros(P, 6, P, 6, dtype=torch.float64)
            for k in range(K):
                rows = torch.nonzero((bk.kk_exp == k) & row_ok).flatten()
                if rows.numel() == 0:
                    continue
                Ek = Eall[rows]  # [r,6,HW]
                Sk = torch.einsum("aip,bjp->aibj", Ek * Q[k], Ek).to(torch.float64)  # EEt6x6 :994-1046
                pa = pose_of_row[rows]
                S.index_put_((pa[:, None, None, None], torch.arange(6)[None, :, None, None],
                              pa[None, None, :, None], torch.arange(6)[None, None, None, :]), Sk, accumulate=True)
            qw = (Q * w)[bk.kk_exp]  # [P+E,HW]
            vrow = torch.einsum("np,nip->ni", qw, Eall).to(torch.float64)  # Ev6x1 :1048-1080
            sv = torch.zeros(P, 6, dtype=torch.float64)
            sv.index_add_(0, pose_of_row[row_ok], vrow[row_ok])  # :1278
            S2 = S.reshape(6 * P, 6 * P)
            x, ok = solve_damped(A - S2, b - sv.reshape(-1), lm, ep)  # :1378
            dx = x.reshape(P, 6).to(dtype)

            # EvT6x1 :1082-1098 with the `idx <= 0` skip (Q4), then dz :1390
            back_ok = (pose_of_row > 0) & (pose_of_row < P)
            dxe = torch.zeros(P + E, 6, dtype=dtype)
            dxe[back_ok] = dx[pose_of_row[back_ok]]
            dw = torch.einsum("nip,ni->np", Eall, dxe)
            dz = Q * (w - accum(dw, bk.ii_exp, bk.kx))
            if trace is not None:
                trace.C, trace.w, trace.Q, trace.S, trace.sv, trace.dw = C, w, Q, S2, sv, dw
        if trace is not None:
            trace.dx, trace.dz = dx, dz
            trace.chol_ok.append(ok)

        # retractions :1353 / :1393-1399
        t_new, q_new = retr_se3(dx, p[t0:t1, :3], p[t0:t1, 3:])
        p[t0:t1, :3] = t_new
        p[t0:t1, 3:] = q_new
        if not motion_only:
            dsp[bk.kx] = dsp[bk.kx] + dz.reshape(K, ht, wd)

    with torch.no_grad():
        poses.copy_(p.to(poses.dtype))
        disps.copy_(dsp.to(disps.dtype))
    return [dx, dz]


def energy(poses, disps, intrinsics, targets, weights, ii, jj, dtype=torch.float64):
    """Weighted squared reprojection error sum_e sum_px w*r^2 with the reference's validity rule
    (not a reference function; used by property tests: BA must not increase it near the optimum)."""
    p = poses.detach().cpu().to(dtype)
    dsp = disps.detach().cpu().to(dtype)
    intr = intrinsics.detach().cpu().to(dtype)
    E = ii.numel()
    ht, wd = dsp.shape[1:]
    HW = ht * wd
    fx, fy, cx, cy = [intr[k] for k in range(4)]
    tij, qij, _ = relative_poses(p, ii, jj)
    v, u = torch.meshgrid(torch.arange(ht, dtype=dtype), torch.arange(wd, dtype=dtype), indexing="ij")
    Xi = torch.stack([((u.reshape(1, HW) - cx) / fx).expand(E, HW), ((v.reshape(1, HW) - cy) / fy).expand(E, HW),
                      torch.ones(E, HW, dtype=dtype)], dim=-1)
    h = dsp.reshape(-1, HW)[ii]
    Xj = act_so3(qij[:, None, :], Xi) + h[..., None] * tij[:, None, :]
    valid = ~(Xj[..., 2] < MIN_DEPTH)
    d = torch.where(valid, 1.0 / torch.where(valid, Xj[..., 2], torch.ones_like(h)), torch.zeros_like(h))
    tg = targets.detach().cpu().to(dtype).reshape(E, 2, HW)
    wt = weights.detach().cpu().to(dtype).reshape(E, 2, HW) * valid[:, None, :]
    ru = tg[:, 0] - (fx * d * Xj[..., 0] + cx)
    rv = tg[:, 1] - (fy * d * Xj[..., 1] + cy)
    return float((WEIGHT_SCALE * (wt[:, 0] * ru * ru + wt[:, 1] * rv * rv)).sum())
