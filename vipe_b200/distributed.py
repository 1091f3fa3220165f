"""Keyframe-sharded bundle adjustment over one node.

One process per GPU.  Disparities couple only to the edges that leave their own frame
(geom_kernels.cu:292 reads disps[ii]; C, w, E are grouped by `ii`, :1365-1373), so every rank eliminates the
disparities of the source frames it owns locally and only the reduced camera system
[H ; b] (fp64, 6P x 6P + 6P) is summed across ranks -- one `all_reduce` per Gauss-Newton iteration over
NCCL / NVLink -- or, where NVSwitch multicast memory is available, no collective at all: the Cholesky kernel reads the
sum over the ranks' partial systems through the switch (`collective="nvls"`).  The damped Cholesky solve is replicated (every rank reduces the same buffer, so every rank
computes the same dx), back-substitution and disparity retraction are local to the owner, and the owned
disparity rows are exchanged once per call, not per iteration.

`ba_sharded` has the signature of `slam_ext.ba` plus a process group.  The engine (what runs between the
collectives) is pluggable so the orchestration is testable with `gloo` on CPU; the product engine is the CUDA one.
"""

from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from .ext import slam_ext


# from this many unknowns on, collective="auto" distributes the factorisation (C4: 5994; C3's 1794 are latency-bound and
# stay replicated)
DIST_SOLVE_MIN_UNKNOWNS = 4096


class PeerSystem:
    """Two instances of the reduced camera system [H ; b ; diag(A)] in NVSwitch multicast (symmetric) memory.

    With them the per-iteration all-reduce disappears: every rank accumulates its partial system into its own
    instance, one cross-rank barrier follows, and the Cholesky kernel of every rank loads its input with
    `multimem.ld_reduce` -- the sum over the ranks' instances, formed inside the switch, tile by tile as the
    factorisation gets there (include/vipe_ba.h: vipe_ba_set_peer_system).  torch supplies the plumbing only:
    the symmetric allocation, the rendezvous that maps the multicast address, and the barrier.

    Construction is collective and every step is agreed on by all ranks before the next one (a rank that cannot
    allocate must not leave the others waiting in the rendezvous); `ok` is the same on every rank."""

    def __init__(self, plan, dev, group):
        g = group if group is not None else dist.group.WORLD
        self.ok, self.k = False, 0  # k: iterations so far, over all calls -- the buffer parity never restarts (vipe_ba.h)
        self.bufs, self.hdls = [], []

        def all_agree(flag: bool) -> bool:
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=g)
            return bool(t.item())

        symm = None
        try:
            import torch.distributed._symmetric_memory as symm

            _, npad = plan.system_view(plan.workspace(dev))
            count = npad * npad + 2 * npad
            self.bufs = [symm.empty(count, dtype=torch.float64, device=dev) for _ in range(3)]  # two accumulators + the reduced system / the factor
            # auxiliary buffer of the distributed factorisation (ready flags, 1/diag, inverted diagonal tiles): zeroed once, here
            aux_bytes = int(_lib.lib().vipe_ba_dist_aux_bytes(plan.handle))
            self.bufs.append(symm.empty((aux_bytes + 7) // 8, dtype=torch.float64, device=dev))
            for b in self.bufs:
                b.zero_()
            good = True
        except Exception:
            good = False
        if not all_agree(good):
            return
        try:
            self.hdls = [symm.rendezvous(b, g) for b in self.bufs]
            good = all(h.multicast_ptr for h in self.hdls)
        except Exception:
            good = False
        self.ok = all_agree(good)

    @staticmethod
    def get(plan, dev, group):
        key = ("peer", dev.index, None if group is None else id(group))
        cache = plan.__dict__.setdefault("_peer_systems", {})
        if key not in cache:
            cache[key] = PeerSystem(plan, dev, group)
        return cache[key]


def nvls_available(plan, dev, group) -> bool:
    """The fused reduction needs multicast memory and the tiled solver (more than 128 unknowns).  Collective: every
    rank gets the same answer."""
    return 6 * plan.P > 128 and PeerSystem.get(plan, dev, group).ok


class CudaShardEngine:
    """Calls the C ABI's two phases on this rank's shard (include/vipe_ba.h: vipe_ba_linearize / vipe_ba_solve_update)."""

    def __init__(self, plan, poses, disps, intrinsics, disps_sens, targets, weights, eta, motion_only):
        dev = poses.device
        self.plan, self.dev, self.motion_only = plan, dev, bool(motion_only)
        HW = plan.ht * plan.wd
        self.dx = torch.zeros(plan.P, 6, dtype=torch.float32, device=dev)
        self.dz = torch.zeros(plan.K, HW, dtype=torch.float32, device=dev)
        self.ws = plan.workspace(dev)
        self.tens = slam_ext._tensors(poses, disps, intrinsics, disps_sens, targets, weights, eta, self.dx, self.dz,
                                      self.motion_only)
        self.system, self.npad = plan.system_view(self.ws)

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def linearize(self) -> torch.Tensor:
        _lib.check(_lib.lib().vipe_ba_linearize(self.plan.handle, C.byref(self.tens), self.ws.data_ptr(),
                                                int(self.motion_only), self._stream()), "vipe_ba_linearize")
        return self.system

    def use_peer_buffer(self, local_ptr, multicast_ptr):
        _lib.check(_lib.lib().vipe_ba_set_peer_system(self.plan.handle, local_ptr, multicast_ptr), "vipe_ba_set_peer_system")

    def use_solve_buffer(self, local_ptr):
        _lib.check(_lib.lib().vipe_ba_set_solve_buffer(self.plan.handle, local_ptr), "vipe_ba_set_solve_buffer")

    def use_dist_solve(self, factor_local, factor_mc, aux_local, aux_mc, rank, world):
        _lib.check(_lib.lib().vipe_ba_set_dist_solve(self.plan.handle, factor_local, factor_mc, aux_local, aux_mc, int(rank), int(world)),
                   "vipe_ba_set_dist_solve")

    def use_owned_rows(self, on: bool):
        _lib.check(_lib.lib().vipe_ba_set_owned_rows(self.plan.handle, int(on)), "vipe_ba_set_owned_rows")

    def peer_reduce(self, accum_mc, reduced_mc, rank, world):
        _lib.check(_lib.lib().vipe_ba_peer_reduce(self.plan.handle, accum_mc, reduced_mc, int(rank), int(world), self._stream()),
                   "vipe_ba_peer_reduce")

    def solve_update(self, lm: float, ep: float):
        _lib.check(_lib.lib().vipe_ba_solve_update(self.plan.handle, C.byref(self.tens), self.ws.data_ptr(), float(lm),
                                                   float(ep), int(self.motion_only), self._stream()),
                   "vipe_ba_solve_update")


def ba_sharded(poses, disps, intrinsics, disps_sens, targets, weights, eta, ii, jj, t0, t1, iterations, lm, ep,
               motion_only, group=None, engine_cls=CudaShardEngine, exchange=True, profile=None, collective="auto", plan=None,
               owned_inputs=False):
    """`slam_ext.ba` sharded by source keyframe across the ranks of `group`.

    `collective`: "allreduce" = one NCCL all-reduce of the reduced camera system per iteration; "nvls" = no collective
    launch at all, the sum is formed by the NVSwitch inside the Cholesky kernel's loads (PeerSystem above); "dist" = the same
    fused input, and the factorisation itself is DISTRIBUTED: tile column j belongs to rank j % world, results are multicast
    to every rank through the switch (include/vipe_ba.h: vipe_ba_set_dist_solve) -- for large systems, where the replicated
    solve is what keeps more GPUs from helping; "nvls2" = the
    same in-switch sum as a small kernel of its own (every rank reduces 1/world of the tiles with multimem.ld_reduce and
    multicasts them to all ranks with multimem.st) followed by a solve on local memory; "auto" = "nvls" on 2 GPUs,
    "nvls2" from 4 GPUs on (the fused loads cost the solve 0.14 ms at 8 GPUs), "allreduce" where multicast memory is missing.

    `owned_inputs`: `targets` / `weights` hold only the rows of this rank's edges, [n_owned, 2, ht, wd] in the order of
    `plan.owned_edges()` (a rank holds the inputs of its edges); otherwise every rank passes the
    full tensors and reads only its rows.  Poses, disparities, eta are replicated either way.  On return `poses` is identical
    on all ranks and, if `exchange`, so are `disps[kx]` and `dz`.

    `profile`: optional dict; CUDA event pairs around (linearise | all-reduce | solve+update) of every iteration are
    appended to profile["events"] and the plan is stored in profile["plan"] (bench.py reads them after a sync)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    t0, t1 = int(t0), int(t1)
    N, ht, wd = disps.shape
    # `plan`: this rank's BAPlan if the caller already has it (e.g. built from a host copy of the edge list); else the
    # identity cache first: the same ii/jj tensor objects as last time => no device-to-host copy of the edge list
    if plan is None:
        plan = slam_ext.ba_plan(ii, jj, N, ht, wd, t0, t1, rank, world)
    eng = engine_cls(plan, poses, disps, intrinsics, disps_sens, targets, weights, eta, motion_only)
    if owned_inputs:
        n_own = int(plan.owned_edges().numel())
        if tuple(targets.shape) != (n_own, 2, ht, wd) or tuple(weights.shape) != (n_own, 2, ht, wd):
            raise RuntimeError(f"owned_inputs: targets/weights must be [{n_own},2,{ht},{wd}] (rows of plan.owned_edges())")
    peer, two_step, dist_solve = None, False, False
    if world > 1 and engine_cls is CudaShardEngine and collective in ("auto", "nvls", "nvls2", "dist"):
        if nvls_available(plan, poses.device, group):
            peer = PeerSystem.get(plan, poses.device, group)
            # the distributed factorisation pays one NVSwitch hop per tile column: worth it once the solve is throughput-bound
            dist_solve = collective == "dist" or (collective == "auto" and 6 * plan.P >= DIST_SOLVE_MIN_UNKNOWNS)
            two_step = not dist_solve and (collective == "nvls2" or (collective == "auto" and world >= 4))
        elif collective in ("nvls", "nvls2", "dist"):
            raise RuntimeError(f"collective='{collective}' needs NVSwitch multicast memory and more than 128 unknowns")
    if profile is not None:
        profile["plan"] = plan
        profile["collective"] = ("dist" if dist_solve else "nvls2" if two_step else "nvls") if peer is not None else "allreduce"
        ev = profile.setdefault("events", [])
    if hasattr(eng, "use_owned_rows"):
        eng.use_owned_rows(bool(owned_inputs))
    elif owned_inputs:
        raise RuntimeError("this engine does not take owner-only inputs")
    try:
        if two_step:  # accumulate in buffer 0, solve in this rank's instance of the reduced system
            eng.use_peer_buffer(peer.bufs[0].data_ptr(), peer.hdls[0].multicast_ptr)
            eng.use_solve_buffer(peer.bufs[2].data_ptr())
        if dist_solve:
            eng.use_dist_solve(peer.bufs[2].data_ptr(), peer.hdls[2].multicast_ptr, peer.bufs[3].data_ptr(), peer.hdls[3].multicast_ptr,
                               rank, world)
        for _ in range(int(iterations)):
            if profile is not None:
                e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                e[0].record()
            if peer is not None and not two_step:
                b = peer.k & 1
                eng.use_peer_buffer(peer.bufs[b].data_ptr(), peer.hdls[b].multicast_ptr)
                peer.k += 1
            system = eng.linearize()
            if profile is not None:
                e[1].record()
            if peer is not None:
                peer.hdls[0].barrier(channel=0)  # every rank's partial system is complete and visible
                if two_step:
                    eng.peer_reduce(peer.hdls[0].multicast_ptr, peer.hdls[2].multicast_ptr, rank, world)
                    peer.hdls[0].barrier(channel=0)  # every rank's share of the sum has landed everywhere
            elif world > 1:
                dist.all_reduce(system, op=dist.ReduceOp.SUM, group=group)
            if profile is not None:
                e[2].record()
            eng.solve_update(lm, ep)
            if profile is not None:
                e[3].record()
                ev.append(e)
    finally:
        if peer is not None:
            eng.use_peer_buffer(None, None)
            if two_step:
                eng.use_solve_buffer(None)
            if dist_solve:
                eng.use_dist_solve(None, None, None, None, 0, 1)
        if owned_inputs and hasattr(eng, "use_owned_rows"):
            eng.use_owned_rows(False)
    if exchange and world > 1 and not motion_only:
        exchange_owned_rows(plan, disps, eng.dz, group)
    return [eng.dx, eng.dz]


def exchange_owned_rows(plan, disps, dz, group=None):
    """Make disps[kx] and dz identical on all ranks: an all-gather of the rows each rank owns (shards differ in size, so the
    rows travel in slots of the largest shard)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    kx = plan.kx.to(disps.device)
    HW = plan.ht * plan.wd
    ranges = [plan.owned_range(r) for r in range(world)]
    m = max(hi - lo for lo, hi in ranges)
    if m == 0:
        return
    lo, hi = ranges[rank]
    mine = torch.zeros(2, m, HW, dtype=disps.dtype, device=disps.device)
    mine[0, : hi - lo] = disps.view(-1, HW)[kx[lo:hi]]
    mine[1, : hi - lo] = dz[lo:hi]
    allrows = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allrows, mine, group=group)
    dflat = disps.view(-1, HW)
    for r, (rlo, rhi) in enumerate(ranges):
        if r == rank or rhi == rlo:
            continue
        dflat[kx[rlo:rhi]] = allrows[r][0, : rhi - rlo]
        dz[rlo:rhi] = allrows[r][1, : rhi - rlo]
