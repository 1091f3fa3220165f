#!/usr/bin/env python
"""bench.py -- dense bundle adjustment throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5]

A "step" is one full `slam_ext.ba` call (all Gauss-Newton iterations of the workload) on one synthetic
problem.  Default workload: C3, the backend global BA (300 keyframes, 3000 edges, 48x64, 8 iterations,
lm=1e-5, ep=1e-2) -- the configuration BASELINE.json quotes "at 1/2/4/8 B200".  With N > 1 the same problem
is sharded by source keyframe over the ranks (strong scaling; one all-reduce of the reduced camera system
per iteration).  Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# torchrun exports OMP_NUM_THREADS=1 to its workers, and the OpenMP runtime reads it when it is loaded (with torch): the reference
# arm's host code (Schur assembly, the LLT stand-in) is OpenMP code and gets every core of the box, as when it is run directly --
# so the variable is fixed BEFORE torch is imported (and the runtime is told again in run_reference)
if "--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1: sys.argv.index("--impl") + 2] == ["reference"]:
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "dense_ba_gn_iterations_per_sec"
UNIT = "iter/s"


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled during the timed region (B200_PROFILING.md recipe): NVML every 20 ms where
    the bindings load (the device-timed region lasts ~0.1 s; one `nvidia-smi` process start takes longer than that),
    else `nvidia-smi` every 0.2 s."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()
        self.nvml, self.handle, self.source = None, None, "nvidia-smi"
        try:
            import pynvml

            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; NVML's does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
        return [str(sm), str(mx), "", *["Active" if r & b else "Not Active" for b in bits]]

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.02 if self.nvml is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = sorted({self.NAMES[i] for s in self.samples if len(s) >= 7 for i in range(4) if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "source": self.source}


def algorithmic_bytes(cfg, E, K, N, HW, motion_only):
    """unique compulsory bytes of one GN iteration, and of the linearise kernel alone."""
    if motion_only:
        it = 16 * E * HW + 4 * K * HW + 56 * N + 16 * E
        lin = 16 * E * HW + 4 * K * HW + 28 * N + 8 * E
    else:
        it = 16 * E * HW + 16 * K * HW + 56 * N + 16 * E
        lin = 16 * E * HW + 12 * K * HW + 28 * N + 8 * E  # targets+weights, disps+disps_sens+eta, poses, edge list
    return it, lin


def cpu_baseline(pr, seconds_budget=25.0):
    """fp32 twin of the oracle (pure torch CPU ops, same math as ba_cuda) on a bounded sample of the workload."""
    from oracle import ba_oracle

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = pr.cfg
    a = pr.args()
    a[11] = 1
    t = time.perf_counter()
    ba_oracle.ba(*a, dtype=torch.float32)
    t1 = time.perf_counter() - t
    iters = max(1, min(cfg.iters, int(seconds_budget / max(t1, 1e-6))))
    a = pr.args()
    a[11] = iters
    t = time.perf_counter()
    ba_oracle.ba(*a, dtype=torch.float32)
    dt = time.perf_counter() - t
    return {"value": iters / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{cfg.name}: one ba call of {iters} GN iteration(s) of the full problem, fp32 torch CPU ops, after a 1-iteration warm-up"}


def run_reference(args, pr):
    """--impl reference: the reference's own slam_ext.ba (CUDA kernels + its host Schur/solve code, compiled
    unmodified from /root/reference with oracle/eigen_stub standing in for Eigen) on the same config; if that build
    is absent, the torch CPU port of the same algorithm."""
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference's host code (Schur assembly, the LLT stand-in) is
    # OpenMP code and is given every core of the box, as when it is run directly
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    torch.set_num_threads(os.cpu_count() or 1)  # omp_set_num_threads on the runtime that is already loaded
    from oracle import build_ref

    cfg = pr.cfg
    sample_iters = cfg.iters  # the same GN iterations per call as this repo's arm
    mod = build_ref.load() if torch.cuda.is_available() else None
    HW = cfg.ht * cfg.wd
    E = pr.ii.numel()
    if mod is not None:
        dev = torch.device("cuda:0")
        base = pr.args(dev)

        def step():
            a = [x.clone() if torch.is_tensor(x) and i < 2 else x for i, x in enumerate(base)]
            a[11] = sample_iters
            mod.slam_ext.ba(*a)

        import ctypes

        llt_clock = None
        try:
            llt_clock = ctypes.CDLL(str(build_ref.so_path())).vipe_eigen_stub_llt_seconds
            llt_clock.restype = ctypes.c_double
            llt_clock.argtypes = [ctypes.c_int]
        except Exception:
            llt_clock = None
        steps = max(1, min(args.steps, 20))
        for _ in range(min(args.warmup, 2)):
            step()
        torch.cuda.synchronize()
        if llt_clock:
            llt_clock(1)
        t = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        llt_s = llt_clock(0) if llt_clock else None
        args.steps = steps
        kind, cores = "reference", os.cpu_count()
        sample = (f"{cfg.name}: {steps} ba calls of {sample_iters} GN iterations each (full problem), reference CUDA kernels on "
                  f"one B200 + reference host Schur/solve code on {cores} host threads (dense-LLT Eigen stand-in), wall clock")
        breakdown = None
        if llt_s is not None:
            per_it = dt / (steps * sample_iters)
            breakdown = {"seconds_per_iteration": per_it, "stand_in_llt_seconds_per_iteration": llt_s / (steps * sample_iters),
                         "stand_in_llt_share": llt_s / dt,
                         "note": "the LLT stand-in (oracle/eigen_stub, dense, OpenMP) is NOT reference code: Eigen's SimplicialLLT is "
                                 "absent from the image; the rest of the time is the reference's own CUDA kernels and host code"}
    else:
        from oracle import ba_oracle

        torch.set_num_threads(os.cpu_count() or 1)

        def step():
            a = pr.args()
            a[11] = sample_iters
            ba_oracle.ba(*a, dtype=torch.float32)

        steps = max(1, min(args.steps, 3))
        step()
        t = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t
        args.steps = steps
        kind, cores = "port", torch.get_num_threads()
        breakdown = None
        sample = f"{cfg.name}: {steps} ba calls of {sample_iters} GN iterations each, fp32 torch CPU port of ba_cuda"
    value = args.steps * sample_iters / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "edge_pixels_per_sec": value * E * HW,
            "config": {"workload": workload_name(cfg), "frames": cfg.n_frames, "edges": E, "ht": cfg.ht, "wd": cfg.wd,
                       "gn_iterations_per_step": sample_iters, "lm": cfg.lm, "ep": cfg.ep, "motion_only": cfg.motion_only},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if breakdown:
        line["reference_breakdown"] = breakdown
    print(json.dumps(line), flush=True)


def side_measurement(name, dev, calls=20):
    """A second, much smaller configuration reported beside the headline one (C2, the frontend window: latency-bound,
    whole working set in L2): BA calls per second and GN iterations per second with inputs resident on the device."""
    from vipe_b200.ext import slam_ext
    from vipe_b200.synthetic import make_problem

    pr = make_problem(name)
    a = pr.args(dev)
    p0, d0 = a[0].clone(), a[1].clone()
    for _ in range(5):
        a[0].copy_(p0)
        a[1].copy_(d0)
        slam_ext.ba(*a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(calls):
        a[0].copy_(p0)
        a[1].copy_(d0)
        slam_ext.ba(*a)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / calls
    return {"workload": workload_name(pr.cfg), "ms_per_call": ms, "gn_iterations_per_sec": pr.cfg.iters / ms * 1e3,
            "edge_pixels_per_sec": pr.edge_pixels * pr.cfg.iters / ms * 1e3}


def side_c4(dev, rank, world, collective, steps=3):
    """C4 (1000 keyframes, 12000 edges, 64x112, 8 GN iterations) beside the headline: one GPU through slam_ext.ba, N GPUs
    keyframe-sharded with owner-only inputs.  Device-timed, inputs resident, max over ranks."""
    import torch.distributed as dist

    from vipe_b200.distributed import ba_sharded
    from vipe_b200.ext import slam_ext
    from vipe_b200.plan import cached_plan
    from vipe_b200.synthetic import make_problem

    pr = make_problem("c4")
    cfg = pr.cfg
    a = pr.args(dev)
    kw = {}
    if world > 1:
        plan = cached_plan(pr.ii.contiguous(), pr.jj.contiguous(), cfg.n_frames, cfg.ht, cfg.wd, pr.t0, pr.t1, rank, world)
        own = plan.owned_edges()
        a[4], a[5] = pr.targets[own].contiguous().to(dev), pr.weights[own].contiguous().to(dev)
        prof = {"events": []}
        kw = dict(plan=plan, owned_inputs=True, collective=collective, profile=prof)
    p0, d0 = a[0].clone(), a[1].clone()

    def run():
        a[0].copy_(p0)
        a[1].copy_(d0)
        if world > 1:
            ba_sharded(*a, **kw)
        else:
            slam_ext.ba(*a)

    for _ in range(2):
        run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a[0].copy_(p0)
        a[1].copy_(d0)
        e0.record()
        if world > 1:
            ba_sharded(*a, **kw)
        else:
            slam_ext.ba(*a)
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out = {"workload": workload_name(cfg), "n_gpus": world, "ms_per_call": ms / steps,
           "gn_iterations_per_sec": cfg.iters * steps / ms * 1e3, "edge_pixels_per_sec": pr.edge_pixels * cfg.iters * steps / ms * 1e3,
           "steps": steps}
    if world > 1:
        evs = prof["events"][-steps * cfg.iters:]
        st = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / len(evs) for i in range(3)]
        out["stage_ms_per_iteration"] = {"linearize_schur_assemble": st[0], "reduce": st[1], "solve_backsub_retract": st[2]}
        out["collective"] = prof.get("collective")
    return out


def side_c5(dev, calls=10):
    """C5: motion-only BA of 64 independent clips (16 keyframes, 120 edges each), one batched call (slam_ext.ba_batch)."""
    from vipe_b200.ext import slam_ext
    from vipe_b200.synthetic import CONFIGS, make_problem

    cfg = CONFIGS["c5"]
    problems = [make_problem(cfg, clip=c) for c in range(cfg.clips)]
    pr, N, nc = problems[0], cfg.n_frames, cfg.clips
    cat = lambda xs: torch.cat(xs, dim=0).to(dev)
    a = [cat([p.poses for p in problems]), cat([p.disps for p in problems]), pr.intrinsics.to(dev),
         cat([p.disps_sens for p in problems]), cat([p.targets for p in problems]), cat([p.weights for p in problems]),
         cat([p.eta for p in problems]), cat([p.ii + c * N for c, p in enumerate(problems)]),
         cat([p.jj + c * N for c, p in enumerate(problems)]), [c * N for c in range(nc + 1)], [c * N + pr.t0 for c in range(nc)],
         [c * N + pr.t1 for c in range(nc)], cfg.iters, cfg.lm, cfg.ep, cfg.motion_only]
    p0 = a[0].clone()
    for _ in range(3):
        a[0].copy_(p0)
        slam_ext.ba_batch(*a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(calls):
        a[0].copy_(p0)
        slam_ext.ba_batch(*a)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / calls
    return {"workload": workload_name(cfg), "ms_per_call": ms, "gn_iterations_per_sec": cfg.iters * nc / ms * 1e3,
            "edge_pixels_per_sec": pr.edge_pixels * nc * cfg.iters / ms * 1e3, "how": "64 clips in one batched plan, inputs resident"}


def workload_name(cfg):
    names = {"c1": "C1 synthetic dense BA: 8 keyframes, 24 edges, 48x64, 2 GN iters",
             "c2": "C2 frontend local BA window: 16 keyframes, 120 edges, 48x64, 4 GN iters",
             "c3": "C3 backend global BA: 300 keyframes, 3000 edges, 48x64, 8 GN iters",
             "c4": "C4 large global BA: 1000 keyframes, 12000 edges, 64x112, 8 GN iters",
             "c5": "C5 motion-only BA: 64 clips x (16 keyframes, 120 edges, 48x64), 4 GN iters"}
    return names[cfg.name]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the C4 / C5 / C2 side measurements")
    ap.add_argument("--collective", default="auto", choices=["auto", "nvls", "nvls2", "dist", "allreduce"],
                    help="N > 1: how the ranks' partial reduced systems are summed (vipe_b200/distributed.py)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from vipe_b200.synthetic import CONFIGS, make_problem

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference(args, make_problem(args.workload))
        return

    from vipe_b200.distributed import ba_sharded
    from vipe_b200.ext import slam_ext
    from vipe_b200 import _lib
    import ctypes as C

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; stdout carries exactly one JSON line, so fd 1
        # points at stderr until that line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = CONFIGS[args.workload]
    clips = cfg.clips
    # C5: the 64 clips are independent -> each rank takes its share of the clips (replicas, no collective)
    my_clips = list(range(rank, clips, world)) if clips > 1 else [0]
    problems = [make_problem(cfg, clip=c) for c in my_clips]
    pr = problems[0]
    HW, E, N = cfg.ht * cfg.wd, pr.ii.numel(), cfg.n_frames
    sharded = world > 1 and clips == 1

    batched = clips > 1
    if batched:
        # C5: this rank's clips are concatenated and solved by ONE batched call (slam_ext.ba_batch)
        nc = len(problems)
        cat = lambda xs: torch.cat(xs, dim=0).to(dev)
        b_ii = cat([p.ii + c * N for c, p in enumerate(problems)])
        b_jj = cat([p.jj + c * N for c, p in enumerate(problems)])
        batch_args = [cat([p.poses for p in problems]), cat([p.disps for p in problems]), pr.intrinsics.to(dev),
                      cat([p.disps_sens for p in problems]), cat([p.targets for p in problems]),
                      cat([p.weights for p in problems]), cat([p.eta for p in problems]), b_ii, b_jj,
                      [c * N for c in range(nc + 1)], [c * N + pr.t0 for c in range(nc)], [c * N + pr.t1 for c in range(nc)],
                      cfg.iters, cfg.lm, cfg.ep, cfg.motion_only]
        dev_args = [batch_args]
    else:
        dev_args = [p.args(dev) for p in problems]
    shard_plan, own_rows = None, None
    if sharded:
        # every rank HOLDS only the targets/weights rows of its own edges, in the plan's CSR order
        from vipe_b200.plan import cached_plan

        shard_plan = cached_plan(pr.ii.contiguous(), pr.jj.contiguous(), N, cfg.ht, cfg.wd, pr.t0, pr.t1, rank, world)
        own_rows = shard_plan.owned_edges()
        dev_args[0][4] = pr.targets[own_rows].contiguous().to(dev)
        dev_args[0][5] = pr.weights[own_rows].contiguous().to(dev)
    init_state = [(a[0].clone(), a[1].clone()) for a in dev_args]
    plans = [slam_ext.ba_plan(p.ii, p.jj, N, cfg.ht, cfg.wd, p.t0, p.t1) for p in problems] if not (sharded or batched) else []
    K = int(torch.unique(torch.cat([torch.arange(pr.t0, pr.t1), pr.ii])).numel())

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def reset():
        for a, (p0, d0) in zip(dev_args, init_state):
            a[0].copy_(p0)
            a[1].copy_(d0)

    shard_prof = {"events": []} if sharded else None

    def step():
        for a in dev_args:
            if sharded:
                ba_sharded(*a, exchange=True, profile=shard_prof, collective=args.collective, plan=shard_plan, owned_inputs=True)
            elif batched:
                slam_ext.ba_batch(*a)
            else:
                slam_ext.ba(*a)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        reset()
        step()
    barrier()
    if batched:
        plans = list(slam_ext._BATCH_PLANS.values())[-1:]  # the plan the warm-up just built

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_ms = [0.0, 0.0, 0.0, 0.0]
    stage_iters = 0
    barrier()
    wall0 = time.perf_counter()
    for s in range(args.steps):
        reset()
        flush_buf.zero_()  # L2 flush between timed steps (outside the event bracket)
        ev[s][0].record()
        step()
        ev[s][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = sum(pl.launch_count for pl in plans) * args.steps if plans else None
    # per-stage times come from a separate, untimed pass: stage events switch the CUDA-graph replay off
    for pl in plans:
        _lib.check(_lib.lib().vipe_ba_profile_enable(pl.handle, 1), "profile_enable")
    for s in range(min(args.steps, 5) if plans else 0):
        reset()
        flush_buf.zero_()
        step()
        torch.cuda.synchronize()
        for pl in plans:
            ms4, its = (C.c_float * 4)(), C.c_int()
            _lib.check(_lib.lib().vipe_ba_profile_read(pl.handle, C.byref(ms4), C.byref(its)), "profile_read")
            stage_ms = [x + y for x, y in zip(stage_ms, ms4)]
            stage_iters += its.value
    # stages 1-2 alone (per-edge Jacobians/Hessians, no Schur Gram): the motion-only variant of the same kernel
    jac_ms = None
    if plans and not batched and not cfg.motion_only:
        tot, its_tot = 0.0, 0
        for s in range(5):
            reset()
            flush_buf.zero_()
            a = list(dev_args[0])
            a[11], a[14] = 1, True
            slam_ext.ba(*a)
            torch.cuda.synchronize()
            ms4, its = (C.c_float * 4)(), C.c_int()
            _lib.check(_lib.lib().vipe_ba_profile_read(plans[0].handle, C.byref(ms4), C.byref(its)), "profile_read")
            if s > 0:  # the first call compiles nothing but warms the motion-only kernel's instruction cache
                tot += ms4[0]
                its_tot += its.value
        jac_ms = tot / max(its_tot, 1)
    for pl in plans:
        _lib.check(_lib.lib().vipe_ba_profile_enable(pl.handle, 0), "profile_enable")
    shard_stage = None
    if sharded:
        evs = shard_prof["events"][-args.steps * cfg.iters:]
        shard_stage = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / len(evs) for i in range(3)]
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    if sharded:
        # the phased entry points accumulate their launch count in the plan: per step = total / (warm-up + timed steps)
        launches = int(shard_prof["plan"].launch_count * args.steps / (args.steps + args.warmup))

    iters_total = args.steps * cfg.iters * clips
    value = iters_total / (total_ms / 1e3)
    edge_px = E * HW
    nprob = len(problems) if batched else 1  # a batched step processes this rank's clips in one launch set
    it_bytes, lin_bytes = algorithmic_bytes(cfg, E * nprob, K * nprob, N * nprob, HW, cfg.motion_only)
    peak, peak_src = measured_peaks()

    # fp32 work of the fused kernel (it is FMA-bound, not HBM-bound, at backend degrees): ~110 FMA-pipe lane-ops per
    # edge-pixel for the linearisation + 36 per block pair and frame-pixel for the Schur Gram
    deg = torch.bincount(pr.ii, minlength=N).double()
    gram_ops = float((36.0 * deg * (deg + 1) / 2 + 12.0 * deg).sum()) * HW
    lin_fma_ops = (110.0 * E * HW + (0.0 if cfg.motion_only else gram_ops)) * nprob
    FP64_DMMA_PEAK_TFLOPS = 37.2  # measured (scripts/dmma_lat.cu): 64 DMMA FMA/clk/SM x 148 SMs x 1.965 GHz x 2
    FP32_PEAK_TFMA = 34.3  # measured on this pool: FFMA/FFMA2 full-chip micro-benchmark (scripts/ffma2_micro.cu), TFMA/s
    roofline = None
    if sharded:
        # this rank's share of the edge-pixels (owned source frames) over its own linearise time
        pl = shard_prof["plan"]
        lo, hi = pl.owned_range()
        ptrs, _ = pl.csr()
        own_E = int(ptrs[hi] - ptrs[lo])
        own_bytes = 16 * own_E * HW + 12 * (hi - lo) * HW + 28 * N + 8 * own_E
        lin_ms = shard_stage[0]
        achieved = own_bytes / (lin_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "vba::linearize2_kernel on rank 0's shard (linearise+Schur+assemble stage)",
                    "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "algorithmic_bytes_per_launch": own_bytes, "avg_launch_ms": lin_ms,
                    "collective": {"nvls": "none: the Cholesky kernel sums the ranks' partial systems through the NVSwitch (multimem.ld_reduce); "
                                           "'all_reduce' below is the cross-rank barrier",
                                   "nvls2": "in-switch reduce-scatter + multicast kernel (multimem.ld_reduce / multimem.st) between two "
                                            "cross-rank barriers ('all_reduce' below), then a local solve",
                                   "dist": "none: the factorisation itself is distributed -- tile columns block-cyclic over the ranks, the owner's "
                                           "loads sum the partial systems in the switch (multimem.ld_reduce) and L, y, L_jj^-T are multicast to "
                                           "every rank (multimem.st) as self-validating words; 'all_reduce' below is the cross-rank barrier",
                                   }.get(shard_prof.get("collective"), "NCCL all_reduce of the reduced camera system"),
                    "stage_ms_per_iteration": {"linearize_schur_assemble": shard_stage[0], "all_reduce": shard_stage[1],
                                               "solve_backsub_retract": shard_stage[2]}}
    elif stage_iters > 0:
        lin_ms = stage_ms[0] / stage_iters  # average duration of one linearise launch (events around the kernel, after the system clear)
        achieved = lin_bytes / (lin_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "vba::linearize2_kernel<256,false> (per-source-frame Jacobian/Hessian + Schur Gram; the system clear is not in avg_launch_ms)",
                    "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "algorithmic_bytes_per_launch": lin_bytes, "avg_launch_ms": lin_ms,
                    "stage_ms_per_iteration": {"linearize_schur": stage_ms[0] / stage_iters, "assemble": stage_ms[1] / stage_iters,
                                               "solve": stage_ms[2] / stage_iters, "backsub_retract": stage_ms[3] / stage_iters},
                    "compute": {"bound": "fp32 FMA pipe", "fma_lane_ops_per_launch": lin_fma_ops,
                                "achieved_tfma_s": lin_fma_ops / (lin_ms * 1e-3) / 1e12, "peak_tfma_s": FP32_PEAK_TFMA,
                                "frac": lin_fma_ops / (lin_ms * 1e-3) / 1e12 / FP32_PEAK_TFMA,
                                "note": "the fused Jacobian+Schur kernel is FMA-bound at backend degrees; see DESIGN.md section 4"},
                    "whole_iteration": {"algorithmic_bytes": it_bytes,
                                        "achieved_gbs": it_bytes * cfg.iters * (clips // nprob if batched else clips) * args.steps / (total_ms * 1e-3) / 1e9 / max(world, 1)}}
        if not cfg.motion_only and clips == 1:
            # the largest share of the iteration: the fp64 Cholesky of the reduced camera system (solve stage = memset + factor +
            # backward substitution), against the fp64 tensor (DMMA) rate measured on B200
            n_sys = 6 * (pr.t1 - pr.t0)
            solve_ms = stage_ms[2] / stage_iters
            roofline["solver"] = {"kernel": "vba::chol_factor_kernel + chol_backward_kernel (tile-dataflow fp64 Cholesky, DMMA)", "bound": "tensor (fp64)",
                                  "unknowns": n_sys, "flop": n_sys ** 3 / 3.0, "avg_stage_ms": solve_ms,
                                  "achieved": n_sys ** 3 / 3.0 / (solve_ms * 1e-3) / 1e12, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                                  "frac": n_sys ** 3 / 3.0 / (solve_ms * 1e-3) / 1e12 / FP64_DMMA_PEAK_TFLOPS,
                                  "note": "C3: bound by the latency of the chain of tile columns, not by the pipe; C4: throughput-bound"}
        if jac_ms:
            _, jac_bytes = algorithmic_bytes(cfg, E, K, N, HW, True)
            roofline["jacobian_stage_alone"] = {
                "kernel": "vba::lin3_motion_kernel (stages 1-2 only: projective transform, J_j, per-edge H/v; TMA-fed, no Schur Gram)",
                "avg_launch_ms": jac_ms, "algorithmic_bytes_per_launch": jac_bytes,
                "achieved": jac_bytes / (jac_ms * 1e-3) / 1e9, "unit": "GB/s", "frac": jac_bytes / (jac_ms * 1e-3) / 1e9 / peak}
        prof = ROOT / "profiles" / "traffic.json"
        if prof.is_file():
            try:
                roofline["traffic"] = json.loads(prof.read_text()).get(args.workload, {}).get("linearize_dram_bytes_per_launch")
            except Exception:
                pass

    # end-to-end through the public API with host buffers (rank-local inputs in pinned memory)
    if batched:
        host_args = [[x.cpu().pin_memory() if torch.is_tensor(x) else x for x in batch_args]]
    elif sharded:
        h = pr.args()
        h[4], h[5] = pr.targets[own_rows].contiguous(), pr.weights[own_rows].contiguous()  # this rank uploads its own edges only
        host_args = [[x.pin_memory() if torch.is_tensor(x) else x for x in h]]
    else:
        host_args = [[x.pin_memory() if torch.is_tensor(x) else x for x in p.args()] for p in problems]
    h2d = sum(x.numel() * x.element_size() for h in host_args for x in h if torch.is_tensor(x))
    d2h = sum((h[0].numel() + h[1].numel()) * 4 for h in host_args)
    out_host = [(torch.empty_like(h[0]).pin_memory(), torch.empty_like(h[1]).pin_memory()) for h in host_args]

    # every step uploads its own inputs from pinned host memory and reads its result back, all inside the timed region;
    # the upload of step s+1 runs on a copy stream under the solve of step s (vipe_b200/host_feed.py)
    from vipe_b200.host_feed import HostFeed

    feed = HostFeed(dev)
    if sharded:
        e2e_fn = lambda *a: ba_sharded(*a, exchange=True, collective=args.collective, plan=shard_plan, owned_inputs=True)
    elif batched:
        e2e_fn = slam_ext.ba_batch
    else:
        e2e_fn = None  # HostFeed's default: slam_ext.ba with the plan taken from the host copy of the edge list
    todo = list(zip(host_args, out_host))

    def e2e_run(nsteps):
        seq = todo * nsteps
        feed.prefetch(seq[0][0])
        for q, (h, (op, od)) in enumerate(seq):
            if q + 1 < len(seq):
                feed.prefetch(seq[q + 1][0])
            feed.run(fn=e2e_fn, out_poses=op, out_disps=od)

    e2e_steps = max(3, min(args.steps, 10))
    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(e2e_steps)
    torch.cuda.current_stream().wait_event(feed.last_done)  # the last step's read-back (third stream) is inside the timed region
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = e2e_steps * cfg.iters * clips / (e2e_ms / 1e3)
    # the sampler ran from the start of the timed region to the end of the end-to-end leg (the device-timed region alone
    # lasts ~0.1 s, one or two nvidia-smi samples)
    clocks = sampler.stop()

    # the sharded result against a single-GPU run of the same operator on rank 0 (the only place the driver sees it)
    sharded_parity = None
    if sharded:
        from vipe_b200.synthetic import disp_error, pose_errors

        reset()
        step()
        barrier()
        if rank == 0:
            full = pr.args(dev)
            slam_ext.ba(*full)
            torch.cuda.synchronize()
            te, re_ = pose_errors(dev_args[0][0], full[0], pr.t0, pr.t1)
            kxs = torch.unique(torch.cat([torch.arange(pr.t0, pr.t1), pr.ii]))
            sharded_parity = {"vs": "slam_ext.ba on one GPU, same inputs, same iterations", "translation_rel": te, "rotation_max_rad": re_,
                              "disparity_rel": disp_error(dev_args[0][1], full[1], kxs),
                              "bounds": {"translation_rel": 1e-4, "rotation_max_rad": 1e-4, "disparity_rel": 1e-3}}
            del full
        barrier()
    side = {}
    if not args.no_side and args.workload == "c3":
        del dev_args, init_state, host_args, out_host, feed
        torch.cuda.empty_cache()
        side["c4"] = side_c4(dev, rank, world, args.collective)
        if world == 1:
            side["c5"] = side_c5(dev)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if clips == 1 else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "edge_pixels_per_sec": value * edge_px,
                "config": {"workload": workload_name(cfg), "frames": N, "edges": E, "ht": cfg.ht, "wd": cfg.wd,
                           "gn_iterations_per_step": cfg.iters, "lm": cfg.lm, "ep": cfg.ep, "motion_only": cfg.motion_only,
                           "clips": clips, "parallelism": (f"keyframe-sharded x{world}, " + {"nvls": "in-switch reduction fused into the solve (NVLS)", "nvls2": "in-switch reduce-scatter + multicast (NVLS), local solve", "dist": "distributed Cholesky over NVSwitch multicast, input summed in the switch"}.get(shard_prof.get("collective"), "NCCL all-reduce per iteration") + ", owner-only targets/weights") if sharded else ("clips batched per rank, no collective" if clips > 1 else "single"),
                           "l2": "flushed between timed steps (256 MB write)", "timing": "cuda events per step, max over ranks",
                           "wall_s_timed_region": wall},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                        "how": "slam_ext.ba fed from pinned host buffers by vipe_b200.host_feed.HostFeed: every step's H2D (copy stream, "
                               "overlapping the previous step's solve) and D2H of poses+disps are inside the timed region"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
        if sharded_parity is not None:
            line["sharded_parity"] = sharded_parity
        for k2, v2 in side.items():
            line[k2] = v2
        if world == 1 and args.workload != "c2" and not args.no_side:
            line["frontend_c2"] = side_measurement("c2", dev)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(pr)
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
