"""Run the linearisation (vipe_ba_linearize) of one config a few times: the command profiled with ncu for the per-kernel
launch list and the full capture of the linearise kernels.  Usage: python scripts/run_lin.py c3 [reps] [motion]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from vipe_b200 import _lib  # noqa: E402
from vipe_b200.ext import slam_ext  # noqa: E402
from vipe_b200.synthetic import make_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
motion = len(sys.argv) > 3 and sys.argv[3] == "motion"
pr = make_problem(name)
cfg = pr.cfg
dev = torch.device("cuda:0")
a = pr.args(dev)
plan = slam_ext.ba_plan(pr.ii, pr.jj, cfg.n_frames, cfg.ht, cfg.wd, pr.t0, pr.t1)
ws = plan.workspace(dev)
dx = torch.zeros(plan.P, 6, device=dev)
dz = torch.zeros(plan.K, cfg.ht * cfg.wd, device=dev)
tens = slam_ext._tensors(a[0], a[1], a[2], a[3], a[4], a[5], a[6], dx, dz, motion)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
for r in range(reps):
    flush.fill_(r)
    ev[2 * r].record()
    _lib.check(_lib.lib().vipe_ba_linearize(plan.handle, C.byref(tens), ws.data_ptr(), int(motion), st), "linearize")
    ev[2 * r + 1].record()
torch.cuda.synchronize()
print(name, "linearize (clear + linearise + assemble) ms:", [round(ev[2 * r].elapsed_time(ev[2 * r + 1]), 4) for r in range(reps)])
