"""Debug: adapter at C3 size with / without the focal variable."""
import sys
import torch
sys.path.insert(0, ".")
from vipe_b200 import adapter
from vipe_b200.synthetic import make_problem, pose_errors

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
import dataclasses
from vipe_b200.synthetic import CONFIGS
pr = make_problem(dataclasses.replace(CONFIGS["c3"], trajectory="orbit") if name == "c3o" else name)
cfg = pr.cfg
dev = torch.device("cuda:0")
E, HW = pr.ii.numel(), cfg.ht * cfg.wd
target = pr.targets.reshape(E, 2, HW).permute(0, 2, 1).contiguous().to(dev)
weight = pr.weights.reshape(E, 2, HW).permute(0, 2, 1).contiguous().to(dev)
for focal, scale, iters in [(False, 1.0, 8), (True, 1.0, 1), (True, 1.0, 8), (True, 1.03, 1), (True, 1.03, 2), (True, 1.03, 8)]:
    intr = (pr.intrinsics * 8.0).to(dev)
    intr[:2] *= scale
    poses, disps = pr.poses.clone().to(dev), pr.disps.clone().to(dev)
    damp = (0.01 * torch.nn.functional.softplus(torch.randn(cfg.n_frames, cfg.ht, cfg.wd, generator=torch.Generator().manual_seed(77)))).to(dev)  # droid_net.py:410
    dx, dz = adapter.bundle_adjustment(poses, disps, pr.disps_sens.to(dev), intr, target, weight, damp, pr.ii.to(dev), pr.jj.to(dev), 1,
                                       cfg.n_frames, iters, cfg.lm, cfg.ep, False, False, optimize_intrinsics=focal)
    torch.cuda.synchronize()
    te, re_ = pose_errors(poses, pr.poses_gt.to(dev), 1, cfg.n_frames)
    print(f"focal={focal} scale={scale} iters={iters}: fx={float(intr[0]):.3f} (gt {float(pr.intrinsics[0])*8:.3f}) te={te:.3e} re={re_:.3e} |dx|max={float(dx.abs().max()):.3e}")
