"""Generate golden vectors from the REFERENCE ITSELF (run on a GPU box):

    python tests/golden/make_golden.py        # writes tests/golden/ref_*.npz (copy back via gpurun_out/)

Runs the unmodified reference `slam_ext.ba` (oracle/_ref/vipe_ref_ext.so, built by oracle/build_ref.py from
/root/reference/csrc/slam_ext with oracle/eigen_stub standing in for Eigen) on the seeded synthetic configs C1 and C2
(full and motion-only) and stores its outputs: updated poses, dx, and every 16th pixel of the updated disparities and
of dz.  tests/test_golden.py checks the CPU oracle against these files, which is what pins the oracle.
"""

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import build_ref  # noqa: E402
from vipe_b200.synthetic import make_problem  # noqa: E402

STRIDE = 16


def main(out_dir: Path):
    mod = build_ref.load()
    assert mod is not None, "oracle/_ref/vipe_ref_ext.so missing"
    dev = torch.device("cuda:0")
    for name in ("c1", "c2"):
        for motion_only in (False, True):
            pr = make_problem(name)
            a = pr.args(dev)
            a[14] = motion_only
            out = mod.slam_ext.ba(*a)
            torch.cuda.synchronize()
            rec = {"poses": a[0].cpu().numpy(), "dx": out[0].cpu().numpy(),
                   "disps_sub": a[1].cpu().view(pr.cfg.n_frames, -1)[:, ::STRIDE].numpy()}
            if not motion_only:
                rec["dz_sub"] = out[1].cpu()[:, ::STRIDE].numpy()
            f = out_dir / f"ref_{name}_{'motion' if motion_only else 'full'}.npz"
            np.savez_compressed(f, **rec)
            print("wrote", f, {k: v.shape for k, v in rec.items()})


def geom(out_dir: Path):
    """projmap / frame_distance / depth_filter / iproj of the reference on the C2 problem (every 16th pixel)."""
    mod = build_ref.load()
    dev = torch.device("cuda:0")
    pr = make_problem("c2")
    p, d, k = pr.poses.to(dev), pr.disps.to(dev), pr.intrinsics.to(dev)
    coords, valid = mod.slam_ext.projmap(p, d, k, pr.ii.to(dev), pr.jj.to(dev))
    gen = torch.Generator().manual_seed(3)
    M = 40
    pi = torch.randint(0, 16, (M,), generator=gen)
    pj = torch.randint(0, 16, (M,), generator=gen)
    K2 = torch.stack([pr.intrinsics, pr.intrinsics * 1.01]).to(dev)
    qi = torch.randint(0, 2, (M,), generator=gen)
    qj = torch.randint(0, 2, (M,), generator=gen)
    dist = mod.slam_ext.frame_distance(p, d, K2, pi.to(dev), pj.to(dev), qi.to(dev), qj.to(dev), pi.to(dev), 0.3)
    ix = torch.tensor([0, 1, 5, 8, 14, 15])
    thresh = torch.tensor([0.05, 0.1, 0.2, 0.4, 0.8, 1.6])
    counter = mod.slam_ext.depth_filter(pr.poses_gt.to(dev), pr.disps_gt.to(dev), k, ix.to(dev), thresh.to(dev))
    pts = mod.slam_ext.iproj(p, d, k)
    torch.cuda.synchronize()
    rec = {"coords_sub": coords.cpu().view(120, -1, 3)[:, ::STRIDE].numpy(), "valid_sub": valid.cpu().view(120, -1)[:, ::STRIDE].numpy(),
           "pi": pi.numpy(), "pj": pj.numpy(), "qi": qi.numpy(), "qj": qj.numpy(), "dist": dist.cpu().numpy(),
           "ix": ix.numpy(), "thresh": thresh.numpy(), "counter_sub": counter.cpu().view(6, -1)[:, ::STRIDE].numpy(),
           "points_sub": pts.cpu().view(16, -1, 3)[:, ::STRIDE].numpy()}
    f = out_dir / "ref_geom_c2.npz"
    np.savez_compressed(f, **rec)
    print("wrote", f, {k_: v.shape for k_, v in rec.items()})


if __name__ == "__main__":
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    main(out)
    geom(out)
