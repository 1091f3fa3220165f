import sys, torch
sys.path.insert(0, ".")
from oracle import build_ref
from vipe_b200.ext import slam_ext
from vipe_b200.synthetic import make_problem, pose_errors
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
mod = build_ref.load()
dev = torch.device("cuda:0")
pr = make_problem(name)
a, b = pr.args(dev), pr.args(dev)
a[11] = b[11] = 1
dxr, dzr = mod.slam_ext.ba(*a)
dx, dz = slam_ext.ba(*b)
torch.cuda.synchronize()
print("ref dx: max", dxr.abs().max().item(), "nan", torch.isnan(dxr).any().item(), "norm", dxr.norm().item())
print("our dx: max", dx.abs().max().item(), "nan", torch.isnan(dx).any().item(), "norm", dx.norm().item())
print("ref dz: max", dzr.abs().max().item(), "nan", torch.isnan(dzr).any().item())
print("our dz: max", dz.abs().max().item(), "nan", torch.isnan(dz).any().item())
d = (dx - dxr).abs()
print("worst rows", d.max(dim=1).values.topk(5))
print("dx[0:3] ref", dxr[:3]); print("dx[0:3] our", dx[:3])
# motion-only comparison too
a, b = pr.args(dev), pr.args(dev)
a[11] = b[11] = 1; a[14] = b[14] = True
dxr, _ = mod.slam_ext.ba(*a)
dx, _ = slam_ext.ba(*b)
torch.cuda.synchronize()
print("motion-only: ref max", dxr.abs().max().item(), "our max", dx.abs().max().item(), "rel diff", ((dx-dxr).norm()/dxr.norm()).item())
