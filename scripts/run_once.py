"""Run slam_ext.ba a few times on one synthetic config (target for ncu).  Usage: run_once.py c3 [calls] [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from vipe_b200.ext import slam_ext
from vipe_b200.synthetic import make_problem

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pr = make_problem(name)
dev = torch.device("cuda:0")
for _ in range(calls):
    a = pr.args(dev)
    if len(sys.argv) > 3:
        a[11] = int(sys.argv[3])
    slam_ext.ba(*a)
torch.cuda.synchronize()
print("done", name)
