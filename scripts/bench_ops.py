"""Throughput of the four streaming slam_ext operators on a B200 (CUDA events, L2 flushed, C3-sized inputs).
Prints one JSON line per op: achieved GB/s of algorithmic bytes vs the measured HBM peak."""
import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vipe_b200.ext import slam_ext
from vipe_b200.synthetic import make_problem

peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").is_file() else 6650.0
dev = torch.device("cuda:0")
pr = make_problem("c3")
p, d, k = pr.poses.to(dev), pr.disps.to(dev), pr.intrinsics.to(dev)
ii, jj = pr.ii.to(dev), pr.jj.to(dev)
N, ht, wd = d.shape
HW, E = ht * wd, ii.numel()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
K2 = k[None].repeat(2, 1).contiguous()
zi = torch.zeros(E, dtype=torch.int64, device=dev)
ix = torch.arange(N, device=dev)
th = torch.full((N,), 0.1, device=dev)
ops = {
    "projmap": (lambda: slam_ext.projmap(p, d, k, ii, jj), E * HW * (4 + 16)),
    "frame_distance": (lambda: slam_ext.frame_distance(p, d, K2, ii, jj, zi, zi, ii, 0.3), E * HW * 4),
    "depth_filter": (lambda: slam_ext.depth_filter(p, d, k, ix, th), N * HW * (4 + 4 + 6 * 16)),
    "iproj": (lambda: slam_ext.iproj(p, d, k), N * HW * (4 + 12)),
}
for name, (fn, nbytes) in ops.items():
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(json.dumps({"op": name, "ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / ms / 1e6, "frac_of_measured_hbm": nbytes / ms / 1e6 / peak,
                      "note": "includes the torch allocation of the outputs; depth_filter bytes count its gathers as 16 B per neighbour"}))
