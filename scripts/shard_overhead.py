"""Where does a sharded BA call spend its time outside the three stages?  torchrun --nproc-per-node 2"""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
from vipe_b200.distributed import ba_sharded, exchange_owned_rows
from vipe_b200.synthetic import make_problem
pr = make_problem("c3")
a = pr.args(dev)
p0, d0 = a[0].clone(), a[1].clone()
for ex in (True, False):
    for _ in range(3):
        a[0].copy_(p0); a[1].copy_(d0)
        ba_sharded(*a, exchange=ex)
    torch.cuda.synchronize(); dist.barrier()
    tot, stages, cpu = 0.0, 0.0, 0.0
    for _ in range(10):
        a[0].copy_(p0); a[1].copy_(d0)
        prof = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t = time.perf_counter()
        e0.record()
        ba_sharded(*a, exchange=ex, profile=prof)
        e1.record()
        cpu += time.perf_counter() - t
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
        ev = prof["events"]
        stages += sum(e[0].elapsed_time(e[3]) for e in ev)
        gaps = sum(ev[i][3].elapsed_time(ev[i + 1][0]) for i in range(len(ev) - 1))
        pre = e0.elapsed_time(ev[0][0]); post = ev[-1][3].elapsed_time(e1)
    if rank == 0:
        print(f"exchange={ex}: call {tot/10:.3f} ms  stages {stages/10:.3f} ms  cpu {cpu/10*1e3:.3f} ms | last call: pre {pre:.3f} gaps {gaps:.3f} post {post:.3f}", flush=True)
dist.destroy_process_group()
