"""Probe: does torch symmetric memory work on this box (peer pointers, multicast, barrier)?  torchrun --nproc-per-node 2"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1 << 20, dtype=torch.float64, device=dev)
t.zero_()
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs], "mc", hex(h.multicast_ptr) if h.multicast_ptr else 0, "signal pads", [hex(p) for p in h.signal_pad_ptrs], "pad size", h.signal_pad_size, flush=True)
h.barrier(channel=0)
peer = h.get_buffer((rank + 1) % world, (16,), torch.float64)
peer.fill_(float(rank + 1))
h.barrier(channel=0)
torch.cuda.synchronize()
print(rank, "my buffer after peer write:", t[:4].tolist(), flush=True)
dist.destroy_process_group()
