"""Callers whose BA inputs live in host memory: uploads on a copy stream, overlapped with the previous solve.

`slam_ext.ba` takes device tensors, like the reference operator (csrc/slam_ext/slam.cpp:24-27).  A caller that keeps
`targets/weights` in pinned host memory pays a host-to-device copy per call -- at backend size (C3: 158 MB) that copy
is 40 % of the solve time when it sits on the same stream.  `HostFeed` keeps two device argument sets and copies call
s+1's arguments on its own stream while call s runs; nothing here touches the BA itself.
"""

from __future__ import annotations

import torch

from .ext import slam_ext

__all__ = ["HostFeed"]


class HostFeed:
    """Double-buffered upload of `slam_ext.ba` argument lists.

        feed = HostFeed("cuda:0")
        feed.prefetch(args_0)                       # host tensors (pinned for a truly asynchronous copy) + scalars
        for s in range(n):
            if s + 1 < n: feed.prefetch(args_{s+1})  # copy stream: runs under call s
            (dx, dz), dev_args = feed.run(out_poses=..., out_disps=...)   # current stream

    `run` makes the current stream wait for the upload, calls `fn` (default `slam_ext.ba`) on the device copies and, if
    host output tensors are given, reads the updated poses / disparities back on a third stream (`synchronize()` waits
    for them).
    A slot is reused only after the call that consumed it -- including its read-back -- has finished on the device."""

    def __init__(self, device, slots: int = 2):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("HostFeed needs a CUDA device (vipe_b200 has no CPU fallback)")
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.down_stream = torch.cuda.Stream(self.dev)  # results go back on their own stream: the next solve does not wait for them
        self._slots = [None] * slots
        self._free = [None] * slots
        self._ready = []
        self._next = 0
        self.h2d_bytes = 0

    def _device_slot(self, b, host_args):
        slot = self._slots[b]
        ok = slot is not None and len(slot) == len(host_args) and all(
            (not torch.is_tensor(h)) or (torch.is_tensor(d) and d.shape == h.shape and d.dtype == h.dtype)
            for d, h in zip(slot, host_args))
        if not ok:
            slot = [torch.empty(h.shape, dtype=h.dtype, device=self.dev) if torch.is_tensor(h) else h for h in host_args]
            self._slots[b] = slot
        return slot

    def prefetch(self, host_args):
        b = self._next
        self._next = (b + 1) % len(self._slots)
        if any(r[0] == b for r in self._ready):
            raise RuntimeError("HostFeed: every slot holds an upload that has not been run yet")
        with torch.cuda.stream(self.copy_stream):
            if self._free[b] is not None:
                self.copy_stream.wait_event(self._free[b])
            slot = self._device_slot(b, host_args)
            n = 0
            for i, h in enumerate(host_args):
                if torch.is_tensor(h):
                    slot[i].copy_(h, non_blocking=True)
                    n += h.numel() * h.element_size()
                else:
                    slot[i] = h
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.h2d_bytes = n
        # the bookkeeping plan comes from the HOST copy of the edge list: hashing 16 bytes per edge on the CPU instead of
        # reading ii/jj back from the device, which would synchronise the stream once per call
        plan = None
        if len(host_args) == 15 and torch.is_tensor(host_args[7]) and host_args[7].device.type == "cpu":
            from .plan import cached_plan

            N, ht, wd = host_args[1].shape
            plan = cached_plan(host_args[7].contiguous(), host_args[8].contiguous(), N, ht, wd, int(host_args[9]),
                               int(host_args[10]))
        self._ready.append((b, ev, list(slot), plan))

    def run(self, fn=None, out_poses=None, out_disps=None):
        if not self._ready:
            raise RuntimeError("HostFeed.run without a prefetch")
        b, ev, a, plan = self._ready.pop(0)
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(ev)
        res = fn(*a) if fn is not None else slam_ext.ba(*a, plan=plan)
        done = torch.cuda.Event()
        done.record(cur)
        if out_poses is not None or out_disps is not None:
            with torch.cuda.stream(self.down_stream):
                self.down_stream.wait_event(done)
                if out_poses is not None:
                    out_poses.copy_(a[0], non_blocking=True)
                if out_disps is not None:
                    out_disps.copy_(a[1], non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.down_stream)
        self._free[b] = done  # the slot is reused only after its results have left
        self.last_done = done
        return res, a

    def synchronize(self):
        """Block until the results of every `run` so far are in their host tensors."""
        self.down_stream.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()
