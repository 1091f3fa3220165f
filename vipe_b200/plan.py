"""Python handle of a `vipe_ba_plan` (index bookkeeping of one BA graph) plus its device workspace."""

from __future__ import annotations

import ctypes as C
import hashlib
from collections import OrderedDict

import torch

from . import _lib


class BAPlan:
    """Index bookkeeping of one BA problem (replaces geom_kernels.cu:1301-1308, :946-981, :1209-1240).

    Host-only: building a plan needs no GPU.  `workspace(device)` allocates the device workspace from the
    torch caching allocator and uploads the index tables into its head."""

    def __init__(self, ii: torch.Tensor, jj: torch.Tensor, n_frames: int, ht: int, wd: int, t0: int, t1: int,
                 rank: int = 0, world: int = 1, batch=None):
        """`batch = (frame_ptr, t0s, t1s)` (int64 host tensors) builds a batched plan of independent small problems
        (vipe_ba_plan_create_batch); t0/t1 are then ignored."""
        L = _lib.lib()
        ii_h = ii.detach().to("cpu", torch.int64).contiguous()
        jj_h = jj.detach().to("cpu", torch.int64).contiguous()
        if ii_h.dim() != 1 or ii_h.shape != jj_h.shape:
            raise RuntimeError("ii and jj must be 1-D tensors of the same length")
        self.E = int(ii_h.numel())
        self.N, self.ht, self.wd, self.t0, self.t1 = int(n_frames), int(ht), int(wd), int(t0), int(t1)
        self.P = self.t1 - self.t0
        self.rank, self.world = rank, world
        h = C.c_void_p()
        if batch is None:
            _lib.check(L.vipe_ba_plan_create(ii_h.data_ptr(), jj_h.data_ptr(), self.E, self.N, self.ht, self.wd, self.t0,
                                             self.t1, rank, world, C.byref(h)), "vipe_ba_plan_create")
        else:
            fp, t0s, t1s = [torch.as_tensor(x, dtype=torch.int64).contiguous() for x in batch]
            _lib.check(L.vipe_ba_plan_create_batch(ii_h.data_ptr(), jj_h.data_ptr(), self.E, self.N, self.ht, self.wd,
                                                   int(t0s.numel()), fp.data_ptr(), t0s.data_ptr(), t1s.data_ptr(), C.byref(h)),
                       "vipe_ba_plan_create_batch")
        self._h = h
        self.K = int(L.vipe_ba_plan_num_kx(h))
        self.P = int(L.vipe_ba_plan_num_free_poses(h))
        self._ws: dict = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().vipe_ba_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    @property
    def handle(self):
        return self._h

    # ---- bookkeeping views (host tensors) -------------------------------------------------------
    @property
    def kx(self) -> torch.Tensor:
        out = torch.empty(self.K, dtype=torch.int64)
        _lib.check(_lib.lib().vipe_ba_plan_copy_kx(self._h, out.data_ptr()), "copy_kx")
        return out

    @property
    def kk_exp(self) -> torch.Tensor:
        out = torch.empty(self.P + self.E, dtype=torch.int64)
        _lib.check(_lib.lib().vipe_ba_plan_copy_kk_exp(self._h, out.data_ptr()), "copy_kk_exp")
        return out

    def csr(self):
        ptrs = torch.empty(self.K + 1, dtype=torch.int64)
        idxs = torch.empty(max(self.E, 1), dtype=torch.int64)
        _lib.check(_lib.lib().vipe_ba_plan_copy_csr(self._h, ptrs.data_ptr(), idxs.data_ptr()), "copy_csr")
        return ptrs, idxs[: self.E]

    def owned_range(self, rank: int | None = None):
        lo, hi = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().vipe_ba_plan_owned_range(self._h, self.rank if rank is None else rank, C.byref(lo),
                                                       C.byref(hi)), "owned_range")
        return int(lo.value), int(hi.value)

    @property
    def num_schur_triples(self) -> int:
        return int(_lib.lib().vipe_ba_plan_num_schur_triples(self._h))

    @property
    def max_degree(self) -> int:
        return int(_lib.lib().vipe_ba_plan_max_degree(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(_lib.lib().vipe_ba_workspace_bytes(self._h))

    def set_options(self, **kw):
        """vipe_ba_set_options: override the reference-CUDA defaults (see include/vipe_ba.h).  `frame_flags` is a CUDA
        uint8 tensor [K] that must stay alive as long as the plan uses it."""
        opt = _lib.Options()
        _lib.lib().vipe_ba_options_default(C.byref(opt))
        self._flags_keepalive = kw.get("frame_flags")
        for k, v in kw.items():
            if k == "frame_flags":
                opt.frame_flags = v.data_ptr() if v is not None else None
            else:
                if not hasattr(opt, k):
                    raise TypeError(f"unknown BA option {k}")
                setattr(opt, k, v)
        _lib.check(_lib.lib().vipe_ba_set_options(self._h, C.byref(opt)), "vipe_ba_set_options")

    @property
    def launch_count(self) -> int:
        return int(_lib.lib().vipe_ba_launch_count(self._h))

    # ---- device side ------------------------------------------------------------------------------
    def workspace(self, device: torch.device) -> torch.Tensor:
        key = (device.type, device.index)
        ws = self._ws.get(key)
        if ws is None:
            if device.type != "cuda":
                raise RuntimeError("vipe_b200 runs on CUDA devices only (no CPU fallback)")
            ws = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=device)
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(_lib.lib().vipe_ba_plan_upload(self._h, ws.data_ptr(), stream), "plan_upload")
            self._ws[key] = ws
        return ws

    def owned_edges(self) -> torch.Tensor:
        """Edge ids of this rank's shard in the order owner-only `targets` / `weights` rows must have (CSR order over the owned
        source frames), int64 host tensor."""
        n = int(_lib.lib().vipe_ba_plan_num_owned_edges(self._h))
        out = torch.empty(max(n, 0), dtype=torch.int64)
        if n > 0:
            _lib.check(_lib.lib().vipe_ba_plan_copy_owned_edges(self._h, out.data_ptr()), "copy_owned_edges")
        return out

    @property
    def sys_order(self) -> torch.Tensor:
        """Position of each free pose in the reduced camera system (a fill-reducing elimination order), [P] int64."""
        out = torch.empty(self.P, dtype=torch.int64)
        _lib.check(_lib.lib().vipe_ba_plan_copy_sys_order(self._h, out.data_ptr()), "copy_sys_order")
        return out

    def system_index(self) -> torch.Tensor:
        """[6P] int64: index in the system buffer of unknown (pose p, component r) in pose order."""
        o = self.sys_order
        return (6 * o[:, None] + torch.arange(6)[None, :]).reshape(-1)

    def system_view(self, ws: torch.Tensor) -> torch.Tensor:
        """fp64 view [npad*npad + 2*npad] of the reduced camera system ([H ; b ; diag of the pose Hessian]) inside the workspace (all-reduce target)."""
        n, cnt = C.c_int64(), C.c_int64()
        ptr = _lib.lib().vipe_ba_system_buffer(self._h, ws.data_ptr(), C.byref(n), C.byref(cnt))
        off = ptr - ws.data_ptr()
        return ws[off: off + 8 * cnt.value].view(torch.float64), int(n.value)

    def debug_q(self, ws: torch.Tensor):
        HW = self.ht * self.wd
        out = []
        for fn in (_lib.lib().vipe_ba_debug_q, _lib.lib().vipe_ba_debug_qw):
            off = fn(self._h, ws.data_ptr()) - ws.data_ptr()
            out.append(ws[off: off + 4 * self.K * HW].view(torch.float32).view(self.K, HW))
        return out


_CACHE: "OrderedDict[tuple, BAPlan]" = OrderedDict()
_CACHE_MAX = 8


def cached_plan(ii_h: torch.Tensor, jj_h: torch.Tensor, n_frames, ht, wd, t0, t1, rank=0, world=1, tag="") -> BAPlan:
    """Plans depend only on the graph, and SLAM calls BA many times on the same graph (every GRU step of
    FactorGraph.update / update_batch, vipe/slam/components/factor_graph.py:296,378), so keep a few."""
    dig = hashlib.blake2b(ii_h.numpy().tobytes() + b"|" + jj_h.numpy().tobytes(), digest_size=16).digest()
    # `tag` separates plans that carry different semantic options (vipe_ba_set_options is per plan)
    key = (dig, int(n_frames), int(ht), int(wd), int(t0), int(t1), rank, world, tag)
    p = _CACHE.get(key)
    if p is None:
        p = BAPlan(ii_h, jj_h, n_frames, ht, wd, t0, t1, rank, world)
        _CACHE[key] = p
        while len(_CACHE) > _CACHE_MAX:
            _CACHE.popitem(last=False)
    else:
        _CACHE.move_to_end(key)
    return p
