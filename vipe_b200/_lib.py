"""ctypes binding of libvipe_ba.so (include/vipe_ba.h).  There is no CPU fallback: if the library is
missing this module raises at import of the first symbol, and every op needs CUDA tensors."""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_SO = Path(__file__).resolve().parent / "lib" / "libvipe_ba.so"
_lib = None

ABI_VERSION = 3


class Tensors(C.Structure):
    """struct vipe_ba_tensors"""

    _fields_ = [
        ("poses", C.c_void_p),
        ("disps", C.c_void_p),
        ("intrinsics", C.c_void_p),
        ("disps_sens", C.c_void_p),
        ("targets", C.c_void_p),
        ("weights", C.c_void_p),
        ("eta", C.c_void_p),
        ("dx_out", C.c_void_p),
        ("dz_out", C.c_void_p),
    ]


class Options(C.Structure):
    """struct vipe_ba_options"""

    _fields_ = [
        ("min_depth", C.c_float),
        ("depth_strict", C.c_int),
        ("alpha", C.c_float),
        ("sensor_mode", C.c_int),
        ("eta_scale", C.c_float),
        ("eta_bias", C.c_float),
        ("dz_max", C.c_float),
        ("renorm_quat", C.c_int),
        ("damp_on_pose_hessian", C.c_int),
        ("backsub_all_poses", C.c_int),
        ("frame_flags", C.c_void_p),
        ("optimize_focal", C.c_int),
        ("focal_jscale", C.c_float),
        ("focal_lm", C.c_float),
        ("focal_ep", C.c_float),
    ]


# name -> (restype, argtypes); this table mirrors include/vipe_ba.h one to one (tests check it)
SIGNATURES = {
    "vipe_ba_abi_version": (C.c_int, []),
    "vipe_ba_last_error": (C.c_char_p, []),
    "vipe_ba_plan_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vipe_ba_plan_create_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "vipe_ba_plan_num_free_poses": (C.c_int64, [C.c_void_p]),
    "vipe_ba_plan_destroy": (None, [C.c_void_p]),
    "vipe_ba_plan_num_kx": (C.c_int64, [C.c_void_p]),
    "vipe_ba_plan_copy_kx": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vipe_ba_plan_copy_kk_exp": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vipe_ba_plan_copy_csr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vipe_ba_plan_owned_range": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vipe_ba_plan_num_schur_triples": (C.c_int64, [C.c_void_p]),
    "vipe_ba_plan_copy_sys_order": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vipe_ba_plan_num_owned_edges": (C.c_int64, [C.c_void_p]),
    "vipe_ba_plan_copy_owned_edges": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vipe_ba_set_owned_rows": (C.c_int, [C.c_void_p, C.c_int]),
    "vipe_ba_set_solve_buffer": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vipe_ba_dist_aux_bytes": (C.c_int64, [C.c_void_p]),
    "vipe_ba_set_dist_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "vipe_ba_peer_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vipe_ba_plan_max_degree": (C.c_int, [C.c_void_p]),
    "vipe_ba_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "vipe_ba_plan_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vipe_ba_run": (C.c_int, [C.c_void_p, C.POINTER(Tensors), C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int,
                              C.c_void_p]),
    "vipe_ba_linearize": (C.c_int, [C.c_void_p, C.POINTER(Tensors), C.c_void_p, C.c_int, C.c_void_p]),
    "vipe_ba_solve_update": (C.c_int, [C.c_void_p, C.POINTER(Tensors), C.c_void_p, C.c_float, C.c_float, C.c_int,
                                       C.c_void_p]),
    "vipe_ba_system_buffer": (C.c_void_p, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vipe_ba_debug_q": (C.c_void_p, [C.c_void_p, C.c_void_p]),
    "vipe_ba_debug_qw": (C.c_void_p, [C.c_void_p, C.c_void_p]),
    "vipe_ba_launch_count": (C.c_int64, [C.c_void_p]),
    "vipe_ba_set_peer_system": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vipe_ba_set_graphs": (C.c_int, [C.c_void_p, C.c_int]),
    "vipe_ba_options_default": (None, [C.POINTER(Options)]),
    "vipe_ba_set_options": (C.c_int, [C.c_void_p, C.POINTER(Options)]),
    "vipe_projmap": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vipe_frame_distance": (C.c_int, [C.c_void_p] * 8 + [C.c_int64, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "vipe_depth_filter": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vipe_iproj": (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vipe_ba_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "vipe_ba_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_float * 4), C.POINTER(C.c_int)]),
}


def so_path() -> Path:
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not _SO.is_file():
            raise ImportError(
                f"{_SO} is missing: build it with `python -m vipe_b200.build` (or __graft_entry__.build()). "
                "vipe_b200 has no CPU or PyTorch fallback for the bundle-adjustment path."
            )
        L = C.CDLL(str(_SO))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.vipe_ba_abi_version() != ABI_VERSION:
            raise ImportError(f"libvipe_ba.so ABI {L.vipe_ba_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().vipe_ba_last_error()
        raise RuntimeError(f"{what}: {msg.decode() if msg else 'unknown error'}")
