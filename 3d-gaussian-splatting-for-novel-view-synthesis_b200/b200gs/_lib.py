"""ctypes binding of libb200gs.so (the C ABI declared in include/b200gs.h).

There is no CPU fallback: importing this module without the built library, or calling into it without
a CUDA device, raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_size_t, c_uint32,
                    c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200gs.so")

OK = 0
ERR_CAPACITY = -5
TILE = 16
CAM_KEEP_OUTSIDE_BAND = 1
CAM_OVERLAPPED = 2
CAM_ROUTED = 4
ROUTE_RECORDS_LATER = 1


class Gaussians(Structure):
    _fields_ = [("n", c_int32), ("pos", c_void_p), ("opacity_raw", c_void_p), ("scale_raw", c_void_p),
                ("q_raw", c_void_p), ("sigma", c_void_p), ("f_dc", c_void_p), ("f_rest", c_void_p),
                ("color", c_void_p)]


class Camera(Structure):
    _fields_ = [("c2w", c_void_p), ("H", c_int32), ("W", c_int32), ("fx", c_double), ("fy", c_double),
                ("cx", c_double), ("cy", c_double), ("near_plane", c_double), ("far_plane", c_double),
                ("pix_guard", c_double), ("min_conis", c_double), ("chi_square_clip", c_double),
                ("alpha_max", c_double), ("alpha_cutoff", c_double), ("tile", c_int32),
                ("tile_row_begin", c_int32), ("tile_row_end", c_int32), ("flags", c_int32)]


class Grads(Structure):
    _fields_ = [("pos", c_void_p), ("opacity_raw", c_void_p), ("scale_raw", c_void_p), ("q_raw", c_void_p),
                ("sigma", c_void_p), ("f_dc", c_void_p), ("f_rest", c_void_p), ("color", c_void_p)]


class Sizes(Structure):
    _fields_ = [("frame_bytes", c_size_t), ("isect_bytes", c_size_t)]


class AdamTensor(Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("numel", ctypes.c_int64), ("lr", c_double), ("step", c_int32), ("reserved", c_int32)]


MAX_PEERS = 16
PEER_MAX_TENSORS = 8
CLIP_MAX_TENSORS = 16          # B200GS_CLIP_MAX_TENSORS
PEER_CTRL_BYTES = 4096


class PeerLayout(Structure):
    _fields_ = [("offset", ctypes.c_int64 * PEER_MAX_TENSORS), ("per", ctypes.c_int64 * PEER_MAX_TENSORS),
                ("shard_offset", ctypes.c_int64 * PEER_MAX_TENSORS), ("flat_total", ctypes.c_int64),
                ("shard_total", ctypes.c_int64)]


class PeerGroup(Structure):
    _fields_ = [("world", c_int32), ("rank", c_int32), ("area", c_void_p * MAX_PEERS), ("multicast", c_void_p)]


class PeerTensor(Structure):
    _fields_ = [("grad", c_void_p), ("numel", ctypes.c_int64), ("lr", c_double), ("step", c_int32), ("clip", c_int32)]


class Route(Structure):
    _fields_ = [("world", c_int32), ("rank", c_int32), ("seg_capacity", c_uint32), ("band_row", c_int32 * (MAX_PEERS + 1)),
                ("band_ws", c_void_p * MAX_PEERS), ("band_ws_bytes", c_size_t), ("flags", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32)]


class FrameStats(Structure):
    _fields_ = [("n_isect", c_uint32), ("n_visible", c_uint32), ("overflow", c_uint32),
                ("n_in_frustum", c_uint32), ("n_super", c_uint32), ("n_sorted", c_uint32), ("n_candidates", c_uint32), ("reserved", c_uint32 * 9)]


# every symbol include/b200gs.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "b200gs_abi_version": (c_int, []),
    "b200gs_last_error": (c_char_p, []),
    "b200gs_workspace_sizes": (c_int, [c_int32, c_int32, c_int32, c_uint32, POINTER(Sizes)]),
    "b200gs_build_sigma": (c_int, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200gs_build_sigma_backward": (c_int, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200gs_evaluate_sh": (c_int, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200gs_evaluate_sh_backward": (c_int, [c_int32] + [c_void_p] * 9),
    "b200gs_render_project": (c_int, [POINTER(Gaussians), POINTER(Camera), c_void_p, c_size_t, c_void_p, c_void_p]),
    "b200gs_render_rasterize": (c_int, [POINTER(Camera), c_int32, c_void_p, c_size_t, c_void_p, c_size_t, c_uint32,
                                        c_void_p, c_void_p, c_void_p]),
    "b200gs_render_rasterize_ev": (c_int, [POINTER(Camera), c_int32, c_void_p, c_size_t, c_void_p, c_size_t, c_uint32,
                                           c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200gs_render_rasterize_split": (c_int, [POINTER(Camera), c_int32, c_void_p, c_size_t, c_void_p, c_size_t, c_uint32,
                                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200gs_render_backward": (c_int, [POINTER(Gaussians), POINTER(Camera), c_void_p, c_size_t, c_void_p, c_size_t,
                                       c_uint32, c_void_p, POINTER(Grads), c_void_p]),
    "b200gs_loss_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "b200gs_l1_ssim_forward": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_double, c_double, c_void_p,
                                       c_size_t, c_int32, c_void_p, c_void_p]),
    "b200gs_l1_ssim_backward": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_double, c_double, c_void_p,
                                        c_size_t, c_void_p, c_void_p, c_void_p]),
    "b200gs_image_to_u8": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200gs_adam_step": (c_int, [POINTER(AdamTensor), c_int32, c_double, c_double, c_double, c_void_p]),
    "b200gs_clip_workspace_bytes": (c_size_t, [ctypes.c_int64]),
    "b200gs_clip_grad_norm": (c_int, [c_void_p, ctypes.c_int64, c_double, c_void_p, c_size_t, c_void_p, c_void_p]),
    "b200gs_clip_workspace_bytes_multi": (c_size_t, [c_void_p, c_int32]),
    "b200gs_clip_grad_norm_multi": (c_int, [c_void_p, c_void_p, c_int32, c_double, c_void_p, c_size_t, c_void_p, c_void_p]),
    "b200gs_densify_workspace_bytes": (c_size_t, [c_int32]),
    "b200gs_densify_plan": (c_int, [c_int32, c_void_p, c_void_p, c_void_p, c_double, c_double, c_double, c_void_p, c_size_t,
                                    c_void_p, c_void_p]),
    "b200gs_densify_apply": (c_int, [c_int32, c_void_p, c_size_t, POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_void_p]),
    "b200gs_peer_layout_compute": (c_int, [POINTER(ctypes.c_int64), c_int32, c_int32, POINTER(PeerLayout)]),
    "b200gs_peer_area_bytes": (c_size_t, [POINTER(PeerLayout)]),
    "b200gs_peer_barrier": (c_int, [POINTER(PeerGroup), POINTER(c_uint32), c_void_p]),
    "b200gs_peer_adam_step": (c_int, [POINTER(PeerGroup), POINTER(PeerLayout), POINTER(PeerTensor), c_int32, c_void_p,
                                      c_void_p, c_double, c_double, c_double, c_double, c_int32, POINTER(c_uint32),
                                      c_void_p, c_void_p]),
    "b200gs_peer_allreduce": (c_int, [POINTER(PeerGroup), POINTER(PeerLayout), POINTER(PeerTensor), c_int32,
                                      POINTER(c_uint32), c_void_p]),
    "b200gs_route_project_slice": (c_int, [POINTER(Gaussians), POINTER(Camera), c_void_p, c_size_t, POINTER(Route), c_void_p]),
    "b200gs_route_records": (c_int, [c_int32, POINTER(Camera), c_void_p, c_size_t, POINTER(Route), c_void_p]),
    "b200gs_render_project_routed": (c_int, [POINTER(Camera), POINTER(Route), c_void_p, c_size_t, c_void_p, c_void_p]),
    "b200gs_render_host": (c_int, [POINTER(Gaussians), POINTER(Camera), c_void_p, c_void_p, POINTER(FrameStats)]),
    "b200gs_debug_export": (c_int, [c_int32, c_void_p, c_size_t, c_int32, c_int32] + [c_void_p] * 10),
    "b200gs_debug_export_lists": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_uint32, c_int32, c_int32, c_int32,
                                          c_void_p, c_void_p, c_uint32, c_void_p, c_void_p]),
    "b200gs_profile_enable": (c_int, [c_int]),
    "b200gs_profile_collect": (c_int, [c_void_p, c_void_p, c_int32]),
    "b200gs_profile_region_name": (c_char_p, [c_int32]),
    "b200gs_kernel_launch_count": (ctypes.c_ulonglong, []),
    "b200gs_exclusive_scan_u32": (c_int, [c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200gs_scan_scratch_bytes": (c_size_t, [c_uint32]),
    "b200gs_radix_sort_pairs": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_int, c_int, c_void_p,
                                        c_size_t, c_void_p]),
    "b200gs_sort_scratch_bytes": (c_size_t, [c_uint32]),
}

_lib = None

_env_data = getattr(os.environ, "_data", None)


def env(name: str, default=None):
    """os.environ.get without the encode / decode round trip of every lookup (the render path reads a handful of
    switches per frame, and they must stay live: tests flip them).  Falls back to os.environ.get."""
    if _env_data is not None:
        try:
            v = _env_data.get(name.encode())
            return default if v is None else v.decode()
        except Exception:  # noqa: BLE001 - not the CPython posix layout
            pass
    return os.environ.get(name, default)


class B200GSError(RuntimeError):
    pass


def load():
    """Load libb200gs.so and type every entry point.  Fails loudly when the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200GSError(
            f"{LIB_PATH} is missing: build it with `python 3d-gaussian-splatting-for-novel-view-synthesis_b200/"
            "build.py` (or __graft_entry__.build()).  b200gs has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.b200gs_abi_version() != 1:
        raise B200GSError("libb200gs.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != OK:
        msg = load().b200gs_last_error()
        raise B200GSError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
