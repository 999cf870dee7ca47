"""ctypes front-end of the host-side check harness (tests only).

Builds tests/hostcheck/hostcheck.cpp - which includes the product's csrc/gs_math.cuh - with g++ and
exposes the per-Gaussian forward/backward arithmetic of the CUDA kernels to CPU tests.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "hostcheck.cpp")
HDR = os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200", "csrc", "gs_math.cuh")
LIB = os.path.join(HERE, "libhostcheck.so")

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    stale = (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR))
    if stale:
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++", SRC,
                        "-o", LIB], check=True)
    _lib = ctypes.CDLL(LIB)
    return _lib


def _p(a, ct=ctypes.c_float):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def project(pos, opacity_raw, c2w, cam, scale_raw=None, q_raw=None, sigma=None, f_dc=None, f_rest=None, color=None):
    """cam = (H, W, fx, fy, cx, cy).  Returns a dict of per-Gaussian arrays."""
    lib = load()
    n = pos.shape[0]
    pos, opacity_raw, c2w = _f32(pos), _f32(opacity_raw), _f32(c2w)
    scale_raw, q_raw, sigma, f_dc, f_rest, color = map(_f32, (scale_raw, q_raw, sigma, f_dc, f_rest, color))
    camd = np.asarray(cam, dtype=np.float64)
    out = dict(vis=np.zeros(n, np.int32), u=np.zeros(n, np.float32), v=np.zeros(n, np.float32),
               z=np.zeros(n, np.float32), op=np.zeros(n, np.float32), conic=np.zeros((n, 3), np.float32),
               rgb=np.zeros((n, 3), np.float32), radius=np.zeros(n, np.int32), rect=np.zeros((n, 4), np.int32),
               tiles=np.zeros(n, np.int32), lam2=np.zeros(n, np.float32), clamped=np.zeros(n, np.int32))
    i32 = ctypes.c_int32
    lib.hc_project(ctypes.c_int(n), _p(pos), _p(scale_raw), _p(q_raw), _p(sigma), _p(opacity_raw), _p(f_dc),
                   _p(f_rest), _p(color), _p(c2w), _p(camd, ctypes.c_double), _p(out["vis"], i32), _p(out["u"]),
                   _p(out["v"]), _p(out["z"]), _p(out["op"]), _p(out["conic"]), _p(out["rgb"]),
                   _p(out["radius"], i32), _p(out["rect"], i32), _p(out["tiles"], i32), _p(out["lam2"]),
                   _p(out["clamped"], i32))
    return out


def backward(pos, scale_raw, q_raw, opacity_raw, f_dc, f_rest, c2w, cam, splat_grads):
    lib = load()
    n = pos.shape[0]
    arrs = [_f32(a) for a in (pos, scale_raw, q_raw, opacity_raw, f_dc, f_rest, c2w)]
    camd = np.asarray(cam, dtype=np.float64)
    sg = _f32(splat_grads)
    out = dict(pos=np.zeros((n, 3), np.float32), scale_raw=np.zeros((n, 3), np.float32),
               q_raw=np.zeros((n, 4), np.float32), opacity_raw=np.zeros(n, np.float32),
               f_dc=np.zeros((n, 3), np.float32), f_rest=np.zeros((n, 45), np.float32))
    lib.hc_backward(ctypes.c_int(n), *[_p(a) for a in arrs], _p(camd, ctypes.c_double), _p(sg), _p(out["pos"]),
                    _p(out["scale_raw"]), _p(out["q_raw"]), _p(out["opacity_raw"]), _p(out["f_dc"]),
                    _p(out["f_rest"]))
    return out
