"""torch-side plumbing for libb200gs: device buffers, streams and the autograd bridge.

PyTorch is used for memory (caching allocator), streams and autograd bookkeeping only; all
arithmetic happens in the hand-written CUDA kernels behind the C ABI (include/b200gs.h).
"""
from __future__ import annotations

import ctypes
import os
import weakref
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import Camera, FrameStats, Gaussians, Grads, Sizes


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.B200GSError(
            f"b200gs: `{name}` lives on {t.device}; this rasterizer runs on CUDA (sm_100a) only and has no CPU "
            "fallback.")


def _require_f32(name: str, t: Optional[torch.Tensor]):
    """The reference is dtype-generic (its image follows pos.dtype, render.py:318); this path computes in fp32 only
    and refuses anything else rather than rounding it silently."""
    if t is None or t.dtype is torch.float32:          # the common case costs one attribute read
        return
    if t.is_floating_point():
        raise TypeError(f"b200gs: `{name}` is {t.dtype}; the CUDA path computes in float32 only (pass .float() tensors)")


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@dataclass
class RenderConfig:
    """The non-tensor arguments of render.py:62-64 plus the tile-row band of this rank."""
    H: int
    W: int
    fx: float
    fy: float
    cx: float
    cy: float
    near: float = 0.01
    far: float = 100.0
    pix_guard: float = 32
    T: int = 16
    min_conis: float = 1e-6
    chi_square_clip: float = 6.25
    alpha_max: float = 0.99
    alpha_cutoff: float = 1 / 128.
    tile_row_begin: int = 0
    tile_row_end: int = 0
    keep_outside_band: bool = False     # a band writes its own pixels only (the image belongs to a tile-row sharded frame)
    out: Optional[torch.Tensor] = None  # caller-owned [H,W,3] fp32 image to render into (returned as the result)
    out_ptr: int = 0                    # raw device address the kernels write instead of out.data_ptr(): the same frame
                                        # buffer on another GPU, peer-mapped (b200gs.dist.TileRowRenderer)

    def to_c(self, c2w: torch.Tensor) -> Camera:
        if int(self.T) != _lib.TILE:
            raise NotImplementedError(f"b200gs renders with T=16 tiles only (got T={self.T}); the reference's "
                                      "callers never change T")
        return Camera(c2w=c2w.data_ptr(), H=int(self.H), W=int(self.W), fx=float(self.fx), fy=float(self.fy),
                      cx=float(self.cx), cy=float(self.cy), near_plane=float(self.near), far_plane=float(self.far),
                      pix_guard=float(self.pix_guard), min_conis=float(self.min_conis),
                      chi_square_clip=float(self.chi_square_clip), alpha_max=float(self.alpha_max),
                      alpha_cutoff=float(self.alpha_cutoff), tile=int(self.T),
                      tile_row_begin=int(self.tile_row_begin), tile_row_end=int(self.tile_row_end),
                      flags=_lib.CAM_KEEP_OUTSIDE_BAND if self.keep_outside_band else 0)


_pinned_stats = {}


def _stats_buffer(device, stream_handle: int):
    """(pinned int32[16] tensor, numpy view of it, data pointer) for this device/stream."""
    key = (device.index, stream_handle)
    ent = _pinned_stats.get(key)
    if ent is None:
        buf = torch.zeros(16, dtype=torch.int32).pin_memory()
        ent = (buf, buf.numpy(), ctypes.c_void_p(buf.data_ptr()))
        _pinned_stats[key] = ent
    return ent


_stats_slots = {}
_STATS_RING = 8          # frames whose statistics may be outstanding per (device, stream): RenderPipeline keeps <= 4


def _stats_slot(device, stream):
    """Next (numpy view, data pointer, torch event, raw cudaEvent_t) of this stream's ring of pinned statistics
    buffers: a frame's counters stay readable until _STATS_RING - 1 later frames have been launched."""
    key = (device.index, stream.cuda_stream)
    ring = _stats_slots.get(key)
    if ring is None:
        bufs = torch.zeros(_STATS_RING, 16, dtype=torch.int32).pin_memory()
        slots = []
        for i in range(_STATS_RING):
            ev = torch.cuda.Event()
            ev.record(stream)                   # torch creates the CUDA event lazily, on first record
            slots.append((bufs[i].numpy(), ctypes.c_void_p(bufs[i].data_ptr()), ev, ctypes.c_void_p(ev.cuda_event)))
        ring = [bufs, slots, 0]
        _stats_slots[key] = ring
    ring[2] = (ring[2] + 1) % _STATS_RING
    return ring[1][ring[2]]


_stats_events = {}


def _stats_event(device, stream):
    """(torch event, raw cudaEvent_t) the library records once a frame's counters are final."""
    key = (device.index, stream.cuda_stream)
    ent = _stats_events.get(key)
    if ent is None:
        ev = torch.cuda.Event()
        ev.record(stream)                       # torch creates the CUDA event lazily, on first record
        ent = (ev, ctypes.c_void_p(ev.cuda_event))
        _stats_events[key] = ent
    return ent


_sizes_cache = {}


def _sizes(lib, n: int, H: int, W: int, capacity: int):
    key = (n, H, W, capacity)
    got = _sizes_cache.get(key)
    if got is None:
        sz = Sizes()
        _lib.check(lib.b200gs_workspace_sizes(n, H, W, capacity, ctypes.byref(sz)), "workspace_sizes")
        got = (int(sz.frame_bytes), int(sz.isect_bytes))
        if len(_sizes_cache) > 4096:
            _sizes_cache.clear()
        _sizes_cache[key] = got
    return got


# Intersection-capacity policy.  "speculative" (default): size the lists from a per-device high-water mark (the
# first frame on a device is sized exactly), queue the whole frame without a mid-frame sync and wait only for the
# event the library records once the frame's counters are final; a frame that did not fit is rasterized again with
# exact buffers.  "sync": read I back between project and rasterize (one stream sync per frame, exact buffers).
# The mark is kept per (device, N, H, W): unrelated scenes or resolutions in one process do not inflate each other's lists.
_high_water = {}
_blend_stream = {}       # device index -> torch stream the blend kernels go to (set by api.RenderPipeline)


class Frame:
    """One rendered view: owns the workspaces the backward needs."""

    def __init__(self, g: Gaussians, keep, cam_cfg: RenderConfig, c2w: torch.Tensor, device):
        self.g, self.keep, self.cfg, self.c2w, self.device = g, keep, cam_cfg, c2w, device
        self.out, self.out_ptr = cam_cfg.out, cam_cfg.out_ptr
        self.cam = cam_cfg.to_c(c2w)
        self.frame_ws = None
        self.isect_ws = None
        self.capacity = 0
        self.n_isect = 0
        self.n_visible = 0
        self.n_in_frustum = 0
        self.n_super = 0
        self.overflow = False
        self._unchecked = None

    # -- forward ----------------------------------------------------------------------------------------
    def render(self, mode: Optional[str] = None) -> torch.Tensor:
        image = self.launch(mode)
        self.finish()
        return image

    def launch(self, mode: Optional[str] = None, buffers=None) -> torch.Tensor:
        """Queues the frame.  In speculative mode nothing is waited for: `finish()` must be called (before
        _STATS_RING - 1 further frames are launched on this stream) to learn whether the frame fitted its buffers
        and to redo it if it did not; the image is only trustworthy after that."""
        lib = _lib.load()
        dev = self.device
        n = int(self.g.n)
        H, W = int(self.cfg.H), int(self.cfg.W)
        mode = mode or _lib.env("B200GS_CAPACITY_MODE", "speculative")
        stream = torch.cuda.current_stream(dev)
        st = ctypes.c_void_p(stream.cuda_stream)
        frame_bytes, _ = _sizes(lib, n, H, W, 0)
        # `buffers` = [frame_ws, isect_ws] owned by the caller (RenderPipeline keeps a ring of them instead of going
        # through the allocator every frame); replaced in place when too small
        self._buffers = buffers
        if buffers is not None and buffers[0] is not None and buffers[0].numel() >= frame_bytes:
            self.frame_ws = buffers[0]
        else:
            self.frame_ws = torch.empty(frame_bytes, dtype=torch.uint8, device=dev)
            if buffers is not None:
                buffers[0] = self.frame_ws
        stats_np, stats_ptr, ev, ev_ptr = _stats_slot(dev, stream)
        image = self.out if self.out is not None else torch.empty((H, W, 3), dtype=torch.float32, device=dev)
        if self.out is not None and (tuple(image.shape) != (H, W, 3) or image.dtype != torch.float32 or
                                     not image.is_contiguous() or image.device != dev):
            raise ValueError("out must be a contiguous [H,W,3] float32 tensor on the device of the Gaussians")
        self._hw_key = (dev.index, n, H, W)
        if _blend_stream.get(dev.index) is not None:      # frame pipeline: the binning runs beside the previous blend
            self.cam.flags |= _lib.CAM_OVERLAPPED
        spec_cap = _high_water.get(self._hw_key, 0) if mode == "speculative" else 0
        self._project(lib, frame_bytes, None if spec_cap else stats_ptr, st)
        self._unchecked = None
        if spec_cap:
            # the whole frame is queued; the host will only wait until the counters are final (end of the binning
            # scan), not for the sort / split / blend kernels behind it
            self._rasterize(lib, n, H, W, spec_cap, image, stats_ptr, st, ev_ptr)
            self._unchecked = (stream, st, stats_np, stats_ptr, ev, spec_cap, image)
        else:
            stream.synchronize()
            self._read_stats(stats_np)
            self._rasterize(lib, n, H, W, self._grow(self.n_isect) if mode == "speculative" else max(self.n_isect, 1),
                            image, stats_ptr, st)
            self._stats_pending = (stream, stats_np)     # n_super is only known after rasterize
        return image

    def _project(self, lib, frame_bytes, stats_ptr, st):
        _lib.check(lib.b200gs_render_project(ctypes.byref(self.g), ctypes.byref(self.cam), _ptr(self.frame_ws),
                                             frame_bytes, stats_ptr, st), "render_project")

    def finish(self):
        """Speculative frames: wait for the frame's counters (NOT for the frame) and rasterize again with exact
        buffers if it did not fit (returns True in that case)."""
        if self._unchecked is None:
            return False
        stream, st, stats_np, stats_ptr, ev, spec_cap, image = self._unchecked
        self._unchecked = None
        ev.synchronize()
        self._read_stats(stats_np)
        if self.n_isect > spec_cap or self.overflow:         # did not fit: redo with exact buffers
            lib = _lib.load()
            with torch.cuda.stream(stream):
                self._rasterize(lib, int(self.g.n), int(self.cfg.H), int(self.cfg.W), self._grow(self.n_isect), image,
                                stats_ptr, st)
            stream.synchronize()
            self._read_stats(stats_np)
            return True
        return False

    def refresh_stats(self):
        """Counters written by the rasterize phase (n_super); synchronises the stream."""
        pend = getattr(self, "_stats_pending", None)
        if pend is not None:
            pend[0].synchronize()
            self._read_stats(pend[1])
            self._stats_pending = None

    def _grow(self, need: int) -> int:
        key = getattr(self, "_hw_key", None) or (self.device.index, int(self.g.n), int(self.cfg.H), int(self.cfg.W))
        cap = max(int(need * 1.25) + 1024, _high_water.get(key, 0))
        if len(_high_water) > 256:
            _high_water.clear()
        _high_water[key] = cap
        return cap

    def _read_stats(self, arr):
        # b200gs_frame_stats: n_isect, n_visible, overflow, n_in_frustum, n_super (uint32 each)
        self.n_isect, self.n_visible = int(arr[0]) & 0xFFFFFFFF, int(arr[1]) & 0xFFFFFFFF
        self.n_in_frustum, self.n_super = int(arr[3]) & 0xFFFFFFFF, int(arr[4]) & 0xFFFFFFFF
        self.overflow = int(arr[2]) != 0

    def _rasterize(self, lib, n, H, W, capacity, image, stats_ptr, st, stats_event=None):
        frame_bytes, isect_bytes = _sizes(lib, n, H, W, capacity)
        self.capacity = capacity
        buffers = getattr(self, "_buffers", None)
        if buffers is not None and buffers[1] is not None and buffers[1].numel() >= isect_bytes:
            self.isect_ws = buffers[1]
        else:
            self.isect_ws = torch.empty(isect_bytes, dtype=torch.uint8, device=self.device)
            if buffers is not None:
                buffers[1] = self.isect_ws
        blend = _blend_stream.get(self.device.index)
        if blend is None:
            _lib.check(lib.b200gs_render_rasterize_ev(ctypes.byref(self.cam), n, _ptr(self.frame_ws), frame_bytes,
                                                      _ptr(self.isect_ws), isect_bytes, capacity, self._image_ptr(image), stats_ptr,
                                                      stats_event, st),
                       "render_rasterize")
        else:
            # frame pipelining (RenderPipeline): the blend runs on its own stream; the buffers were allocated in the
            # binning stream's pool, so the allocator must know about their second user
            for t in (self.frame_ws, self.isect_ws, image):
                t.record_stream(blend)
            _lib.check(lib.b200gs_render_rasterize_split(ctypes.byref(self.cam), n, _ptr(self.frame_ws), frame_bytes,
                                                         _ptr(self.isect_ws), isect_bytes, capacity, self._image_ptr(image),
                                                         stats_ptr, stats_event, st, ctypes.c_void_p(blend.cuda_stream)),
                       "render_rasterize_split")

    def _image_ptr(self, image):
        return ctypes.c_void_p(int(self.out_ptr)) if self.out_ptr else _ptr(image)

    # -- backward ---------------------------------------------------------------------------------------
    def backward(self, grad_image: torch.Tensor, grads: Grads):
        lib = _lib.load()
        _lib.check(lib.b200gs_render_backward(ctypes.byref(self.g), ctypes.byref(self.cam), _ptr(self.frame_ws),
                                              self.frame_ws.numel(), _ptr(self.isect_ws), self.isect_ws.numel(),
                                              self.capacity, _ptr(grad_image), ctypes.byref(grads),
                                              _stream(self.device)), "render_backward")

    # -- introspection (parity tests) -------------------------------------------------------------------
    def export(self):
        lib = _lib.load()
        self.refresh_stats()
        n, dev = int(self.g.n), self.device
        H, W = int(self.cfg.H), int(self.cfg.W)
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        i = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
        out = dict(xy=f(n, 2), depth=f(n), conic=f(n, 3), opacity=f(n), color=f(n, 3), radius=i(n), rect=i(n, 4),
                   tiles_touched=i(n), depth_order=i(n))
        _lib.check(lib.b200gs_debug_export(n, _ptr(self.frame_ws), self.frame_ws.numel(), H, W, _ptr(out["xy"]),
                                           _ptr(out["depth"]), _ptr(out["conic"]), _ptr(out["opacity"]),
                                           _ptr(out["color"]), _ptr(out["radius"]), _ptr(out["rect"]),
                                           _ptr(out["tiles_touched"]), _ptr(out["depth_order"]), _stream(dev)),
                   "debug_export")
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        cnt = min(self.n_isect, self.capacity)
        out.update(list_tile=i(max(cnt, 1)), list_id=i(max(cnt, 1)), ranges=i(tiles, 2))
        _lib.check(lib.b200gs_debug_export_lists(_ptr(self.frame_ws), self.frame_ws.numel(), _ptr(self.isect_ws),
                                                 self.isect_ws.numel(), self.capacity, n, H, W, _ptr(out["list_tile"]),
                                                 _ptr(out["list_id"]), cnt, _ptr(out["ranges"]), _stream(dev)),
                   "debug_export_lists")
        out["list_tile"], out["list_id"] = out["list_tile"][:cnt], out["list_id"][:cnt]
        torch.cuda.current_stream(dev).synchronize()
        out = {k: v.cpu() for k, v in out.items()}
        # behind the survivors the sorted order lists the culled Gaussians: report the survivors, in depth order
        order = out["depth_order"].long()
        order = order[(order >= 0) & (order < n)]
        out["depth_order"] = order[out["tiles_touched"][order] >= 0].to(torch.int32)
        return out


def _gaussians(pos, opacity_raw, scale_raw, q_raw, sigma, f_dc, f_rest, color):
    keep = [_f32c(t) for t in (pos, opacity_raw, scale_raw, q_raw, sigma, f_dc, f_rest, color)]
    pos_, op_, sr_, q_, sg_, dc_, fr_, col_ = keep
    n = pos_.shape[0]
    if pos_.dim() != 2 or pos_.shape[1] != 3:
        raise ValueError(f"pos must be [N,3], got {tuple(pos_.shape)}")
    if op_.numel() != n:
        raise ValueError("opacity_raw must have one entry per Gaussian")
    if fr_ is not None and (fr_.dim() != 2 or fr_.shape[1] != 45):
        raise ValueError("f_rest must be [N,45] (spherical_harmonics.py:125-127)")
    # the kernels index every array with the same Gaussian id: a row-count mismatch (e.g. tags of a model that has
    # been densified since) would read out of bounds where the reference raises a shape error
    for name, t, tail in (("scale_raw", sr_, (3,)), ("q_raw", q_, (4,)), ("sigma", sg_, (3, 3)), ("f_dc", dc_, (3,)),
                          ("f_rest", fr_, (45,)), ("color", col_, (3,))):
        if t is not None and tuple(t.shape) != (n,) + tail:
            raise ValueError(f"{name} must be {(n,) + tail} to match pos [N={n},3], got {tuple(t.shape)}")
    if q_ is not None and q_.data_ptr() % 16:
        raise ValueError("q_raw must be 16-byte aligned (a contiguous [N,4] float32 tensor is)")
    g = Gaussians(n=n, pos=pos_.data_ptr(), opacity_raw=op_.data_ptr(),
                  scale_raw=None if sr_ is None else sr_.data_ptr(), q_raw=None if q_ is None else q_.data_ptr(),
                  sigma=None if sg_ is None else sg_.data_ptr(), f_dc=None if dc_ is None else dc_.data_ptr(),
                  f_rest=None if fr_ is None else fr_.data_ptr(), color=None if col_ is None else col_.data_ptr())
    return g, keep


class RoutedFrame(Frame):
    """A band of a tile-row sharded frame whose splat records were routed into `band_ws` by the ranks that projected
    them (b200gs_route_project_slice; dist.TileRowRenderer): the projection step is the gather + depth sort of the routed
    entries, everything behind it is the ordinary frame.  Entry ids are positions in the band workspace; forward only."""

    def __init__(self, route, cam_cfg: RenderConfig, c2w: torch.Tensor, device):
        g = Gaussians(n=int(route.world) * int(route.seg_capacity))
        super().__init__(g, None, cam_cfg, c2w, device)
        self.route = route
        self.cam.flags |= _lib.CAM_ROUTED
        self.after_project = None         # called between the depth sort and the rasterize phase (records-later barrier)

    def _project(self, lib, frame_bytes, stats_ptr, st):
        _lib.check(lib.b200gs_render_project_routed(ctypes.byref(self.cam), ctypes.byref(self.route), _ptr(self.frame_ws),
                                                    frame_bytes, stats_ptr, st), "render_project_routed")
        if self.after_project is not None:
            self.after_project()

    def backward(self, grad_image, grads):
        raise _lib.B200GSError("routed band frames are forward only")


def route_project_slice(pos, opacity_raw, scale_raw, q_raw, sigma, f_dc, f_rest, color, c2w, cfg: RenderConfig, route,
                        begin: int, end: int, buffers):
    """Source role of a routed tile-row frame: project Gaussians [begin, end) with the FULL frame's camera and route the
    survivors into the band workspaces named by `route`.  buffers = [slice workspace or None] (reused between frames)."""
    lib = _lib.load()
    sl = lambda t: None if t is None else t[begin:end]
    g, keep = _gaussians(sl(pos), sl(opacity_raw), sl(scale_raw), sl(q_raw), sl(sigma), sl(f_dc), sl(f_rest), sl(color))
    dev = pos.device
    with torch.cuda.device(dev):
        cam = cfg.to_c(c2w)
        frame_bytes, _ = _sizes(lib, int(g.n), int(cfg.H), int(cfg.W), 0)
        if buffers[0] is None or buffers[0].numel() < frame_bytes:
            buffers[0] = torch.empty(frame_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.b200gs_route_project_slice(ctypes.byref(g), ctypes.byref(cam), _ptr(buffers[0]), buffers[0].numel(),
                                                  ctypes.byref(route), _stream(dev)), "route_project_slice")
    return keep


def route_records(n_slice: int, c2w, cfg: RenderConfig, route, buffers):
    """Second half of the source role (route.flags has ROUTE_RECORDS_LATER): the splat records of the slice projected by
    the last `route_project_slice` on `buffers`, written on the CURRENT stream (ordered after that call by the caller)."""
    lib = _lib.load()
    dev = buffers[0].device
    with torch.cuda.device(dev):
        cam = cfg.to_c(c2w)
        _lib.check(lib.b200gs_route_records(int(n_slice), ctypes.byref(cam), _ptr(buffers[0]), buffers[0].numel(),
                                            ctypes.byref(route), _stream(dev)), "route_records")


def launch_frame(pos, opacity_raw, scale_raw, q_raw, sigma, f_dc, f_rest, color, c2w, cfg, buffers=None):
    """Forward-only frame without autograd bookkeeping (RenderPipeline): returns (image, Frame); the caller must
    call Frame.finish() before trusting the image."""
    g, keep = _gaussians(pos, opacity_raw, scale_raw, q_raw, sigma, f_dc, f_rest, color)
    with torch.cuda.device(pos.device):
        frame = Frame(g, keep, cfg, c2w, pos.device)
        image = frame.launch("speculative", buffers)
    return image, frame


# Gradient sinks: a caller that is going to hand the leaf gradients to other GPUs (b200gs.PeerAdam) registers, per leaf,
# a view of its peer-visible staging buffer; the backward then writes that leaf's gradient THERE and autograd adopts the
# tensor as `.grad` (it takes over a gradient it is handed instead of copying it), so the 236 MB staging copy of the
# optimizer step disappears.  A sink is handed out at most once between two optimizer steps and only while the leaf
# has no `.grad` yet (a second backward into the same leaf - several views per iteration - gets fresh tensors, which
# autograd adds into the first one in place).
_grad_sinks = {}      # data_ptr of the leaf -> [weakref(leaf), sink view, handed out]


def register_grad_sink(leaf: torch.Tensor, view: torch.Tensor) -> None:
    _grad_sinks[leaf.data_ptr()] = [weakref.ref(leaf), view, False]


def release_grad_sinks(leaves) -> None:
    """The optimizer has consumed the gradients: the sinks may be handed out again."""
    for p in leaves:
        e = _grad_sinks.get(p.data_ptr())
        if e is not None:
            e[2] = False


def unregister_grad_sinks(leaves) -> None:
    for p in leaves:
        _grad_sinks.pop(p.data_ptr(), None)


def _grad_sink_for(src: torch.Tensor, shape):
    e = _grad_sinks.get(src.data_ptr())
    if e is None or e[2]:
        return None
    leaf = e[0]()
    if leaf is None:
        _grad_sinks.pop(src.data_ptr(), None)
        return None
    if leaf.grad is not None or tuple(e[1].shape) != tuple(shape) or e[1].device != src.device:
        return None
    e[2] = True
    return e[1]


class _Rasterize(torch.autograd.Function):
    """render.py:62-410 (+ gaussian.py / spherical_harmonics.py when raw parameters are given)."""

    @staticmethod
    def forward(ctx, pos, opacity_raw, scale_raw, q_raw, sigma, f_dc, f_rest, color, c2w, cfg, strict):
        dev = pos.device
        g, keep = _gaussians(pos, opacity_raw, scale_raw, q_raw, sigma, f_dc, f_rest, color)
        with torch.cuda.device(dev):
            frame = Frame(g, keep, cfg, c2w, dev)
            image = frame.render()
        # (a band's counters describe the band - Gaussians that cannot touch it are dropped before the on-screen test -
        # so the reference's whole-frame exception is not raised for a band)
        band = cfg.tile_row_begin > 0 or cfg.tile_row_end > 0
        if strict and not band and frame.n_in_frustum > 0 and frame.n_visible == 0:
            raise Exception("All projected points are off-screen")      # render.py:235-236
        ctx.frame = frame
        ctx.shapes = [None if t is None else (t.shape, t.dtype) for t in
                      (pos, opacity_raw, scale_raw, q_raw, sigma, f_dc, f_rest, color)]
        # the backward recomputes the projection from the inputs' live memory (nothing per Gaussian is saved twice):
        # an in-place change between forward and backward must raise, as it does for tensors autograd saves itself
        ctx.inputs = [(name, t, t._version) for name, t in
                      zip(("pos", "opacity_raw", "scale_raw", "q_raw", "sigma", "f_dc", "f_rest", "color", "c2w"),
                          (pos, opacity_raw, scale_raw, q_raw, sigma, f_dc, f_rest, color, c2w)) if t is not None]
        return image

    @staticmethod
    def backward(ctx, grad_image):
        frame: Frame = ctx.frame
        for name, t, version in ctx.inputs:
            if t._version != version:
                raise RuntimeError(f"b200gs.render: `{name}`, needed for the gradient computation, has been modified by an "
                                   f"inplace operation since the forward (version {t._version}, expected {version})")
        dev = frame.device
        gi = _f32c(grad_image)
        n = int(frame.g.n)
        names = ("pos", "opacity_raw", "scale_raw", "q_raw", "sigma", "f_dc", "f_rest", "color")
        dims = dict(pos=(n, 3), opacity_raw=(n,), scale_raw=(n, 3), q_raw=(n, 4), sigma=(n, 3, 3), f_dc=(n, 3),
                    f_rest=(n, 45), color=(n, 3))
        out = {}
        for name, src in zip(names, frame.keep):
            if src is None:
                out[name] = None
                continue
            sink = _grad_sink_for(src, dims[name]) if _grad_sinks else None
            out[name] = sink if sink is not None else torch.empty(dims[name], dtype=torch.float32, device=dev)
        grads = Grads(**{k: (None if v is None else v.data_ptr()) for k, v in out.items()})
        with torch.cuda.device(dev):
            frame.backward(gi, grads)
        res = []
        for name, meta in zip(names, ctx.shapes):
            gten = out[name]
            if gten is None or meta is None:
                res.append(None)
            else:
                res.append(gten.reshape(meta[0]).to(meta[1]))
        # ctx.frame stays: backward(retain_graph=True) / a second autograd.grad on the same image must work as it does
        # with the reference's autograd graph; the workspaces are released with the graph node
        return (*res, None, None, None)


class _BuildSigma(torch.autograd.Function):
    """gaussian.py:71-127."""

    @staticmethod
    def forward(ctx, scale_raw, q_raw):
        lib = _lib.load()
        sr, q = _f32c(scale_raw), _f32c(q_raw)
        n = sr.shape[0]
        out = torch.empty((n, 3, 3), dtype=torch.float32, device=sr.device)
        with torch.cuda.device(sr.device):
            _lib.check(lib.b200gs_build_sigma(n, _ptr(sr), _ptr(q), _ptr(out), _stream(sr.device)), "build_sigma")
        ctx.save_for_backward(sr, q)
        ctx.dtypes = (scale_raw.dtype, q_raw.dtype)
        return out.to(scale_raw.dtype)

    @staticmethod
    def backward(ctx, g_sigma):
        lib = _lib.load()
        sr, q = ctx.saved_tensors
        n = sr.shape[0]
        gs = _f32c(g_sigma)
        g_sr, g_q = torch.empty_like(sr), torch.empty_like(q)
        with torch.cuda.device(sr.device):
            _lib.check(lib.b200gs_build_sigma_backward(n, _ptr(sr), _ptr(q), _ptr(gs), _ptr(g_sr), _ptr(g_q),
                                                       _stream(sr.device)), "build_sigma_backward")
        return g_sr.to(ctx.dtypes[0]), g_q.to(ctx.dtypes[1])


class _EvaluateSH(torch.autograd.Function):
    """spherical_harmonics.py:70-166."""

    @staticmethod
    def forward(ctx, f_dc, f_rest, points, c2w):
        lib = _lib.load()
        dc, fr, pts, cw = _f32c(f_dc), _f32c(f_rest), _f32c(points), _f32c(c2w)
        n = pts.shape[0]
        out = torch.empty((n, 3), dtype=torch.float32, device=pts.device)
        with torch.cuda.device(pts.device):
            _lib.check(lib.b200gs_evaluate_sh(n, _ptr(dc), _ptr(fr), _ptr(pts), _ptr(cw), _ptr(out),
                                              _stream(pts.device)), "evaluate_sh")
        ctx.save_for_backward(dc, fr, pts, cw)
        ctx.dtypes = (f_dc.dtype, f_rest.dtype, points.dtype)
        return out.to(points.dtype)

    @staticmethod
    def backward(ctx, g_color):
        lib = _lib.load()
        dc, fr, pts, cw = ctx.saved_tensors
        n = pts.shape[0]
        gc = _f32c(g_color)
        g_dc, g_fr, g_pts = torch.empty_like(dc), torch.empty_like(fr), torch.empty_like(pts)
        with torch.cuda.device(pts.device):
            _lib.check(lib.b200gs_evaluate_sh_backward(n, _ptr(dc), _ptr(fr), _ptr(pts), _ptr(cw), _ptr(gc), _ptr(g_dc),
                                                       _ptr(g_fr), _ptr(g_pts), _stream(pts.device)),
                       "evaluate_sh_backward")
        return g_dc.to(ctx.dtypes[0]), g_fr.to(ctx.dtypes[1]), g_pts.to(ctx.dtypes[2]), None
