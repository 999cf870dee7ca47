"""Host-side mirror of the reference's render-path interface (same names, arguments and error
behaviour), backed by the CUDA kernels of libb200gs.

Reference signatures mirrored (paths relative to the reference checkout):
  gaussian_splatting/gaussian.py:71             build_sigma_from_params(scale_raw, q_raw) -> [N,3,3]
  gaussian_splatting/spherical_harmonics.py:70  evaluate_sh(f_dc, f_rest, points, c2w) -> [N,3]
  gaussian_splatting/render.py:62-64            render(pos, color, opacity_raw, sigma, c2w, H, W, fx, fy, cx, cy,
                                                       near, far, pix_guard, T, min_conis, chi_square_clip,
                                                       alpha_max, alpha_cutoff) -> [H,W,3]

Fusion across the three calls: the tensors returned by `build_sigma_from_params` / `evaluate_sh` carry
a tag naming the leaves they were computed from.  When `render` receives tagged `sigma` / `color` whose
sources are unchanged (and `color` was evaluated at the same points and pose), it runs the fused
kernels straight from the raw parameters, so gradients flow to scale_raw / q_raw / f_dc / f_rest / pos
in one preprocess-backward kernel; otherwise it consumes `sigma` / `color` as given, exactly like the
reference.  Callers need no change (scripts/train.py:463,502,505-508 keep working verbatim).
"""
from __future__ import annotations

import os
import weakref

import torch

from . import _lib, ops

# Tags (what a derived tensor was computed from) live in a side table keyed by the tensor's identity, NOT in the
# tensor's __dict__: Python state on a tensor changes how torch pickles it (torch.save of a tagged tensor would
# need the tag to be picklable and loadable under weights_only), and a saved / copied tensor must be a plain tensor.
_tags = {}


def _set_tag(t, src) -> None:
    key = id(t)
    _tags[key] = (weakref.ref(t, lambda _r, k=key: _tags.pop(k, None)), src)


def _get_tag(t):
    e = _tags.get(id(t))
    return e[1] if e is not None and e[0]() is t else None


def _fusion_enabled() -> bool:
    return _lib.env("B200GS_FUSE", "1") != "0"


def _lazy_enabled() -> bool:
    return _lib.env("B200GS_LAZY", "1") != "0"


_META_GETTERS = {"shape", "dtype", "device", "requires_grad", "ndim", "is_cuda", "layout", "is_leaf", "names",
                 "is_sparse", "is_quantized", "is_meta", "grad_fn", "_version", "is_nested"}
_META_METHODS = {"size", "dim", "numel", "ndimension", "nelement", "element_size", "is_floating_point",
                 "is_complex", "is_contiguous", "get_device", "stride", "storage_offset", "__len__"}


class _Deferred(torch.Tensor):
    """Result of `build_sigma_from_params` / `evaluate_sh` whose kernel has not run yet.

    It has the right shape/dtype/device and behaves like the tensor it stands for: the first torch
    operation that touches it runs the stand-alone kernel (with autograd recorded) and proceeds on the
    real tensor.  `render` recognises it and, when it can fuse, never materialises it - so the
    reference's call sequence costs one fused preprocess kernel instead of three passes over the
    parameters."""

    @staticmethod
    def __new__(cls, thunk, shape, dtype, device, requires_grad):
        t = torch.Tensor._make_wrapper_subclass(cls, shape, dtype=dtype, device=device, requires_grad=False)
        t._thunk, t._value, t._grad_mode, t._wants_grad = thunk, None, torch.is_grad_enabled(), requires_grad
        return t

    def materialize(self) -> torch.Tensor:
        if self._value is None:
            src = _get_tag(self)
            if src is not None and src.resolve() is None:
                # the reference computes sigma / colours at the call; computing them now from modified parameters
                # would silently give different values
                raise RuntimeError("b200gs: the parameters this tensor was computed from have been modified in place "
                                   "since build_sigma_from_params / evaluate_sh was called (set B200GS_LAZY=0 for eager "
                                   "evaluation at the call)")
            with torch.set_grad_enabled(self._grad_mode):
                self._value = self._thunk()
            src = _get_tag(self)
            if src is not None:
                _set_tag(self._value, src)
            self._thunk = None
        return self._value

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        owner = getattr(func, "__self__", None)
        if name == "__get__" and getattr(owner, "__name__", "") in _META_GETTERS and \
                getattr(owner, "__name__", "") not in ("requires_grad", "grad_fn", "is_leaf"):
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if name in _META_METHODS:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        unwrap = lambda x: x.materialize() if isinstance(x, _Deferred) else x
        args = torch.utils._pytree.tree_map(unwrap, args)
        kwargs = torch.utils._pytree.tree_map(unwrap, kwargs)
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **kwargs)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # reached only when something bypasses __torch_function__: same policy, materialise and go on
        unwrap = lambda x: x.materialize().detach() if isinstance(x, _Deferred) else x
        args = torch.utils._pytree.tree_map(unwrap, args)
        kwargs = torch.utils._pytree.tree_map(unwrap, kwargs or {})
        return func(*args, **kwargs)

    def __reduce_ex__(self, proto):
        # torch.save / pickle / copy.deepcopy store the tensor this object stands for (a plain tensor)
        return self.materialize().__reduce_ex__(proto)

    def __repr__(self):
        return f"_Deferred({'pending' if self._value is None else 'materialized'}, shape={tuple(self.shape)})"


def _real(t):
    return t.materialize() if isinstance(t, _Deferred) else t


class _Source:
    """What a derived tensor was computed from (weak references + version counters)."""

    def __init__(self, *tensors):
        self.refs = [weakref.ref(t) for t in tensors]
        self.versions = [t._version for t in tensors]

    def resolve(self):
        out = []
        for r, v in zip(self.refs, self.versions):
            t = r()
            if t is None or t._version != v:
                return None
            out.append(t)
        return out


def build_sigma_from_params(scale_raw: torch.Tensor, q_raw: torch.Tensor) -> torch.Tensor:
    """Sigma = R S S^T R^T with s = max(exp(scale_raw), 1e-6), q normalised (gaussian.py:71-127)."""
    ops._require_cuda(scale_raw, "scale_raw")
    ops._require_f32("scale_raw", scale_raw)
    ops._require_f32("q_raw", q_raw)
    if _lazy_enabled() and _fusion_enabled():
        n = scale_raw.shape[0]
        needs = torch.is_grad_enabled() and (scale_raw.requires_grad or q_raw.requires_grad)
        sigma = _Deferred(lambda: ops._BuildSigma.apply(scale_raw, q_raw), (n, 3, 3), scale_raw.dtype,
                          scale_raw.device, needs)
    else:
        sigma = ops._BuildSigma.apply(scale_raw, q_raw)
    _set_tag(sigma, _Source(scale_raw, q_raw))
    return sigma


def evaluate_sh(f_dc: torch.Tensor, f_rest: torch.Tensor, points: torch.Tensor, c2w: torch.Tensor) -> torch.Tensor:
    """sigmoid(sum_k sh_k Y_k(dir)), degree-3 real SH with the reference's signs (spherical_harmonics.py:70-166)."""
    ops._require_cuda(points, "points")
    for name, t in (("f_dc", f_dc), ("f_rest", f_rest), ("points", points)):
        ops._require_f32(name, t)
    if f_rest.dim() != 2 or f_rest.shape[1] != 45:
        raise RuntimeError(f"evaluate_sh expects f_rest of shape [N,45], got {tuple(f_rest.shape)}")
    if _lazy_enabled() and _fusion_enabled():
        needs = torch.is_grad_enabled() and (f_dc.requires_grad or f_rest.requires_grad or points.requires_grad)
        color = _Deferred(lambda: ops._EvaluateSH.apply(f_dc, f_rest, points, c2w.to(device=points.device)),
                          (points.shape[0], 3), points.dtype, points.device, needs)
    else:
        color = ops._EvaluateSH.apply(f_dc, f_rest, points, c2w.to(device=points.device))
    _set_tag(color, _Source(f_dc, f_rest, points, c2w))
    return color


def render(pos, color, opacity_raw, sigma, c2w, H, W, fx, fy, cx, cy,
           near=0.01, far=100.0, pix_guard=32, T=16, min_conis=1e-6,
           chi_square_clip=6.25, alpha_max=0.99, alpha_cutoff=1 / 128., *, tile_rows=None, out=None):
    """Drop-in for render.py:62-410.  Returns [H,W,3] in [0,1], same dtype/device as `pos`.

    `tile_rows=(begin, end)` (keyword-only extension) renders only that band of 16-pixel tile rows; the
    rest of the image is zero (tile-row sharding of large frames across GPUs).  `out` (keyword-only): a
    contiguous [H,W,3] float32 CUDA tensor to render into."""
    args, strict = _resolve(pos, color, opacity_raw, sigma, c2w, H, W, fx, fy, cx, cy, near, far, pix_guard, T, min_conis,
                            chi_square_clip, alpha_max, alpha_cutoff, tile_rows)
    if out is not None:
        args[-1].out = out
    image = ops._Rasterize.apply(*args, strict)
    return image if image.dtype == pos.dtype else image.to(pos.dtype)


def _resolve(pos, color, opacity_raw, sigma, c2w, H, W, fx, fy, cx, cy, near, far, pix_guard, T, min_conis,
             chi_square_clip, alpha_max, alpha_cutoff, tile_rows):
    """Arguments of render() -> (pos, opacity_raw, scale_raw, q_raw, sigma, f_dc, f_rest, color, c2w, cfg), strict:
    the fused route (raw parameters) when the tags of `sigma` / `color` resolve, the tensors themselves otherwise."""
    ops._require_cuda(pos, "pos")
    for name, t in (("pos", pos), ("color", color), ("opacity_raw", opacity_raw), ("sigma", sigma)):
        if type(t) is not _Deferred:                   # a deferred tensor's dtype was checked when it was created
            ops._require_f32(name, t)
    H, W = int(H), int(W)           # callers pass Python ints, 0-dim tensors (train.py:499) or floats
    cfg = ops.RenderConfig(H=H, W=W, fx=float(fx), fy=float(fy), cx=float(cx), cy=float(cy), near=float(near),
                           far=float(far), pix_guard=float(pix_guard), T=int(T), min_conis=float(min_conis),
                           chi_square_clip=float(chi_square_clip), alpha_max=float(alpha_max),
                           alpha_cutoff=float(alpha_cutoff))
    if tile_rows is not None:
        cfg.tile_row_begin, cfg.tile_row_end = int(tile_rows[0]), int(tile_rows[1])
    c2w_d = ops._f32c(c2w.to(device=pos.device))
    scale_raw = q_raw = f_dc = f_rest = None
    if _fusion_enabled():
        src = _get_tag(sigma)
        got = src.resolve() if src is not None else None
        if got is not None:
            scale_raw, q_raw = got
        src = _get_tag(color)
        got = src.resolve() if src is not None else None
        if got is not None and got[2] is pos and (got[3] is c2w or torch.equal(got[3].to(c2w_d.device), c2w_d)):
            f_dc, f_rest = got[0], got[1]
    strict = _lib.env("B200GS_STRICT_OFFSCREEN", "1") != "0"
    return (pos, opacity_raw, scale_raw, q_raw, None if scale_raw is not None else _real(sigma),
            f_dc, f_rest, None if f_dc is not None else _real(color), c2w_d, cfg), strict


def to_uint8(image: torch.Tensor) -> torch.Tensor:
    """Frame sink: `(image * 255).astype(uint8)` - the conversion the reference scripts do on the host after
    downloading the fp32 image (render_trained.py:357, inference.py:117) - done on the device, so a frame costs
    3 bytes per pixel on PCIe instead of 12.  Returns a uint8 tensor of the same shape on the same device."""
    from . import _lib
    ops._require_cuda(image, "image")
    img = ops._f32c(image.detach())
    out = torch.empty(img.shape, dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.load().b200gs_image_to_u8(ops._ptr(img), ops._ptr(out), img.numel(), ops._stream(img.device)),
                   "image_to_u8")
    return out


class RenderPipeline:
    """Software pipelining of a SEQUENCE of independent frames (an orbit render: scripts/render_trained.py:333-358,
    scripts/inference.py:100-119) on one GPU.

    A frame is two halves with opposite bottlenecks: project + binning (one HBM-bound kernel and seven small,
    latency-bound ones) and the blend (FP32-issue bound, nearly no HBM traffic).  Frames are queued with the first
    half on a high-priority stream and the blend on a second stream, so the binning of frame i+1 runs while frame i
    is being blended, and the host never waits for a frame it has just queued:

        pipe = b200gs.RenderPipeline()
        t = pipe.submit(pos, colors, opacity_raw, sigma, c2w, H, W, fx, fy, cx, cy)     # returns at once
        ...                                                                              # submit the next frame(s)
        img = pipe.result(t)        # waits for that frame's COUNTERS only (re-rasterizes if the lists overflowed)

    `render()` = submit + result.  Same arguments and image as `b200gs.render` (forward only).  The image is valid
    in `pipe.blend_stream` order: consume it under `torch.cuda.stream(pipe.blend_stream)` or after
    `pipe.synchronize()`.  The Gaussian parameters must not be modified while frames are in flight."""

    MAX_IN_FLIGHT = 4

    def __init__(self, device=None, front_streams=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        _, high = torch.cuda.Stream.priority_range()
        if front_streams is None:
            front_streams = int(os.environ.get("B200GS_FRONT_STREAMS", "2"))
        # the binning chain is a dozen dependent, latency-bound kernels: with two front streams the chains of frames
        # i+1 and i+2 interleave beside the blend of frame i (the caller must then keep `lag` frames queued)
        self.front_streams = [torch.cuda.Stream(self.device, priority=high) for _ in range(max(1, int(front_streams)))]
        self.front_stream = self.front_streams[0]
        self.lag = len(self.front_streams)
        self.blend_stream = torch.cuda.Stream(self.device)
        self._open = []                               # submitted, not yet checked (oldest first)
        # ring of workspace sets ([frame_ws, isect_ws], completion event of their last user): the ~260 MB of a frame
        # do not go through the caching allocator every frame (buffers used on two streams return to it slowly)
        self._ring = [[[None, None], None] for _ in range(self.MAX_IN_FLIGHT + 1)]
        self._next = 0
        torch.cuda.synchronize(self.device)          # everything created so far is visible to both streams

    def submit(self, pos, color, opacity_raw, sigma, c2w, H, W, fx, fy, cx, cy, near=0.01, far=100.0, pix_guard=32,
               T=16, min_conis=1e-6, chi_square_clip=6.25, alpha_max=0.99, alpha_cutoff=1 / 128.):
        while len(self._open) >= self.MAX_IN_FLIGHT:
            self._check(self._open[0])
        prev = ops._blend_stream.get(self.device.index)
        ops._blend_stream[self.device.index] = self.blend_stream
        front = self.front_streams[self._next % len(self.front_streams)]
        try:
            with torch.no_grad(), torch.cuda.stream(front):
                c2w_d = c2w.to(self.device, non_blocking=True)
                args, strict = _resolve(pos, color, opacity_raw, sigma, c2w_d, H, W, fx, fy, cx, cy, near, far, pix_guard,
                                        T, min_conis, chi_square_clip, alpha_max, alpha_cutoff, None)
                slot = self._ring[self._next % len(self._ring)]
                self._next += 1
                if slot[1] is not None:
                    front.wait_event(slot[1])                  # the frame that last used these workspaces is blended
                image, frame = ops.launch_frame(*args, buffers=slot[0])
        finally:
            if prev is None:
                ops._blend_stream.pop(self.device.index, None)
            else:
                ops._blend_stream[self.device.index] = prev
        done = torch.cuda.Event()
        done.record(self.blend_stream)                # the frame's blend has been queued: completion marker
        slot[1] = done
        ticket = [image, frame, strict, pos.dtype, done, slot]
        self._open.append(ticket)
        return ticket

    def _check(self, ticket):
        idx = next((i for i, t in enumerate(self._open) if t is ticket), None)
        if idx is None:
            return
        del self._open[idx]
        prev = ops._blend_stream.get(self.device.index)
        ops._blend_stream[self.device.index] = self.blend_stream
        try:
            if ticket[1].finish():                    # rasterized again: the completion marker moves
                ticket[4] = torch.cuda.Event()
                ticket[4].record(self.blend_stream)
                ticket[5][1] = ticket[4]
        finally:
            if prev is None:
                ops._blend_stream.pop(self.device.index, None)
            else:
                ops._blend_stream[self.device.index] = prev

    def done_event(self, ticket):
        """CUDA event that completes when the frame's image is ready (call after `result`): lets a consumer on
        another stream (e.g. a device-to-host copy) wait for exactly this frame, not for the blend stream's tail."""
        self._check(ticket)
        return ticket[4]

    def result(self, ticket):
        self._check(ticket)
        image, frame, strict, dtype = ticket[:4]
        if strict and frame.n_in_frustum > 0 and frame.n_visible == 0:
            raise Exception("All projected points are off-screen")      # render.py:235-236
        return image if image.dtype == dtype else image.to(dtype)

    def render(self, *args, **kwargs):
        return self.result(self.submit(*args, **kwargs))

    def synchronize(self):
        for t in list(self._open):
            self._check(t)
        for st in self.front_streams:
            st.synchronize()
        self.blend_stream.synchronize()

    @property
    def next_front_stream(self):
        """The stream the NEXT `submit` queues its first half on: inputs of that frame that are produced
        asynchronously (e.g. the pose uploaded from pinned memory) belong on this stream."""
        return self.front_streams[self._next % len(self.front_streams)]

    def wait_event(self, event):
        """Every stream of the pipeline waits for `event` (e.g. the start marker of a timed region)."""
        for st in self.front_streams:
            st.wait_event(event)
        self.blend_stream.wait_event(event)

    def join(self, stream):
        """`stream` waits for everything queued on the pipeline's streams so far."""
        for st in self.front_streams:
            stream.wait_stream(st)
        stream.wait_stream(self.blend_stream)
