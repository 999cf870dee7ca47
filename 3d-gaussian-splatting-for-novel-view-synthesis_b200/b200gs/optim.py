"""Fused optimizer step for the Gaussian parameter set.

  scripts/train.py:394-401   optim.Adam([{'params': [model.pos], 'lr': ...}, ... six groups ...], lr=lr, eps=1e-15)
  scripts/train.py:536       torch.nn.utils.clip_grad_norm_(model.pos, max_norm=1.0)
  scripts/train.py:538       optimizer.step()

`FusedAdam` takes the same constructor arguments as `torch.optim.Adam` and keeps the same per-parameter
state (`step`, `exp_avg`, `exp_avg_sq`: state dicts are interchangeable), but updates every tensor of every
group in ONE kernel launch that touches each array once (28 B per element); `clip_grad_norm_` is the
reference's call without its host round trip.  CUDA fp32 parameters only; anything else raises - there is
no fallback.  The reference script constructs `optim.Adam` itself, so using this class is opt-in:
`b200gs.install(optimizer=True)` (or `B200GS_PATCH_ADAM=1` with `python -m b200gs.run`) rebinds
`torch.optim.Adam` for the process.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **unsupported):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("b200gs.FusedAdam implements the reference's configuration: weight_decay=0, "
                                      "amsgrad=False")
        for k, v in unsupported.items():
            if v not in (None, False):
                raise NotImplementedError(f"b200gs.FusedAdam does not support {k}={v!r}")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        # one launch per distinct (betas, eps, device) - a single one for the reference's groups
        buckets = {}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise _lib.B200GSError("b200gs.FusedAdam: parameters and gradients must be dense CUDA fp32 tensors "
                                           "(no CPU or mixed-precision fallback)")
                if not p.is_contiguous():
                    raise _lib.B200GSError("b200gs.FusedAdam: parameters must be contiguous")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                key = (tuple(group["betas"]), float(group["eps"]), p.device)
                buckets.setdefault(key, []).append((p, g, st, float(group["lr"])))
        for (betas, eps, dev), items in buckets.items():
            table = (_lib.AdamTensor * len(items))()
            for i, (p, g, st, lr) in enumerate(items):
                table[i] = _lib.AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                           st["exp_avg_sq"].data_ptr(), p.numel(), lr, int(st["step"]), 0)
            with torch.cuda.device(dev):
                _lib.check(lib.b200gs_adam_step(table, len(items), betas[0], betas[1], eps, ops._stream(dev)), "adam_step")
        return loss


_clip_ws = {}


def clip_grad_norm_(parameters, max_norm, norm_type=2.0, error_if_nonfinite=False, foreach=None):
    """torch.nn.utils.clip_grad_norm_ for CUDA fp32 gradients (L2 norm), no host sync.  Returns the total norm
    as a 0-dim device tensor.  One tensor (the reference clips `model.pos` only) is a norm kernel + an in-place
    scale; several tensors are clipped by their joint norm like torch does, by the same two kernels."""
    if float(norm_type) != 2.0 or error_if_nonfinite:
        raise NotImplementedError("b200gs.clip_grad_norm_ implements the L2 norm without error_if_nonfinite")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    lib = _lib.load()
    for g in grads:
        if not g.is_cuda or g.dtype != torch.float32 or not g.is_contiguous():
            raise _lib.B200GSError("b200gs.clip_grad_norm_: gradients must be contiguous CUDA fp32 tensors")
    dev = grads[0].device
    if any(g.device != dev for g in grads):
        raise _lib.B200GSError("b200gs.clip_grad_norm_: all gradients must live on one device")
    if len(grads) > _lib.CLIP_MAX_TENSORS:
        raise NotImplementedError(f"b200gs.clip_grad_norm_ clips at most {_lib.CLIP_MAX_TENSORS} tensors per call "
                                  f"(got {len(grads)})")
    # one tensor (the reference: model.pos) or several: the same two kernels (sum of squares over all tensors with the
    # coefficient computed by the last block, then one in-place scale), the tensors given as a table of pointers
    ptrs = (ctypes.c_void_p * len(grads))(*[g.data_ptr() for g in grads])
    numels = (ctypes.c_int64 * len(grads))(*[g.numel() for g in grads])
    nbytes = int(lib.b200gs_clip_workspace_bytes_multi(numels, len(grads)))
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _clip_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _clip_ws[key] = ws
    total = torch.empty((), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.b200gs_clip_grad_norm_multi(ptrs, numels, len(grads), float(max_norm), ops._ptr(ws), ws.numel(),
                                                   ops._ptr(total), ops._stream(dev)), "clip_grad_norm")
    return total
