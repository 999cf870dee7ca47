"""Host-side mirror of the reference's training loss (gaussian_splatting/losses.py), backed by the fused
L1 + SSIM kernels of libb200gs (csrc/loss.cu).  Same names, arguments and return values:

  gaussian_splatting/losses.py:27    l1_loss(pred, target) -> scalar tensor
  gaussian_splatting/losses.py:44    ssim_loss(pred, target, window_size=11, size_average=True) -> scalar tensor
  gaussian_splatting/losses.py:158   compute_loss(pred, target, lambda_l1=0.8, lambda_ssim=0.2)
                                     -> (total loss tensor, {'l1': float, 'ssim': float, 'total': float})

`pred` / `target` are [H,W,3] or [B,H,W,3] (the render output layout).  The reference spends fifteen
conv2d calls, ~40 elementwise kernels and three `.item()` syncs per view here (scripts/train.py:511);
this is one kernel forward, one backward and one 12-byte read-back.  CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops


class _L1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, lambda_l1, lambda_ssim):
        lib = _lib.load()
        p, t = ops._f32c(pred), ops._f32c(target)
        if p.dim() == 3:
            p, t = p.unsqueeze(0), t.unsqueeze(0)
        if p.dim() != 4 or p.shape[-1] != 3 or p.shape != t.shape:
            raise ValueError(f"pred/target must both be [H,W,3] or [B,H,W,3], got {tuple(pred.shape)} and "
                             f"{tuple(target.shape)}")
        B, H, W, _ = p.shape
        with_grad = bool(ctx.needs_input_grad[0])
        nbytes = int(lib.b200gs_loss_workspace_bytes(B, H, W, int(with_grad)))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=p.device)
        out = torch.empty(3, dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            _lib.check(lib.b200gs_l1_ssim_forward(ops._ptr(p), ops._ptr(t), B, H, W, float(lambda_l1), float(lambda_ssim),
                                                  ops._ptr(ws), nbytes, int(with_grad), ops._ptr(out),
                                                  ops._stream(p.device)), "l1_ssim_forward")
        ctx.save_for_backward(p, t, ws)
        ctx.meta = (B, H, W, float(lambda_l1), float(lambda_ssim), pred.shape, pred.dtype)
        ctx.mark_non_differentiable(out)
        return out[2], out

    @staticmethod
    def backward(ctx, g_total, _g_out):
        lib = _lib.load()
        p, t, ws = ctx.saved_tensors
        B, H, W, l1w, sw, shape, dtype = ctx.meta
        g = ops._f32c(g_total)
        grad = torch.empty_like(p)
        with torch.cuda.device(p.device):
            _lib.check(lib.b200gs_l1_ssim_backward(ops._ptr(p), ops._ptr(t), B, H, W, l1w, sw, ops._ptr(ws), ws.numel(),
                                                   ops._ptr(g), ops._ptr(grad), ops._stream(p.device)),
                       "l1_ssim_backward")
        return grad.reshape(shape).to(dtype), None, None, None


def _check(pred, target):
    ops._require_cuda(pred, "pred")
    ops._require_cuda(target, "target")


def compute_loss(pred, target, lambda_l1=0.8, lambda_ssim=0.2):
    """losses.py:158-185: (lambda_l1 * L1 + lambda_ssim * (1 - SSIM), {'l1', 'ssim', 'total'} as Python floats)."""
    _check(pred, target)
    total, out = _L1SSIM.apply(pred, target.detach(), lambda_l1, lambda_ssim)
    vals = out.tolist()                      # one 12-byte read-back (the reference does three .item() syncs)
    return total, {"l1": vals[0], "ssim": vals[1], "total": vals[2]}


def compute_loss_tensors(pred, target, lambda_l1=0.8, lambda_ssim=0.2):
    """Same loss without the host read-back: (total, device tensor [l1, ssim, total])."""
    _check(pred, target)
    return _L1SSIM.apply(pred, target.detach(), lambda_l1, lambda_ssim)


def l1_loss(pred, target):
    """losses.py:27-41."""
    _check(pred, target)
    return _L1SSIM.apply(pred, target.detach(), 1.0, 0.0)[0]


def ssim_loss(pred, target, window_size=11, size_average=True):
    """losses.py:44-88 (the reference ignores size_average; only its default 11x11 window is built)."""
    if int(window_size) != 11:
        raise NotImplementedError("b200gs.ssim_loss implements the reference's default 11x11 window only")
    _check(pred, target)
    return _L1SSIM.apply(pred, target.detach(), 0.0, 1.0)[0]
