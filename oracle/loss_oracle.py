"""CPU oracle for the training loss (L1 + SSIM).  TEST INFRASTRUCTURE ONLY.

CPU restatement (torch CPU ops) of the reference's `gaussian_splatting/losses.py`, the consumer that
sits right after the render path in every training iteration (scripts/train.py:511; SURVEY.md
section 8f row N2).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may
import it; the product (`b200gs.compute_loss`) runs its own CUDA kernels and never touches this file.

Parity pin: the reference ships no tests or vectors for the loss either, so the pin is the reference
itself - `oracle/make_golden_loss.py` imports the unmodified reference in the build container, asserts
that this restatement reproduces its three loss values and its autograd gradient exactly, and commits
the vectors under `tests/golden/loss_*.npz`.

Reference lines followed (paths relative to the reference checkout):
  gaussian_splatting/losses.py:27-41    l1_loss          = F.l1_loss (mean absolute error)
  gaussian_splatting/losses.py:44-88    ssim_loss        = 1 - mean over channels of the per-channel SSIM mean
  gaussian_splatting/losses.py:91-132   _ssim_single_channel: 11x11 Gaussian window (sigma 1.5), ZERO padding,
                                        C1 = 0.01^2, C2 = 0.03^2
  gaussian_splatting/losses.py:135-155  _create_gaussian_window: g = exp(-c^2 / (2 sigma^2)) / sum, outer product
  gaussian_splatting/losses.py:158-185  compute_loss     = lambda_l1 * l1 + lambda_ssim * ssim, plus a dict of floats
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

C1 = 0.01 ** 2
C2 = 0.03 ** 2


def gaussian_window_1d(window_size: int = 11, sigma: float = 1.5, dtype=torch.float32) -> torch.Tensor:
    """losses.py:148-151 - the normalised 1-D window the 2-D one is the outer product of."""
    coords = torch.arange(window_size, dtype=dtype) - window_size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def l1_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """losses.py:27-41."""
    return (pred - target).abs().mean()


def ssim_map_single_channel(x: torch.Tensor, y: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    """losses.py:91-130 for one [B,1,H,W] channel, returning the SSIM map (the reference returns its mean)."""
    g = gaussian_window_1d(window_size, 1.5, x.dtype).to(x.device)
    win = (g.unsqueeze(1) * g.unsqueeze(0)).unsqueeze(0).unsqueeze(0)
    pad = window_size // 2
    mu1 = F.conv2d(x, win, padding=pad)
    mu2 = F.conv2d(y, win, padding=pad)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    s1 = F.conv2d(x * x, win, padding=pad) - mu1_sq
    s2 = F.conv2d(y * y, win, padding=pad) - mu2_sq
    s12 = F.conv2d(x * y, win, padding=pad) - mu1_mu2
    return ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))


def ssim_loss(pred: torch.Tensor, target: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    """losses.py:44-88 - [H,W,3] or [B,H,W,3] in, scalar out."""
    if pred.dim() == 3:
        pred, target = pred.unsqueeze(0), target.unsqueeze(0)
    p, t = pred.permute(0, 3, 1, 2), target.permute(0, 3, 1, 2)
    per_channel = [ssim_map_single_channel(p[:, c:c + 1], t[:, c:c + 1], window_size).mean() for c in range(p.shape[1])]
    return 1 - torch.stack(per_channel).mean()


def compute_loss(pred: torch.Tensor, target: torch.Tensor, lambda_l1: float = 0.8, lambda_ssim: float = 0.2):
    """losses.py:158-185."""
    l1 = l1_loss(pred, target)
    ssim = ssim_loss(pred, target)
    total = lambda_l1 * l1 + lambda_ssim * ssim
    return total, {"l1": l1.item(), "ssim": ssim.item(), "total": total.item()}
