"""CUDA-event times of the fused L1 + SSIM loss (forward, forward + backward) at 1080p, next to the reference's
formulation in torch ops on the same GPU (15 conv2d + elementwise; oracle/loss_oracle.py = losses.py restated)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch  # noqa: E402
import b200gs  # noqa: E402
from oracle import loss_oracle as L  # noqa: E402

H, W = 1080, 1920
pred = torch.rand(H, W, 3, device="cuda")
target = torch.rand(H, W, 3, device="cuda")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def mine_fwd():
    with torch.no_grad():
        b200gs.compute_loss_tensors(pred, target)


def mine_fb():
    p = pred.clone().requires_grad_(True)
    b200gs.compute_loss_tensors(p, target)[0].backward()


def ref_fb():
    p = pred.clone().requires_grad_(True)
    (0.8 * L.l1_loss(p, target) + 0.2 * L.ssim_loss(p, target)).backward()


print(f"b200gs forward {timeit(mine_fwd):.1f} us, forward+backward {timeit(mine_fb):.1f} us; "
      f"reference formulation in torch ops (same GPU) forward+backward {timeit(ref_fb, 5):.1f} us")
