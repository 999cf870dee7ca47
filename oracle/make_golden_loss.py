"""Pins oracle/loss_oracle.py to the unmodified reference (gaussian_splatting/losses.py) and writes the
golden vectors tests/golden/loss_*.npz.  Runs in the build container only (needs /root/reference).

    python oracle/make_golden_loss.py
"""
import json
import os
import sys

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = json.load(open(os.path.join(ROOT, "BASELINE.json")))["reference_path"]
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from gaussian_splatting import losses as ref_losses  # noqa: E402  (the unmodified reference)
from oracle import loss_oracle as L  # noqa: E402

# (name, H, W, seed, kind): smooth = blurred random fields (image-like), noise = white noise, equal = pred == target
# on part of the frame (sign(0) in the L1 gradient), edge sizes exercise partial 32x32 kernel tiles and the zero padding
CASES = [("loss_smooth_97x71", 71, 97, 1, "smooth"), ("loss_noise_64x48", 48, 64, 2, "noise"),
         ("loss_equal_40x33", 33, 40, 3, "equal"), ("loss_tiny_7x5", 5, 7, 4, "noise"),
         ("loss_wide_130x20", 20, 130, 5, "smooth")]


def make_pair(H, W, seed, kind):
    g = torch.Generator().manual_seed(seed)
    a, b = torch.rand(H, W, 3, generator=g), torch.rand(H, W, 3, generator=g)
    if kind == "smooth":
        k = torch.ones(1, 1, 5, 5) / 25.0
        blur = lambda t: torch.nn.functional.conv2d(t.permute(2, 0, 1).unsqueeze(1), k, padding=2).squeeze(1).permute(1, 2, 0)
        a = blur(a)
        b = (0.8 * a + 0.2 * blur(b)).clamp(0, 1)
    if kind == "equal":
        b = b.clone()
        b[: H // 2] = a[: H // 2]
    return a.contiguous(), b.contiguous()


def main():
    summary = {}
    for name, H, W, seed, kind in CASES:
        pred, target = make_pair(H, W, seed, kind)
        p_ref = pred.clone().requires_grad_(True)
        total_ref, d_ref = ref_losses.compute_loss(p_ref, target)
        total_ref.backward()
        p_or = pred.clone().requires_grad_(True)
        total_or, d_or = L.compute_loss(p_or, target)
        total_or.backward()
        assert d_ref == d_or, (name, d_ref, d_or)                         # the three values, exactly
        assert torch.equal(p_ref.grad, p_or.grad), name                   # and the autograd gradient
        assert float(ref_losses.l1_loss(pred, target)) == float(L.l1_loss(pred, target))
        assert float(ref_losses.ssim_loss(pred, target)) == float(L.ssim_loss(pred, target))
        # other weights + fp64 arbiter
        t2, _ = ref_losses.compute_loss(pred, target, lambda_l1=0.3, lambda_ssim=0.7)
        p64 = pred.double().requires_grad_(True)
        t64, d64 = L.compute_loss(p64, target.double())
        t64.backward()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), pred=pred.numpy(), target=target.numpy(),
                            l1=np.float32(d_ref["l1"]), ssim=np.float32(d_ref["ssim"]), total=np.float32(d_ref["total"]),
                            total_03_07=np.float32(float(t2)), grad=p_ref.grad.numpy(), grad64=p64.grad.numpy(),
                            l1_64=d64["l1"], ssim_64=d64["ssim"], total_64=d64["total"])
        g, g64 = p_ref.grad.double(), p64.grad
        summary[name] = {**d_ref, "oracle_equals_reference": True,
                         "ref32_vs_64_grad_maxrel": float((g - g64).abs().max() / g64.abs().max()),
                         "ref32_vs_64_total": abs(d_ref["total"] - d64["total"])}
        print(name, summary[name], flush=True)
    # batched input [B,H,W,3] (losses.py:71-74)
    a, b = make_pair(33, 40, 9, "noise")
    a2, b2 = make_pair(33, 40, 10, "smooth")
    pb, tb = torch.stack([a, a2]), torch.stack([b, b2])
    assert float(ref_losses.ssim_loss(pb, tb)) == float(L.ssim_loss(pb, tb))
    tot, d = ref_losses.compute_loss(pb, tb)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "loss_batch2_40x33.npz"), pred=pb.numpy(), target=tb.numpy(),
                        l1=np.float32(d["l1"]), ssim=np.float32(d["ssim"]), total=np.float32(d["total"]))
    summary["loss_batch2_40x33"] = d
    json.dump(summary, open(os.path.join(ROOT, "tests", "golden", "LOSS_SUMMARY.json"), "w"), indent=1)
    # the 1-D window the kernels hard-code (fp32, as torch computes it)
    print("window:", [float(v) for v in L.gaussian_window_1d()])
    print(["%a" % float(v) for v in L.gaussian_window_1d()])


if __name__ == "__main__":
    main()
