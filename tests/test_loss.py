"""The training loss (L1 + SSIM; gaussian_splatting/losses.py): oracle vs the golden vectors generated from
the unmodified reference (CPU), and the fused CUDA kernels vs both (GPU, through the C ABI behind
b200gs.compute_loss)."""
import glob
import os

import numpy as np
import pytest
import torch

from common import ROOT

GOLDEN = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(ROOT, "tests", "golden", "loss_*.npz")))
SINGLE = [g for g in GOLDEN if "batch" not in g]
VAL_TOL = 2e-6       # absolute, on loss values in [0, 1]
GRAD_TOL = 1e-3      # max-norm relative, the bar SURVEY.md section 8c sets for gradients


def _load(name):
    return {k: v for k, v in np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")).items()}


def test_golden_set_is_present():
    assert len(SINGLE) >= 5 and "loss_batch2_40x33" in GOLDEN


@pytest.mark.parametrize("name", SINGLE)
def test_oracle_reproduces_reference_loss_and_gradient(name):
    from oracle import loss_oracle as L
    G = _load(name)
    pred = torch.from_numpy(G["pred"]).requires_grad_(True)
    total, d = L.compute_loss(pred, torch.from_numpy(G["target"]))
    total.backward()
    assert np.float32(d["l1"]) == G["l1"] and np.float32(d["ssim"]) == G["ssim"] and np.float32(d["total"]) == G["total"]
    assert np.array_equal(pred.grad.numpy(), G["grad"])
    t2, _ = L.compute_loss(pred.detach(), torch.from_numpy(G["target"]), 0.3, 0.7)
    assert np.float32(float(t2)) == G["total_03_07"]


def test_oracle_batched_input():
    from oracle import loss_oracle as L
    G = _load("loss_batch2_40x33")
    _, d = L.compute_loss(torch.from_numpy(G["pred"]), torch.from_numpy(G["target"]))
    assert np.float32(d["l1"]) == G["l1"] and np.float32(d["ssim"]) == G["ssim"]


def _relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("name", SINGLE)
def test_cuda_loss_against_golden(name):
    import b200gs
    G = _load(name)
    pred = torch.from_numpy(G["pred"]).cuda().requires_grad_(True)
    target = torch.from_numpy(G["target"]).cuda()
    total, d = b200gs.compute_loss(pred, target)
    assert total.dim() == 0 and total.requires_grad and isinstance(d["l1"], float)
    assert abs(d["l1"] - float(G["l1_64"])) <= VAL_TOL and abs(d["ssim"] - float(G["ssim_64"])) <= VAL_TOL
    assert abs(d["total"] - float(G["total_64"])) <= VAL_TOL and abs(float(total) - d["total"]) == 0.0
    (total * 3.0).backward()                                           # upstream gradient != 1
    g = pred.grad.cpu().numpy() / 3.0
    # vs the reference's fp32 autograd and vs the fp64 arbiter
    assert _relerr(g, G["grad64"]) <= max(GRAD_TOL * 0.1, 3 * _relerr(G["grad"], G["grad64"])), name
    assert _relerr(g, G["grad"]) <= GRAD_TOL
    # other weights, and the two single-term entry points
    t2, _ = b200gs.compute_loss(pred.detach(), target, lambda_l1=0.3, lambda_ssim=0.7)
    assert abs(float(t2) - float(G["total_03_07"])) <= VAL_TOL
    assert abs(float(b200gs.l1_loss(pred.detach(), target)) - float(G["l1"])) <= VAL_TOL
    assert abs(float(b200gs.ssim_loss(pred.detach(), target)) - float(G["ssim"])) <= VAL_TOL
    with pytest.raises(NotImplementedError):
        b200gs.ssim_loss(pred.detach(), target, window_size=7)


@pytest.mark.gpu
def test_cuda_loss_batched_and_live_oracle():
    import b200gs
    from oracle import loss_oracle as L
    G = _load("loss_batch2_40x33")
    _, d = b200gs.compute_loss(torch.from_numpy(G["pred"]).cuda(), torch.from_numpy(G["target"]).cuda())
    assert abs(d["l1"] - float(G["l1"])) <= VAL_TOL and abs(d["ssim"] - float(G["ssim"])) <= VAL_TOL
    # a render-sized frame with partial tiles on both axes (1297 = 40*32 + 17, 840 = 26*32 + 8), live oracle
    g = torch.Generator().manual_seed(11)
    H, W = 840, 1297
    a = torch.rand(H // 8 + 1, W // 8 + 1, 3, generator=g)
    up = lambda t: torch.nn.functional.interpolate(t.permute(2, 0, 1)[None], size=(H, W), mode="bilinear")[0].permute(1, 2, 0)
    pred = up(a).contiguous()
    target = (0.9 * pred + 0.1 * torch.rand(H, W, 3, generator=g)).contiguous()
    p_ref = pred.clone().requires_grad_(True)
    t_ref, d_ref = L.compute_loss(p_ref, target)
    t_ref.backward()
    p = pred.cuda().requires_grad_(True)
    total, d = b200gs.compute_loss(p, target.cuda())
    total.backward()
    assert abs(d["l1"] - d_ref["l1"]) <= VAL_TOL and abs(d["ssim"] - d_ref["ssim"]) <= 2 * VAL_TOL
    assert _relerr(p.grad.cpu().numpy(), p_ref.grad.numpy()) <= GRAD_TOL
    # size-independent properties at the headline resolution: identical images -> l1 = 0, ssim loss = 0, and the
    # gradient of the L1 term vanishes (sign(0) = 0); the loss is symmetric in (pred, target)
    x = torch.rand(1080, 1920, 3, device="cuda")
    y = torch.rand(1080, 1920, 3, device="cuda")
    _, same = b200gs.compute_loss(x, x.clone())
    assert same["l1"] == 0.0 and abs(same["ssim"]) <= 1e-6
    _, ab = b200gs.compute_loss(x, y)
    _, ba = b200gs.compute_loss(y, x)
    assert abs(ab["l1"] - ba["l1"]) <= 1e-7 and abs(ab["ssim"] - ba["ssim"]) <= 1e-6
    xg = x.clone().requires_grad_(True)
    b200gs.l1_loss(xg, x.clone()).backward()
    assert float(xg.grad.abs().max()) == 0.0
