"""Fused Adam step + gradient-norm clipping (scripts/train.py:394-401,536-538) against torch's own
implementations - the reference's optimizer IS torch.optim.Adam (third-party dependency of the reference,
torch >= 1.9 per requirements.txt:1; pinned here by the torch build in the image), so torch is the oracle."""
import pytest
import torch


def _groups(n, dev, seed):
    g = torch.Generator().manual_seed(seed)
    shapes = dict(pos=(n, 3), opacity_raw=(n,), f_dc=(n, 3), f_rest=(n, 45), scale_raw=(n, 3), q_raw=(n, 4))
    lrs = dict(pos=1.6e-4, opacity_raw=0.05, f_dc=2.5e-3, f_rest=2.5e-3 / 20, scale_raw=5e-3, q_raw=1e-3)
    params = {k: torch.randn(*s, generator=g).to(dev) for k, s in shapes.items()}
    return params, lrs, g


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 1023, 4097, 100_003])
def test_fused_adam_matches_torch_adam(n):
    import b200gs
    dev = torch.device("cuda")
    params, lrs, g = _groups(n, dev, seed=n)
    mine = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    mk = lambda ps: [{"params": [ps[k]], "lr": lrs[k], "name": k} for k in ps]          # train.py:394-401
    opt_m = b200gs.FusedAdam(mk(mine), lr=1e-3, eps=1e-15)
    opt_r = torch.optim.Adam(mk(ref), lr=1e-3, eps=1e-15, foreach=False)
    for it in range(6):
        opt_m.param_groups[0]["lr"] = opt_r.param_groups[0]["lr"] = 1.6e-4 * 0.9 ** it      # the pos-LR schedule (:445-457)
        for k in params:
            gr = (torch.randn(params[k].shape, generator=g) * (10.0 ** (it - 3))).to(dev)    # wide dynamic range
            if it == 4 and k == "f_rest":
                gr.zero_()
            mine[k].grad, ref[k].grad = gr.clone(), gr.clone()
        if it == 2:
            mine["q_raw"].grad = ref["q_raw"].grad = None                                 # a tensor without gradient is skipped
        opt_m.step()
        opt_r.step()
    for k in params:
        a, b = mine[k].detach(), ref[k].detach()
        assert float((a - b).abs().max() / b.abs().max()) <= 2e-6, k
        sm, sr = opt_m.state[mine[k]], opt_r.state[ref[k]]
        assert float(sm["step"]) == float(sr["step"])
        assert float((sm["exp_avg"] - sr["exp_avg"]).abs().max() / sr["exp_avg"].abs().max().clamp_min(1e-30)) <= 2e-6
        assert float((sm["exp_avg_sq"] - sr["exp_avg_sq"]).abs().max() / sr["exp_avg_sq"].abs().max().clamp_min(1e-30)) <= 2e-6
    # state dicts are interchangeable with torch.optim.Adam
    opt_r2 = torch.optim.Adam(mk(ref), lr=1e-3, eps=1e-15)
    opt_r2.load_state_dict(opt_m.state_dict())
    with pytest.raises(NotImplementedError):
        b200gs.FusedAdam(mk(mine), amsgrad=True)
    cpu = torch.zeros(4, requires_grad=True)
    cpu.grad = torch.ones(4)
    with pytest.raises(b200gs.B200GSError):
        b200gs.FusedAdam([cpu]).step()


@pytest.mark.gpu
def test_fused_adam_zero_and_denormal_range_gradients_match_torch():
    """Culled Gaussians have g = m = v = 0 and barely visible ones g*g in the denormal range: the kernel rescales
    such operands around sqrt / division instead of taking the slow IEEE paths (csrc/common.cuh) - the results
    must still be torch's."""
    import b200gs
    dev = torch.device("cuda")
    n = 40_000
    g = torch.Generator().manual_seed(7)
    p0 = torch.randn(n, 3, generator=g)
    mags = torch.tensor([0.0, 1e-30, 1e-24, 1e-21, 1e-19, 1e-12, 1e-6, 1.0]).repeat_interleave(n // 8)
    mine = p0.clone().to(dev).requires_grad_(True)
    ref = p0.clone().to(dev).requires_grad_(True)
    opt_m = b200gs.FusedAdam([mine], lr=1e-2, eps=1e-15)
    opt_r = torch.optim.Adam([ref], lr=1e-2, eps=1e-15, foreach=False)
    for it in range(5):
        gr = (torch.randn(n, 3, generator=g) * mags[:, None]).to(dev)
        mine.grad, ref.grad = gr.clone(), gr.clone()
        opt_m.step()
        opt_r.step()
    sm, sr = opt_m.state[mine], opt_r.state[ref]
    k = n // 8
    for gi in range(8):                                    # every magnitude group on its own scale
        sl = slice(gi * k, (gi + 1) * k)
        for name in ("exp_avg", "exp_avg_sq"):
            a, b = sm[name][sl], sr[name][sl]
            tol = 2e-6 * float(b.abs().max()) + 3e-45      # + two denormal steps where the moments underflow
            assert float((a - b).abs().max()) <= tol, (gi, name, float((a - b).abs().max()), tol)
        a, b = mine.detach()[sl], ref.detach()[sl]
        # the UPDATE (not just the parameter) must agree: compare the displacement from the start
        da, db = a - p0[sl].to(dev), b - p0[sl].to(dev)
        assert float((da - db).abs().max()) <= 2e-6 * float(db.abs().max()) + 2.5e-7, (gi, float((da - db).abs().max()))
    assert torch.equal(mine.detach()[:k], p0[:k].to(dev))            # zero gradient, zero moments: untouched


@pytest.mark.gpu
@pytest.mark.parametrize("n,scale", [(1, 5.0), (4096, 1e-3), (4097, 1.0), (1_000_003, 0.01), (300_000, 3.0)])
def test_clip_grad_norm_matches_torch(n, scale):
    import b200gs
    g = torch.Generator().manual_seed(n)
    grad = (torch.randn(n, 3, generator=g) * scale).cuda()
    a = torch.zeros(n, 3, device="cuda", requires_grad=True)
    b = torch.zeros(n, 3, device="cuda", requires_grad=True)
    a.grad, b.grad = grad.clone(), grad.clone()
    tn_a = b200gs.clip_grad_norm_(a, max_norm=1.0)                       # train.py:536
    tn_b = torch.nn.utils.clip_grad_norm_(b, max_norm=1.0)
    assert tn_a.is_cuda and abs(float(tn_a) - float(tn_b)) <= 2e-6 * float(tn_b)
    if float(tn_b) <= 1.0 - 1e-5:
        assert torch.equal(a.grad, grad)                                  # below the threshold: untouched
    assert float((a.grad - b.grad).abs().max()) <= 2e-6 * float(b.grad.abs().max())
    # several tensors: joint norm
    c = torch.zeros(7, device="cuda", requires_grad=True)
    d = torch.zeros(7, device="cuda", requires_grad=True)
    c.grad, d.grad = torch.full((7,), 2.0, device="cuda"), torch.full((7,), 2.0, device="cuda")
    t1 = b200gs.clip_grad_norm_([a, c], 0.5)
    t2 = torch.nn.utils.clip_grad_norm_([b, d], 0.5)
    assert abs(float(t1) - float(t2)) <= 1e-5 * float(t2) and float((c.grad - d.grad).abs().max()) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("max_norm", [0.25, 1e6])
def test_clip_grad_norm_of_the_six_parameter_tensors_matches_torch(max_norm):
    """torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm): one joint norm, one coefficient, every tensor
    scaled - through the table-driven kernels (no torch ops on the way), incl. an empty tensor and odd sizes."""
    import b200gs
    g = torch.Generator().manual_seed(11)
    n = 70_001
    shapes = [(n, 3), (n,), (n, 3), (n, 45), (n, 3), (n, 4), (0, 3), (5,)]
    mine = [torch.zeros(s, device="cuda", requires_grad=True) for s in shapes]
    ref = [torch.zeros(s, device="cuda", requires_grad=True) for s in shapes]
    for a, b in zip(mine, ref):
        gr = (torch.randn(a.shape, generator=g) * 1e-3).cuda()
        a.grad, b.grad = gr.clone(), gr.clone()
    before = [a.grad.clone() for a in mine]
    t_mine = b200gs.clip_grad_norm_(mine, max_norm)
    t_ref = torch.nn.utils.clip_grad_norm_(ref, max_norm)
    assert t_mine.is_cuda and abs(float(t_mine) - float(t_ref)) <= 2e-6 * float(t_ref)
    for a, b, g0 in zip(mine, ref, before):
        if a.numel() == 0:
            continue
        if max_norm > float(t_ref):
            assert torch.equal(a.grad, g0)                               # below the threshold: bit-for-bit untouched
        assert float((a.grad - b.grad).abs().max()) <= 2e-6 * float(b.grad.abs().max())
    with pytest.raises(NotImplementedError):
        b200gs.clip_grad_norm_([_with_grad(3) for _ in range(17)], 1.0)


def _with_grad(n):
    p = torch.zeros(n, device="cuda", requires_grad=True)
    p.grad = torch.ones(n, device="cuda")
    return p


def test_install_can_swap_the_optimizer(monkeypatch):
    import sys
    import types
    import b200gs
    pkg = types.ModuleType("fake_gs2")
    pkg.__path__ = []
    for sub, attrs in (("render", ("render",)), ("gaussian", ("build_sigma_from_params",)),
                       ("spherical_harmonics", ("evaluate_sh",)), ("losses", ("compute_loss", "l1_loss", "ssim_loss"))):
        m = types.ModuleType(f"fake_gs2.{sub}")
        for attr in attrs:
            setattr(m, attr, lambda *a, **k: "reference")
        monkeypatch.setitem(sys.modules, f"fake_gs2.{sub}", m)
    monkeypatch.setitem(sys.modules, "fake_gs2", pkg)
    adam, clip = torch.optim.Adam, torch.nn.utils.clip_grad_norm_
    b200gs.install("fake_gs2", optimizer=True)
    try:
        assert torch.optim.Adam is b200gs.FusedAdam and torch.nn.utils.clip_grad_norm_ is b200gs.clip_grad_norm_
    finally:
        b200gs.uninstall()
    assert torch.optim.Adam is adam and torch.nn.utils.clip_grad_norm_ is clip
