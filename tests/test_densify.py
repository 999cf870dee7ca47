"""Adaptive density control (b200gs.densify) against tests/golden/densify_*.npz, which were produced by
the REFERENCE's own GaussianModel.densify_and_prune (scripts/train.py:89-195) on the CPU with the torch.randn_like draw
recorded (oracle/make_golden_densify.py).  Copied rows must be bit-exact; displaced positions agree to 1e-6 (CPU vs GPU
expf); the golden inputs keep every threshold decision away from its threshold."""
import os
import types

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARAMS = ("pos", "opacity_raw", "f_dc", "f_rest", "scale_raw", "q_raw")
CASES = ("split", "clone", "prune_only", "cold", "both")


def _load(name):
    return np.load(os.path.join(GOLD, f"densify_{name}.npz"))


def test_golden_set_is_present_and_consistent():
    for name in CASES:
        z = _load(name)
        n_in, n_out = z["in_pos"].shape[0], z["out_pos"].shape[0]
        assert z["in_f_rest"].shape == (n_in, 45) and z["out_q_raw"].shape == (n_out, 4)
    assert str(_load("both")["raised"]) != "" and str(_load("split")["raised"]) == ""


def test_densify_has_no_cpu_fallback():
    import b200gs
    z = _load("cold")
    with pytest.raises(b200gs.B200GSError):
        b200gs.densify_tensors({k: torch.from_numpy(z["in_" + k]) for k in PARAMS}, None)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["split", "clone", "prune_only", "cold"])
def test_densify_matches_the_reference_method(name):
    import b200gs
    z = _load(name)
    dev = torch.device("cuda")
    params = {k: torch.from_numpy(z["in_" + k]).to(dev) for k in PARAMS}
    has_grads = "gin_pos" in z.files
    pos_grad = torch.from_numpy(z["gin_pos"]).to(dev) if has_grads else None
    thr_op, max_grad, thr_scale = (float(x) for x in z["kw"])
    noise = torch.from_numpy(z["noise"]).to(dev) if z["noise"].shape[0] else None
    out, info = b200gs.densify_tensors(params, pos_grad, thr_op, max_grad, thr_scale, noise=noise)
    n_out = z["out_pos"].shape[0]
    assert info["n_keep"] + info["n_split"] + info["n_clone"] == n_out and info["n_split"] == z["noise"].shape[0]
    nk, ns = info["n_keep"], info["n_split"]
    for k in PARAMS:
        got, want = out[k].cpu().numpy(), z["out_" + k]
        assert got.shape == want.shape, k
        if k == "pos":
            assert np.array_equal(got[:nk], want[:nk]) and np.array_equal(got[nk + ns:], want[nk + ns:])
            assert np.abs(got[nk:nk + ns] - want[nk:nk + ns]).max(initial=0.0) <= 1e-6
        else:
            assert np.array_equal(got, want), k                      # copies and scale_raw - 0.5: bit-exact
    if has_grads:                                                     # the grads dict is pruned like train.py:122-126
        keep = info["keep"].cpu().numpy()
        assert np.array_equal(z["gin_pos"][keep], z["gout_pos"]) and np.array_equal(z["gin_opacity_raw"][keep], z["gout_opacity_raw"])


@pytest.mark.gpu
def test_split_and_clone_in_one_call_raises_like_the_reference_unless_told_otherwise():
    import b200gs
    z = _load("both")
    dev = torch.device("cuda")
    params = {k: torch.from_numpy(z["in_" + k]).to(dev) for k in PARAMS}
    pos_grad = torch.from_numpy(z["gin_pos"]).to(dev)
    with pytest.raises(IndexError):
        b200gs.densify_tensors(params, pos_grad)
    out, info = b200gs.densify_tensors(params, pos_grad, strict=False, noise=torch.from_numpy(z["noise"]).to(dev))
    # the reference had already applied the split when it raised: its state = kept rows + split copies
    nk, ns, nc = info["n_keep"], info["n_split"], info["n_clone"]
    assert ns == z["noise"].shape[0] and nc > 0 and z["out_pos"].shape[0] == nk + ns
    for k in PARAMS:
        got, want = out[k].cpu().numpy(), z["out_" + k]
        if k == "pos":
            assert np.array_equal(got[:nk], want[:nk]) and np.abs(got[nk:nk + ns] - want[nk:]).max() <= 1e-6
        else:
            assert np.array_equal(got[:nk + ns], want), k
    # the clones are identical copies of kept rows that were hot and small
    sm = np.exp(z["in_scale_raw"]).max(-1)
    hot = np.linalg.norm(z["gin_pos"], axis=-1) > 0.01
    keep = info["keep"].cpu().numpy()
    src = np.nonzero(keep & hot & (sm <= 0.01))[0]
    assert len(src) == nc and np.array_equal(out["f_rest"].cpu().numpy()[nk + ns:], z["in_f_rest"][src])


@pytest.mark.gpu
def test_model_level_call_replaces_parameters_and_draws_the_noise_like_the_reference():
    import b200gs
    z = _load("split")
    dev = torch.device("cuda")
    model = types.SimpleNamespace(**{k: torch.nn.Parameter(torch.from_numpy(z["in_" + k]).to(dev)) for k in PARAMS})
    grads = {"pos": torch.from_numpy(z["gin_pos"]).to(dev), "opacity_raw": torch.from_numpy(z["gin_opacity_raw"]).to(dev)}
    torch.manual_seed(123)
    info = b200gs.densify_and_prune(model, grads)
    after = torch.cuda.get_rng_state(dev)
    torch.manual_seed(123)
    noise = torch.randn((info["n_split"], 3), device=dev)            # what _split_points draws (train.py:164)
    assert torch.equal(after, torch.cuda.get_rng_state(dev))          # the same generator state afterwards
    assert isinstance(model.pos, torch.nn.Parameter) and model.pos.shape[0] == z["out_pos"].shape[0]
    assert grads["pos"].shape[0] == info["n_keep"]
    nk, ns = info["n_keep"], info["n_split"]
    src = torch.from_numpy(np.nonzero(info["keep"].cpu().numpy())[0]).to(dev)
    kept_pos, kept_scale = model.pos.detach()[:nk], model.scale_raw.detach()[:nk]
    assert torch.equal(kept_pos, torch.from_numpy(z["in_pos"]).to(dev)[src])
    hot = (grads["pos"].norm(dim=-1) > 0.01) & (torch.exp(kept_scale).max(dim=-1)[0] > 0.01)
    want = kept_pos[hot] + noise * torch.exp(kept_scale[hot]) * 0.1
    assert float((model.pos.detach()[nk:nk + ns] - want).abs().max()) <= 1e-6
