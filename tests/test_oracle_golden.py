"""CPU: the oracle restatement against the committed golden vectors (generated from the unmodified
reference by oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from common import GOLDEN_CASES, GRAD_CASES, PARAMS, golden_inputs, load_golden
from oracle import gs_oracle as O


@pytest.mark.parametrize("name", ["sh3_4k_200x136_rot", "edge_1500_97x71", "dense_600_48x40"])
def test_oracle_reproduces_reference_image_and_stages(name):
    G = load_golden(name)
    sc, cam = golden_inputs(G)
    sig = O.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    col = O.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], cam["c2w"])
    img, proj, bins = O.render(sc["pos"], col, sc["opacity_raw"], sig, cam["c2w"], cam["H"], cam["W"], cam["fx"],
                               cam["fy"], cam["cx"], cam["cy"], return_stages=True)
    # torch CPU kernels are deterministic for a given build; allow 1e-6 for other builds
    assert np.abs(img.numpy() - G["image"]).max() <= 1e-6
    assert np.array_equal(proj.ids.numpy(), G["ids"])
    assert np.array_equal(proj.radius.numpy(), G["radius"])
    assert np.array_equal(proj.rect.numpy(), G["rect"])
    assert np.array_equal(proj.tiles_touched.numpy(), G["tiles_touched"])
    assert np.array_equal(proj.ids[bins.ranks].numpy(), G["list_id"])
    assert np.array_equal(bins.start.numpy(), G["start"]) and np.array_equal(bins.end.numpy(), G["end"])


def test_oracle_gradients_match_reference_autograd():
    name = "dense_600_48x40"
    G = load_golden(name)
    sc, cam = golden_inputs(G)
    leaves = {k: sc[k].clone().requires_grad_(True) for k in PARAMS}
    img = O.render_from_params(leaves["pos"], leaves["scale_raw"], leaves["q_raw"], leaves["opacity_raw"],
                               leaves["f_dc"], leaves["f_rest"], cam["c2w"], cam["H"], cam["W"], cam["fx"], cam["fy"],
                               cam["cx"], cam["cy"])
    grads = torch.autograd.grad((img * torch.from_numpy(G["loss_w"])).sum(), [leaves[k] for k in PARAMS])
    for k, g in zip(PARAMS, grads):
        ref = G["grad_" + k]
        assert np.abs(g.numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


def test_oracle_edge_behaviours():
    sc = O.make_scene(50, seed=2)
    cam = O.make_camera(64, 64)
    sig = O.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    col = O.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], cam["c2w"])
    img = O.render(sc["pos"], col, torch.full((50,), -20.0), sig, cam["c2w"], 64, 64, 57.6, 57.6, 32., 32.)
    assert img.shape == (64, 64, 3) and float(img.abs().max()) == 0.0     # render.py:109-112
    G = load_golden("offscreen_case")
    sig2 = O.build_sigma_from_params(torch.from_numpy(G["scale_raw"]), torch.from_numpy(G["q_raw"]))
    with pytest.raises(Exception, match="off-screen"):                      # render.py:235-236
        O.render(torch.from_numpy(G["pos"]), col, torch.full((50,), 2.0), sig2, torch.from_numpy(G["c2w"]), 64, 64,
                 57.6, 57.6, 32., 32.)
    # H, W as 0-dim tensors (scripts/train.py:499)
    img2 = O.render(sc["pos"], col, sc["opacity_raw"], sig, cam["c2w"], torch.tensor(64), torch.tensor(64), 57.6, 57.6,
                    32., 32.)
    assert img2.shape == (64, 64, 3)


def test_scene_generator_is_seeded():
    a, b = O.make_scene(100, seed=5), O.make_scene(100, seed=5)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert float(O.make_scene(10, sh_degree=0)["f_rest"].abs().max()) == 0.0
