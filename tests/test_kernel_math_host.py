"""CPU: the arithmetic the CUDA kernels run (csrc/gs_math.cuh compiled for the host by tests/hostcheck)
against the golden vectors - projection, integer stages, SH, and the whole backward chain (with the
blend-backward formulas mirrored in tests/blend_mirror.py)."""
import numpy as np
import pytest

import hostcheck as hc
from blend_mirror import blend_and_grads
from common import GOLDEN_CASES, GRAD_CASES, PARAMS, grad_relerr, image_report, load_golden


def _project(G):
    H, W, fx, fy, cx, cy = G["cam"]
    cam = (int(H), int(W), fx, fy, cx, cy)
    P = hc.project(G["in_pos"], G["in_opacity_raw"], G["c2w"], cam, scale_raw=G["in_scale_raw"], q_raw=G["in_q_raw"],
                   f_dc=G["in_f_dc"], f_rest=G["in_f_rest"])
    return P, cam


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_projection_and_integer_stages(name):
    G = load_golden(name)
    P, cam = _project(G)
    ids = G["ids"]
    vis_ref = np.zeros(P["vis"].shape[0], bool)
    vis_ref[ids] = True
    assert np.array_equal(vis_ref, P["vis"] > 0)                 # same survivors
    assert np.abs(P["u"][ids] - G["u"]).max() <= 1e-4
    assert np.abs(P["v"][ids] - G["v"]).max() <= 1e-4
    assert np.array_equal(P["z"][ids], G["z"])                   # depth: bit-exact (sort keys)
    assert np.array_equal(P["radius"][ids], G["radius"])         # bit-exact integer stages
    assert np.array_equal(P["rect"][ids], G["rect"])
    assert np.array_equal(P["tiles"][ids], G["tiles_touched"])
    assert np.abs(P["op"][ids] - G["opacity"]).max() <= 2e-7
    assert np.abs(P["rgb"][ids] - G["color"][ids]).max() <= 5e-7


@pytest.mark.parametrize("name", GRAD_CASES)
def test_backward_chain(name):
    G = load_golden(name)
    P, cam = _project(G)
    H, W = cam[0], cam[1]
    img, sg = blend_and_grads(H, W, (W + 15) // 16, G["uniq_tiles"], G["start"], G["end"], G["list_id"], P["u"], P["v"],
                              P["conic"], P["op"], P["rgb"], G["loss_w"])
    rep = image_report(img, G["image"], G["image64"])
    assert rep["n_bad_min"] <= 2 * rep["n_bad_ref32_vs_ref64"] + 2, rep
    B = hc.backward(G["in_pos"], G["in_scale_raw"], G["in_q_raw"], G["in_opacity_raw"], G["in_f_dc"], G["in_f_rest"],
                    G["c2w"], cam, sg)
    for k in PARAMS:
        e32 = grad_relerr(B[k], G["grad_" + k])
        e64 = grad_relerr(B[k], G["grad64_" + k])
        noise = grad_relerr(G["grad_" + k], G["grad64_" + k])      # the reference's own fp32 round-off
        assert min(e32, e64) <= max(1e-3, 2 * noise), (k, e32, e64, noise)
