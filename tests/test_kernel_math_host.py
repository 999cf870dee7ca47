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


def test_adam_arithmetic_without_ieee_slow_paths_is_the_plain_arithmetic():
    """csrc/gs_math.cuh: sqrt_no_slow_path / div_no_slow_path rescale zero and tiny operands by exact powers of two so
    that the GPU never branches to the IEEE slow paths; for every input whose result is a normal number they must give
    the plain, correctly rounded sqrt / quotient bit for bit (g++ evaluates both with IEEE arithmetic)."""
    import ctypes
    import hostcheck
    lib = hostcheck.load()
    rng = np.random.default_rng(5)
    n = 400_000
    mag = 10.0 ** rng.uniform(-44, 6, n)                       # from the denormal range up
    a = (mag * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    a[:1000] = 0.0
    a[1000:2000] = np.float32(1e-45)                            # the smallest denormal
    b = (10.0 ** rng.uniform(-15, 3, n)).astype(np.float32)     # denominators: eps = 1e-15 and up
    out = [np.zeros(n, np.float32) for _ in range(4)]
    fp = ctypes.POINTER(ctypes.c_float)
    lib.hc_sqrt_div(ctypes.c_int(n), a.ctypes.data_as(fp), b.ctypes.data_as(fp), *[o.ctypes.data_as(fp) for o in out])
    s_r, s_p, d_r, d_p = out
    assert np.array_equal(s_r.view(np.uint32), s_p.view(np.uint32))            # sqrt of a denormal is a normal number
    tiny = np.float32(1.17549435e-38)
    normal = (np.abs(d_p) >= tiny) | (d_p == 0)
    assert normal.mean() > 0.6 and np.array_equal(d_r[normal].view(np.uint32) & 0x7FFFFFFF, d_p[normal].view(np.uint32) & 0x7FFFFFFF)
    assert np.all(np.abs(d_r[~normal] - d_p[~normal]) <= 1.5e-45 * 2)          # denormal quotients: one rounding step apart at most
    # one full Adam update against the formula of torch/optim/adam.py evaluated in float32 with numpy
    g = (rng.standard_normal(n) * 10.0 ** rng.uniform(-24, 0, n)).astype(np.float32)
    g[:5000] = 0.0
    p = rng.standard_normal(n).astype(np.float32)
    m = (g * np.float32(0.3)).astype(np.float32)
    v = (g * g * np.float32(0.01)).astype(np.float32)
    b1, b2, eps, step = np.float32(0.9), np.float32(0.999), np.float32(1e-15), 7
    step_size = np.float32(1e-2 / (1.0 - 0.9 ** step))
    bc2 = np.float32(np.sqrt(1.0 - 0.999 ** step))
    m_ref = m + (g - m) * np.float32(1.0 - 0.9)
    v_ref = (np.float32(1.0 - 0.999) * g).astype(np.float32).astype(np.float64) * g.astype(np.float64) + (v * b2).astype(np.float64)
    v_ref = v_ref.astype(np.float32)                                            # one rounding: the kernel uses an fma
    denom = np.sqrt(v_ref) / bc2 + eps
    p_ref = p - step_size * (m_ref / denom)
    pp, mm, vv = p.copy(), m.copy(), v.copy()
    lib.hc_adam(ctypes.c_int(n), pp.ctypes.data_as(fp), g.ctypes.data_as(fp), mm.ctypes.data_as(fp), vv.ctypes.data_as(fp),
                ctypes.c_double(0.9), ctypes.c_double(0.999), ctypes.c_double(1e-15), ctypes.c_float(float(step_size)),
                ctypes.c_float(float(bc2)))
    assert np.array_equal(mm, m_ref.astype(np.float32)) and np.array_equal(vv, v_ref)
    assert np.array_equal(pp, p_ref.astype(np.float32))
    assert np.array_equal(pp[:5000], p[:5000])                                  # g = m = v = 0: untouched
