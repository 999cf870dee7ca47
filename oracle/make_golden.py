"""Pin the oracle against the UNMODIFIED reference and write tests/golden/*.npz.

Run in the build container only (needs /root/reference; the GPU box has no reference checkout):

    python oracle/make_golden.py

For every case below it
  1. renders with the reference's own `build_sigma_from_params` / `evaluate_sh` / `render`
     (imported from BASELINE.json's reference_path, CPU, fp32),
  2. asserts the oracle restatement (`oracle/gs_oracle.py`) gives the same image bit-for-bit
     and the same autograd gradients for the six parameter tensors,
  3. stores inputs, the reference image, the reference gradients (loss = sum(img * w) with a
     fixed random w) and the oracle's intermediate integer stages (survivor ids in depth order,
     radii, tile rects, tile counts, per-tile sorted lists, ranges) as a compressed fixture.

TEST INFRASTRUCTURE ONLY (see oracle/gs_oracle.py header).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = json.load(open(os.path.join(ROOT, "BASELINE.json")))["reference_path"]
sys.path.insert(0, REF)

from oracle import gs_oracle as O  # noqa: E402
import gaussian_splatting  # noqa: E402,F401
# NB: the package __init__ rebinds the attribute `gaussian_splatting.render` to the FUNCTION, so the
# modules have to be fetched from sys.modules.
ref_gaussian = sys.modules["gaussian_splatting.gaussian"]
ref_render = sys.modules["gaussian_splatting.render"]
ref_sh = sys.modules["gaussian_splatting.spherical_harmonics"]

PARAMS = ("pos", "scale_raw", "q_raw", "opacity_raw", "f_dc", "f_rest")


def edge_scene(n=1500, seed=7):
    """Adversarial mix: hits every cull / clamp branch of render.py S1-S15."""
    sc = O.make_scene(n, seed=seed, log_scale=-3.0, sh_degree=3)
    g = torch.Generator().manual_seed(seed + 100)
    k = n // 10
    sc["pos"][0:k, 2] -= 4.5                        # behind / very near the camera
    sc["opacity_raw"][k:2 * k] = -7.0                 # below the 1/256 pre-cull
    sc["scale_raw"][2 * k:3 * k] = -0.2               # huge splats: lambda_max clamp, radius 250
    sc["pos"][2 * k:3 * k, 2] = -2.4 + 0.01 * torch.rand(k, generator=g)
    sc["scale_raw"][2 * k:2 * k + 12] = 1.5           # giants near the origin: lambda_max > 1e4 -> clamp, r = 250
    sc["pos"][2 * k:2 * k + 12] *= 0.3
    sc["opacity_raw"][2 * k:2 * k + 12] = -3.0
    sc["scale_raw"][3 * k:4 * k] = -11.0              # needles: lambda clamp at 1e-6 (exp(-11)*fx/z)^2 ~ 1e-7
    sc["scale_raw"][4 * k:5 * k, 0] = -1.0            # long thin: anisotropic
    sc["scale_raw"][4 * k:5 * k, 1:] = -7.5
    sc["pos"][5 * k:6 * k, 0] += 2.2                  # off to the side: guard band / off-screen
    sc["opacity_raw"][6 * k:7 * k] = 9.0              # sigmoid > 0.999 -> opacity clamp, alpha clamp at 0.99
    return sc


CASES = {
    # name: (scene fn, camera kwargs, with_grad)
    "c1_10k_sh0_256": (lambda: O.make_scene(10000, seed=0, log_scale=-3.5, sh_degree=0),
                       dict(W=256, H=256, view=0, n_views=1), False),
    "sh3_4k_200x136_rot": (lambda: O.make_scene(4000, seed=1, log_scale=-3.2, sh_degree=3),
                           dict(W=200, H=136, view=1, n_views=8), True),
    "edge_1500_97x71": (edge_scene, dict(W=97, H=71, view=3, n_views=16), True),
    "dense_600_48x40": (lambda: O.make_scene(600, seed=3, log_scale=-2.0, sh_degree=3),
                        dict(W=48, H=40, view=5, n_views=8), True),
}


def run_case(name, scene_fn, cam_kw, with_grad):
    sc = scene_fn()
    cam = O.make_camera(**cam_kw)
    c2w = cam["c2w"]
    H, W, fx, fy, cx, cy = cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    leaves = {k: sc[k].clone().requires_grad_(with_grad) for k in PARAMS}

    # --- reference ---------------------------------------------------------------------------
    sig = ref_gaussian.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
    col = ref_sh.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
    img_ref = ref_render.render(leaves["pos"], col, leaves["opacity_raw"], sig, c2w, H, W, fx, fy, cx, cy)
    out = {f"in_{k}": sc[k].numpy() for k in PARAMS}
    out.update(c2w=c2w.numpy(), cam=np.array([H, W, fx, fy, cx, cy], dtype=np.float64),
               image=img_ref.detach().numpy(), sigma=sig.detach().numpy(), color=col.detach().numpy())
    gw = torch.Generator().manual_seed(1234)
    wimg = torch.rand(H, W, 3, generator=gw)
    if with_grad:
        grads_ref = torch.autograd.grad((img_ref * wimg).sum(), [leaves[k] for k in PARAMS])
        for k, gr in zip(PARAMS, grads_ref):
            out[f"grad_{k}"] = gr.numpy()
        out["loss_w"] = wimg.numpy()

    # --- oracle must agree with the reference ----------------------------------------------------
    l2 = {k: sc[k].clone().requires_grad_(with_grad) for k in PARAMS}
    sig_o = O.build_sigma_from_params(l2["scale_raw"], l2["q_raw"])
    col_o = O.evaluate_sh(l2["f_dc"], l2["f_rest"], l2["pos"], c2w)
    img_o, proj, bins = O.render(l2["pos"], col_o, l2["opacity_raw"], sig_o, c2w, H, W, fx, fy, cx, cy,
                                 return_stages=True)
    assert torch.equal(sig_o, sig), f"{name}: sigma differs"
    assert torch.equal(col_o, col), f"{name}: colour differs"
    assert torch.equal(img_o, img_ref), f"{name}: oracle image is not bit-identical to the reference"
    report = {"image_bit_exact": True}
    if with_grad:
        grads_o = torch.autograd.grad((img_o * wimg).sum(), [l2[k] for k in PARAMS])
        for k, a, b in zip(PARAMS, grads_o, grads_ref):
            fin = torch.isfinite(b)
            assert torch.equal(torch.isfinite(a), fin), f"{name}: grad {k} finite pattern differs"
            err = (a[fin] - b[fin]).abs().max().item() / max(b[fin].abs().max().item(), 1e-30)
            assert err < 1e-5, f"{name}: oracle grad {k} off by {err}"
            report[f"grad_{k}_relerr"] = err
            report[f"grad_{k}_nonfinite"] = int((~fin).sum())

    # --- fp64 run of the oracle: the arbiter for fp32 round-off (threshold flips, ill-conditioned
    # covariances): a result counts as matching when it is as close to this as the fp32 reference is.
    l3 = {k: sc[k].double().clone().requires_grad_(with_grad) for k in PARAMS}
    c2w64 = c2w.double()
    sig64 = O.build_sigma_from_params(l3["scale_raw"], l3["q_raw"])
    col64 = O.evaluate_sh(l3["f_dc"], l3["f_rest"], l3["pos"], c2w64)
    img64 = O.render(l3["pos"], col64, l3["opacity_raw"], sig64, c2w64, H, W, fx, fy, cx, cy)
    out["image64"] = img64.detach().numpy()
    d = (img_ref.detach().double() - img64.detach()).abs()
    report["ref32_vs_64_image_max"] = float(d.max())
    report["ref32_vs_64_image_n_gt_1e-4"] = int((d > 1e-4).sum())
    if with_grad:
        grads64 = torch.autograd.grad((img64 * wimg.double()).sum(), [l3[k] for k in PARAMS])
        for k, g64, g32 in zip(PARAMS, grads64, grads_ref):
            out[f"grad64_{k}"] = g64.numpy()
            report[f"ref32_vs_64_grad_{k}"] = float((g32.double() - g64).abs().max() / g64.abs().max())

    # --- intermediates -----------------------------------------------------------------------------
    out.update(ids=proj.ids.numpy().astype(np.int32), u=proj.u.detach().numpy(), v=proj.v.detach().numpy(),
               z=proj.z.detach().numpy(), opacity=proj.opacity.detach().numpy(),
               lam_max=proj.lam_max.detach().numpy(),
               radius=proj.radius.numpy().astype(np.int32), rect=proj.rect.numpy().astype(np.int32),
               tiles_touched=proj.tiles_touched.numpy().astype(np.int32),
               conic=proj.conic.detach().numpy()[:, [0, 0, 1], [0, 1, 1]],
               list_tile=bins.tile_ids.numpy().astype(np.int32),
               list_id=proj.ids[bins.ranks].numpy().astype(np.int32),
               uniq_tiles=bins.uniq_tiles.numpy().astype(np.int32),
               start=bins.start.numpy().astype(np.int32), end=bins.end.numpy().astype(np.int32))
    report.update(proj.stage_counts)
    report["depth_ties"] = int((proj.z[1:] == proj.z[:-1]).sum())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    return report


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    summary = {"torch": torch.__version__, "reference_path": REF, "cases": {}}
    for name, (fn, cam_kw, wg) in CASES.items():
        summary["cases"][name] = run_case(name, fn, cam_kw, wg)
        print(name, summary["cases"][name], flush=True)
    # empty-result behaviours (render.py:109-112, 235-236)
    sc = O.make_scene(50, seed=2)
    cam = O.make_camera(64, 64)
    sig = ref_gaussian.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    col = ref_sh.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], cam["c2w"])
    z = ref_render.render(sc["pos"], col, torch.full((50,), -20.0), sig, cam["c2w"], 64, 64, 57.6, 57.6, 32., 32.)
    assert z.shape == (64, 64, 3) and float(z.abs().max()) == 0.0
    summary["all_transparent_returns_zero_image"] = True
    try:
        far = sc["pos"].clone()
        far[:, 0] = 2.0 + 0.001 * torch.arange(50)      # inside the guard band, AABB off the screen
        far[:, 1] *= 0.1
        far[:, 2] = 0.0
        tiny = torch.full((50, 3), -9.0)
        sig2 = ref_gaussian.build_sigma_from_params(tiny, sc["q_raw"])
        ref_render.render(far, col, sc["opacity_raw"] * 0 + 2, sig2, cam["c2w"], 64, 64, 57.6, 57.6, 32., 32.)
        summary["offscreen_raises"] = False
    except Exception as e:  # noqa: BLE001
        summary["offscreen_raises"] = str(e)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "offscreen_case.npz"),
                            pos=far.numpy(), scale_raw=tiny.numpy(), q_raw=sc["q_raw"].numpy(),
                            c2w=cam["c2w"].numpy())
    json.dump(summary, open(os.path.join(ROOT, "tests", "golden", "SUMMARY.json"), "w"), indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
