"""Generates tests/golden/densify_*.npz by running the REFERENCE's own GaussianModel.densify_and_prune
(scripts/train.py:89-195) on the CPU on small seeded parameter sets, with the torch.randn_like draw of _split_points
recorded so that the CUDA path can be fed the same noise.  Cases: prune + split, prune + clone, prune only (grads=None),
nothing to do, and split + clone in one call (the reference raises IndexError: recorded as such).

Test infrastructure only; needs /root/reference (build container).  python oracle/make_golden_densify.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = os.environ.get("B200GS_REFERENCE_ROOT", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PARAMS = ("pos", "opacity_raw", "f_dc", "f_rest", "scale_raw", "q_raw")


def make(n, seed, scale_centre, grad_scale):
    g = torch.Generator().manual_seed(seed)
    p = {"pos": torch.randn(n, 3, generator=g), "opacity_raw": torch.randn(n, generator=g) * 3.0,
         "f_dc": torch.randn(n, 3, generator=g), "f_rest": torch.randn(n, 45, generator=g) * 0.1,
         "scale_raw": torch.randn(n, 3, generator=g) * 0.3 + scale_centre, "q_raw": torch.randn(n, 4, generator=g)}
    # keep every decision well away from its threshold (CPU and GPU exp / sigmoid differ in the last bit)
    op = torch.sigmoid(p["opacity_raw"])
    p["opacity_raw"][(op - 0.01).abs() < 2e-3] = 2.0
    grads = {"pos": torch.randn(n, 3, generator=g) * grad_scale, "opacity_raw": torch.randn(n, generator=g)}
    gn = grads["pos"].norm(dim=-1)
    grads["pos"][(gn - 0.01).abs() < 2e-3] *= 3.0
    smax = torch.exp(p["scale_raw"]).max(dim=-1)[0]
    p["scale_raw"][(smax - 0.01).abs() < 1e-3] += 0.5
    return p, grads


def run_case(ref_train, name, p, grads, **kw):
    model = ref_train.GaussianModel({k: v.clone() for k, v in p.items()}, device="cpu")
    noise = []
    real = torch.randn_like

    def recording(x, *a, **k):
        r = real(x, *a, **k)
        noise.append(r.clone())
        return r
    torch.randn_like = recording
    raised = ""
    g_in = None if grads is None else {k: v.clone() for k, v in grads.items()}
    try:
        model.densify_and_prune(g_in, **kw)
    except IndexError as e:
        raised = str(e)
    finally:
        torch.randn_like = real
    out = {"in_" + k: p[k].numpy() for k in PARAMS}
    out.update({"out_" + k: getattr(model, k).detach().numpy() for k in PARAMS})
    if grads is not None:
        out.update({"gin_" + k: v.numpy() for k, v in grads.items()})
        out.update({"gout_" + k: v.numpy() for k, v in g_in.items()})
    out["noise"] = noise[0].numpy() if noise else np.zeros((0, 3), np.float32)
    out["raised"] = np.array(raised)
    out["kw"] = np.array([kw.get("opacity_threshold", 0.01), kw.get("max_grad", 0.01), kw.get("scale_threshold", 0.01)])
    np.savez_compressed(os.path.join(OUT, f"densify_{name}.npz"), **out)
    print(name, "N", p["pos"].shape[0], "->", getattr(model, "pos").shape[0], "noise rows", out["noise"].shape[0],
          "raised" if raised else "")


def main():
    spec = importlib.util.spec_from_file_location("ref_train", os.path.join(REF, "scripts", "train.py"))
    ref_train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_train)
    # large Gaussians (exp(scale) ~ 0.05 > 0.01) with hot gradients -> split only
    p, g = make(700, 1, -3.0, 0.02)
    run_case(ref_train, "split", p, g)
    # small Gaussians (exp(scale) ~ 0.0025 <= 0.01) -> clone only
    p, g = make(700, 2, -6.0, 0.02)
    run_case(ref_train, "clone", p, g)
    # prune only
    p, g = make(500, 3, -3.0, 0.02)
    run_case(ref_train, "prune_only", p, None)
    # cold gradients: prune, nothing else; different thresholds
    p, g = make(300, 4, -3.0, 1e-4)
    run_case(ref_train, "cold", p, g, opacity_threshold=0.05, max_grad=0.02, scale_threshold=0.02)
    # both kinds present: the reference raises
    p, g = make(400, 5, -4.6, 0.02)
    run_case(ref_train, "both", p, g)


if __name__ == "__main__":
    main()
