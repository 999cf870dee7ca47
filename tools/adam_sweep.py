"""Why was the opt-in fused optimizer slower than torch's through the reference's training script at small N
(profiles/r01_reference_scripts_run.log: 15.2 vs 54.0 it/s at N = 5k)?  The loop of scripts/train.py:530-541 -
backward, empty_cache, clip_grad_norm_, step, empty_cache - with synthetic gradients, N = 5k ... 1M, stock
(torch.optim.Adam + torch.nn.utils.clip_grad_norm_) against fused (b200gs.FusedAdam + b200gs.clip_grad_norm_), with and
without the script's torch.cuda.empty_cache() calls, wall clock per iteration with a final synchronize; the time spent
inside empty_cache is reported separately.  One JSON line per configuration."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch  # noqa: E402
import b200gs  # noqa: E402

SHAPES = {"pos": 3, "opacity_raw": 1, "f_dc": 3, "f_rest": 45, "scale_raw": 3, "q_raw": 4}


def run(n, fused, empty, iters=100):
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    params = {k: torch.nn.Parameter(torch.randn((n, w) if w > 1 else (n,), device=dev, generator=g)) for k, w in SHAPES.items()}
    grads = {k: torch.randn_like(p) * 1e-3 for k, p in params.items()}
    Opt = b200gs.FusedAdam if fused else torch.optim.Adam
    clip = b200gs.clip_grad_norm_ if fused else torch.nn.utils.clip_grad_norm_
    opt = Opt([{"params": [p], "lr": 1e-3, "name": k} for k, p in params.items()], lr=1e-3, eps=1e-15)
    t_empty = 0.0

    def it():
        nonlocal t_empty
        opt.zero_grad()
        for k, p in params.items():
            p.grad = grads[k].clone()              # stands for backward(): fresh gradient tensors every iteration
        if empty:
            t0 = time.perf_counter(); torch.cuda.empty_cache(); t_empty += time.perf_counter() - t0
        clip(params["pos"], max_norm=1.0)
        opt.step()
        if empty:
            t0 = time.perf_counter(); torch.cuda.empty_cache(); t_empty += time.perf_counter() - t0
    for _ in range(10):
        it()
    torch.cuda.synchronize()
    t_empty = 0.0
    t0 = time.perf_counter()
    for _ in range(iters):
        it()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"N": n, "optimizer": "fused" if fused else "stock", "empty_cache": empty, "ms_per_iter": dt / iters * 1e3,
            "ms_in_empty_cache_per_iter": t_empty / iters * 1e3}


if __name__ == "__main__":
    for n in (5_000, 50_000, 500_000, 1_000_000):
        for empty in (False, True):
            for fused in (False, True):
                print(json.dumps(run(n, fused, empty)), flush=True)
