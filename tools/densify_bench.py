"""densify / prune / split at N = 1M: b200gs.densify_tensors (csrc/densify.cu) next to the reference's formulation
(boolean-mask gathers + torch.cat per tensor, scripts/train.py:109-195) run with torch ops on the same GPU."""
import os, sys, time
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch, b200gs
from oracle import gs_oracle as O
n = 1_000_000
sc = {k: v.cuda() for k, v in O.make_scene(n, seed=0, log_scale=-4.2).items()}     # exp(scale) around 0.015: splits, no clones
g = torch.randn(n, 3, device="cuda") * 0.02
g[torch.exp(sc["scale_raw"]).max(dim=-1)[0] <= 0.01] = 0.0                          # keep the clone set empty (the reference raises otherwise)
PARAMS = ("pos", "opacity_raw", "f_dc", "f_rest", "scale_raw", "q_raw")

def torch_way():
    p = {k: sc[k] for k in PARAMS}
    prune = torch.sigmoid(p["opacity_raw"]) < 0.01
    if prune.any():
        p = {k: v[~prune] for k, v in p.items()}
    gg = g[~prune]
    gn = gg.norm(dim=-1)
    smax = torch.exp(p["scale_raw"]).max(dim=-1)[0]
    split = (smax > 0.01) & (gn > 0.01)
    if split.any():
        new = {k: v[split].clone() for k, v in p.items()}
        new["pos"] = new["pos"] + torch.randn_like(new["pos"]) * torch.exp(p["scale_raw"][split]) * 0.1
        new["scale_raw"] = new["scale_raw"] - 0.5
        p = {k: torch.cat([p[k], new[k]], dim=0) for k in PARAMS}
    clone = (smax <= 0.01) & (gn > 0.01)
    if clone.any():
        p = {k: torch.cat([p[k], p[k][clone]], dim=0) for k in PARAMS}
    return p

def mine():
    return b200gs.densify_tensors(sc, g)[0]

for name, fn in (("torch ops (the reference's formulation)", torch_way), ("b200gs.densify_tensors", mine)):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        out = fn()
    torch.cuda.synchronize()
    print(f"{name:42s} {(time.perf_counter() - t0) * 100:.2f} ms per call, {n} -> {out['pos'].shape[0]} rows")
