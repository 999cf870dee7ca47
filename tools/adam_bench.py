"""adam_step stand-alone at N = 1M: b200gs.FusedAdam next to torch.optim.Adam (foreach) and torch's fused=True (CUDA events)."""
import sys, os
sys.path.insert(0, "3d-gaussian-splatting-for-novel-view-synthesis_b200")
import torch, b200gs
n = 1_000_000
shapes = dict(pos=(n, 3), opacity_raw=(n,), f_dc=(n, 3), f_rest=(n, 45), scale_raw=(n, 3), q_raw=(n, 4))
ps = {k: torch.randn(*s, device="cuda").requires_grad_(True) for k, s in shapes.items()}
opt = b200gs.FusedAdam([{"params": [p], "lr": 1e-3} for p in ps.values()], eps=1e-15)
ref = {k: v.detach().clone().requires_grad_(True) for k, v in ps.items()}
opt_r = torch.optim.Adam([{"params": [p], "lr": 1e-3} for p in ref.values()], eps=1e-15)
opt_f = torch.optim.Adam([{"params": [p], "lr": 1e-3} for p in [v.detach().clone().requires_grad_(True) for v in ps.values()]], eps=1e-15, fused=True)
for o, pp in ((opt, ps.values()), (opt_r, ref.values()), (opt_f, [p for g in opt_f.param_groups for p in g["params"]])):
    for p in pp:
        p.grad = torch.randn_like(p)
    for _ in range(3):
        o.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        o.step()
    e1.record()
    torch.cuda.synchronize()
    print(type(o).__name__, getattr(o, "defaults", {}).get("fused"), round(e0.elapsed_time(e1) / 20 * 1e3, 1), "us/step", round(28 * 59 * n / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9), "GB/s")
