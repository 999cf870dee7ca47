"""Host-side cost of clip + optimizer step at a small N (5 000 Gaussians): torch vs b200gs, with and without empty_cache()."""
import sys, time
sys.path.insert(0, "3d-gaussian-splatting-for-novel-view-synthesis_b200")
import torch, b200gs
n = 5000
shapes = dict(pos=(n, 3), opacity_raw=(n,), f_dc=(n, 3), f_rest=(n, 45), scale_raw=(n, 3), q_raw=(n, 4))
for name, Opt, clip in (("torch", torch.optim.Adam, torch.nn.utils.clip_grad_norm_), ("b200gs", b200gs.FusedAdam, b200gs.clip_grad_norm_)):
    ps = {k: torch.nn.Parameter(torch.randn(*s, device="cuda")) for k, s in shapes.items()}
    opt = Opt([{"params": [p], "lr": 1e-3} for p in ps.values()], lr=0.01, eps=1e-15)
    for ec in (False, True):
        for it in range(3):
            for p in ps.values(): p.grad = torch.randn_like(p)
            clip(ps["pos"], max_norm=1.0); opt.step()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for it in range(100):
            opt.zero_grad()
            for p in ps.values(): p.grad = torch.randn_like(p)
            if ec: torch.cuda.empty_cache()
            clip(ps["pos"], max_norm=1.0)
            opt.step()
            if ec: torch.cuda.empty_cache()
        torch.cuda.synchronize()
        print(name, "empty_cache" if ec else "plain", round((time.perf_counter() - t0) * 10, 2), "ms/iter")
